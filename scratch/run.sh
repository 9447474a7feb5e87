set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_relaxed.py -m gpu -x -q -k "linear or k4 or K4" > gpurun_out/r3g_pytest_linear.log 2>&1; tail -15 gpurun_out/r3g_pytest_linear.log
python -m pytest tests/test_gpu_relaxed.py tests/test_gpu_soak.py tests/test_gpu_margins.py tests/test_capi_load.py -m gpu -x -q > gpurun_out/r3g_pytest.log 2>&1; tail -5 gpurun_out/r3g_pytest.log
python scratch/k4_hbm.py > gpurun_out/r3g_k4.log 2>&1; cat gpurun_out/r3g_k4.log
( time timeout 120 python scratch/soak_relaxed_guesses.py 1048576 ) > gpurun_out/r3g_soak2.log 2>&1; grep "K4\|violate" gpurun_out/r3g_soak2.log
( time timeout 120 python scratch/soak_relaxed_scaled.py 1048576 ) > gpurun_out/r3g_soak3.log 2>&1; grep "K4\|violate" gpurun_out/r3g_soak3.log
python scratch/kbench.py 5 1 1048576 8 > gpurun_out/r3g_kbench.log 2>&1; cat gpurun_out/r3g_kbench.log
