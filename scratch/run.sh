set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_relaxed.py tests/test_gpu_soak.py tests/test_gpu_margins.py -m gpu -x -q > gpurun_out/r3f_pytest.log 2>&1; tail -5 gpurun_out/r3f_pytest.log
( time timeout 200 python scratch/soak_relaxed.py 1048576 0x8000 ) > gpurun_out/r3f_soak1.log 2>&1; tail -2 gpurun_out/r3f_soak1.log
( time timeout 120 python scratch/soak_relaxed_scaled.py 1048576 ) > gpurun_out/r3f_soak3.log 2>&1; tail -2 gpurun_out/r3f_soak3.log
( time timeout 120 python scratch/soak_relaxed_guesses.py 524288 ) > gpurun_out/r3f_soak2.log 2>&1; tail -2 gpurun_out/r3f_soak2.log
python scratch/kbench.py 5 1,2,3,4,5 524288 2 > gpurun_out/r3f_kbench.log 2>&1
cat gpurun_out/r3f_kbench.log
