set -x
mkdir -p gpurun_out
( time timeout 200 python scratch/soak_relaxed.py 2097152 0x5000 ) > gpurun_out/r3c_soak1.log 2>&1; tail -4 gpurun_out/r3c_soak1.log
( time timeout 120 python scratch/soak_relaxed_guesses.py 1048576 ) > gpurun_out/r3c_soak2.log 2>&1; tail -4 gpurun_out/r3c_soak2.log
( time timeout 120 python scratch/soak_relaxed_scaled.py 1048576 ) > gpurun_out/r3c_soak3.log 2>&1; tail -4 gpurun_out/r3c_soak3.log
( time SOAK_VARIANT=8 timeout 120 python scratch/soak_relaxed.py 1048576 0x6000 ) > gpurun_out/r3c_soak4.log 2>&1; tail -4 gpurun_out/r3c_soak4.log
python -m pytest tests -m gpu -x -q > gpurun_out/r3c_pytest.log 2>&1; tail -5 gpurun_out/r3c_pytest.log
python scratch/kbench.py 5 1,2,3,5 524288 2 > gpurun_out/r3c_kbench.log 2>&1
cat gpurun_out/r3c_kbench.log
