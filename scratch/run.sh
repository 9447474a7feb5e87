set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_soak.py -m gpu -x -q > gpurun_out/r3k_pytest.log 2>&1; tail -8 gpurun_out/r3k_pytest.log
( time timeout 200 python scratch/soak_relaxed.py 2097152 0x9000 ) > gpurun_out/r3k_soak1.log 2>&1; tail -4 gpurun_out/r3k_soak1.log
( time timeout 200 python scratch/soak_relaxed.py 2097152 0xA000 ) > gpurun_out/r3k_soak2.log 2>&1; tail -4 gpurun_out/r3k_soak2.log
