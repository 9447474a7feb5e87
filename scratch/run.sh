set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_relaxed.py -m gpu -x -q -k "degenerate" > gpurun_out/r3j_pytest.log 2>&1; tail -15 gpurun_out/r3j_pytest.log
