set -x
mkdir -p gpurun_out
python -m pytest tests/test_merge3.py -m gpu -x -q > gpurun_out/r3q_merge3.log 2>&1; tail -5 gpurun_out/r3q_merge3.log
