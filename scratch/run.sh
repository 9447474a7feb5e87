set -x
mkdir -p gpurun_out
python -m pytest tests/test_merge3.py -m gpu -x -q -k golden > gpurun_out/r3r_merge3.log 2>&1; tail -8 gpurun_out/r3r_merge3.log
