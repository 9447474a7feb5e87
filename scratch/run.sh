set -x
for v in 3 7; do python scratch/kbench.py $v 1,2,3,5 2>&1 | grep variant; done
python -m pytest tests/test_gpu_relaxed.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
