python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -2
python scratch/kbench.py 3 1,2,3,5
python scratch/kbench.py 3 1 524288 8
