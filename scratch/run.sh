python -m pytest tests/test_gpu_relaxed.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
python scratch/kbench.py 5 1,5 2>&1 | grep variant | cut -c1-90
python scratch/kbench.py 7 1,5 2>&1 | grep variant | cut -c1-90
