python -m pytest tests/test_gpu_relaxed.py -x -q 2>&1 | tail -3
python scratch/kbench.py 9 1,2,3,5 2>&1 | grep variant | cut -c1-170
python scratch/kbench.py 6 1,2,3,5 2>&1 | grep variant | cut -c1-80
python scratch/kbench.py 9 1,3 131072 8 2>&1 | grep variant | cut -c1-80
