python -m pytest tests/test_gpu_golden.py -x -q 2>&1 | tail -3
python scratch/soak_ref.py 1048576 5 2>&1 | tail -12
