python -m pytest tests/test_gpu_golden.py -x -q 2>&1 | tail -8
