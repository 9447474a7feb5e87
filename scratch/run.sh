python -m pytest tests/test_merge3.py -m gpu -x -q 2>&1 | tail -5
