python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python scratch/e2e_probe.py | grep both
