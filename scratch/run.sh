python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scratch/kbench.py 1 1,2,3,4,5
python scratch/kbench.py 2 1,5
