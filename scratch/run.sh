python scratch/k4_hbm.py
