python scratch/e2e_probe.py | grep both
python scratch/e2e_sub.py 131072 131072
python scratch/e2e_sub.py 131072 65536
python scratch/e2e_sub.py 65536 65536
GCS_B200_TRACE=1 python scratch/e2e_sub.py 131072 65536 2>&1 | grep trace | tail -42 | grep -E "up-begin|flags-end" | head -30
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
