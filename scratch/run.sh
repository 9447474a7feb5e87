python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
python scratch/kbench.py 3 1,2,3,4,5
echo t256; GCS_B200_LIB=scratch/libgcs_t256.so python scratch/kbench.py 3 1,2,3,5
