python scratch/kbench.py 8 1,2,3,5 2>&1 | grep variant | cut -c1-80
GCS_B200_LIB=build/libgcs_rpair8.so python scratch/kbench.py 8 1,2,3,5 2>&1 | grep variant | cut -c1-80
GCS_B200_LIB=build/libgcs_rpair10.so python scratch/kbench.py 8 1,2,3,5 2>&1 | grep variant | cut -c1-80
