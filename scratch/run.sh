for n in 32768 131072 262144 524288 1048576 2097152; do python scratch/kbench.py 6 1,5 $n 2>&1 | grep variant | cut -c1-75; done
