python -m pytest tests/test_gpu_relaxed.py -x -q 2>&1 | tail -15
python scratch/kbench.py 5 1,2,3,5 2>&1 | grep variant | cut -c1-200
python scratch/soak_relaxed.py 1048576 > gpurun_out/soak_relaxed2.log 2>&1; tail -2 gpurun_out/soak_relaxed2.log
