python -m pytest tests/test_gpu_relaxed.py tests/test_gpu_golden.py -x -q 2>&1 | tail -4
python scratch/soak_relaxed.py 4194304 > gpurun_out/soak_relaxed4.log 2>&1; tail -1 gpurun_out/soak_relaxed4.log; grep VIOLATION gpurun_out/soak_relaxed4.log | cut -c1-300
python scratch/kbench.py 5 1,2,3,5 2>&1 | grep variant | cut -c1-170
