python scratch/soak_relaxed.py 4194304 0x100000 > gpurun_out/soak_relaxed5.log 2>&1; tail -1 gpurun_out/soak_relaxed5.log; grep VIOLATION gpurun_out/soak_relaxed5.log | cut -c1-300
