timeout 1000 python scratch/soak_relaxed_scaled.py 2097152 > gpurun_out/soak_relaxed_scaled.log 2>&1; tail -26 gpurun_out/soak_relaxed_scaled.log
