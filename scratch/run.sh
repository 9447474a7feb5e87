python scratch/soak.py 2097152 2>&1 | tail -30 > gpurun_out/soak_r1.log; cat gpurun_out/soak_r1.log
