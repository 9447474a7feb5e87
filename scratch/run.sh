python scratch/soak_ref.py 1048576 2>&1 | tail -14 > gpurun_out/soak_ref_r1.log; cat gpurun_out/soak_ref_r1.log
