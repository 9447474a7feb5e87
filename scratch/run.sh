timeout 800 python scratch/soak_relaxed_guesses.py 262144 > gpurun_out/soak_relaxed_guesses.log 2>&1; tail -30 gpurun_out/soak_relaxed_guesses.log
