set -x
mkdir -p gpurun_out
for v in 6 7 8; do
  SOAK_VARIANT=$v python scratch/soak_relaxed.py 1048576 0x4000 2>&1 | grep -v "within contract" | tail -3
  SOAK_VARIANT=$v python scratch/soak_relaxed_guesses.py 524288 2>&1 | grep -v "within contract" | tail -3
done > gpurun_out/r2r_soak_variants.log 2>&1; cat gpurun_out/r2r_soak_variants.log
