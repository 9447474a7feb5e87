set -x
mkdir -p gpurun_out
for L in "" build/alt_rlx6.so build/alt_rlx10.so build/alt_rlx12.so; do
  echo "== lib ${L:-default}" >> gpurun_out/r3l_occ.log
  GCS_B200_LIB=${L:+$PWD/$L} python scratch/kbench.py 5 1,5 524288 2 2>&1 | sed 's/literal re-runs.*//' >> gpurun_out/r3l_occ.log
  GCS_B200_LIB=${L:+$PWD/$L} python scratch/kbench.py 5 1,5 4194304 2 2>&1 | sed 's/literal re-runs.*//' >> gpurun_out/r3l_occ.log
  GCS_B200_LIB=${L:+$PWD/$L} python scratch/k4_hbm.py 2>&1 | grep "variant 5" >> gpurun_out/r3l_occ.log
done
cat gpurun_out/r3l_occ.log
