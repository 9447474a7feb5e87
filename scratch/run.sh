set -x
mkdir -p gpurun_out
for L in "" build/alt_seq8.so build/alt_seq10.so; do
  echo "== lib ${L:-default}" >> gpurun_out/r3h_seq.log
  GCS_B200_LIB=${L:+$PWD/$L} python scratch/kbench.py 8 1,5 524288 2 >> gpurun_out/r3h_seq.log 2>&1
  GCS_B200_LIB=${L:+$PWD/$L} python scratch/kbench.py 8 1,5 4194304 2 >> gpurun_out/r3h_seq.log 2>&1
done
cat gpurun_out/r3h_seq.log
ncu --set full --clock-control none --import-source on -k regex:newton_seq_kernel -s 3 -c 1 -f -o gpurun_out/r3h_prof_seq_k1 python scratch/kbench.py 8 1 2097152 2 > gpurun_out/r3h_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton_seq_kernel -s 3 -c 1 -f -o gpurun_out/r3h_prof_seq_k5 python scratch/kbench.py 8 5 2097152 2 >> gpurun_out/r3h_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton_static_kernel -s 3 -c 1 -f -o gpurun_out/r3h_prof_static_k5 python scratch/kbench.py 5 5 2097152 2 >> gpurun_out/r3h_ncu.log 2>&1
ls -la gpurun_out/r3h_prof*
