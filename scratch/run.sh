set -x
mkdir -p gpurun_out
python scratch/small_launch.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:newton_static_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_small python scratch/small_launch.py > gpurun_out/r2_ncu_small.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:newton_static_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_small_nocc python scratch/small_launch.py > gpurun_out/r2_ncu_small2.log 2>&1
ls -la gpurun_out/r2_prof_small*
