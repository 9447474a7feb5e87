set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_prebench.json 2> gpurun_out/r2_prebench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton_static_kernel -s 10 -c 2 -f -o gpurun_out/r2_prof_static python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_static.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton_sorted_kernel -s 4 -c 2 -f -o gpurun_out/r2_prof_sorted python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_sorted.log 2>&1
python scratch/k4_hbm.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:newton_seq_kernel -s 4 -c 1 -f -o gpurun_out/r2_prof_seq_k4 python scratch/k4_hbm.py > gpurun_out/r2_ncu_seq.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton_seq_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_seq_k1x8 python scratch/kbench.py 5 1 1048576 8 > gpurun_out/r2_ncu_seq8.log 2>&1
python -c "
import importlib; print(importlib.import_module('2d_geometry_constraint_solver_b200').capi.load().gcs_b200_version().decode())" > gpurun_out/r2_ncu_libversion.txt
ls -la gpurun_out/r2_prof_*.ncu-rep; cat gpurun_out/r2_ncu_libversion.txt
