set -x
mkdir -p gpurun_out
P=r2f
python -m pytest tests -m gpu -x -q > gpurun_out/${P}_pytest.log 2>&1; tail -4 gpurun_out/${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; tail -2 gpurun_out/${P}_smoke.log
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/${P}_prebench.json 2> gpurun_out/${P}_prebench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches.csv $B > gpurun_out/${P}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton_static_kernel -s 10 -c 2 -f -o gpurun_out/${P}_prof_static $B > gpurun_out/${P}_ncu_static.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton_sorted_kernel -s 4 -c 2 -f -o gpurun_out/${P}_prof_sorted $B > gpurun_out/${P}_ncu_sorted.log 2>&1
python scratch/k4_hbm.py > gpurun_out/${P}_k4_hbm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:newton_seq_kernel -s 4 -c 1 -f -o gpurun_out/${P}_prof_seq_k4 python scratch/k4_hbm.py > gpurun_out/${P}_ncu_seq.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton_linear_kernel -s 4 -c 1 -f -o gpurun_out/${P}_prof_linear_k4 python scratch/k4_hbm.py > gpurun_out/${P}_ncu_lin.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton_seq_kernel -s 3 -c 1 -f -o gpurun_out/${P}_prof_seq_k1x8 python scratch/kbench.py 5 1 1048576 8 > gpurun_out/${P}_ncu_seq8.log 2>&1
python -c "
import importlib; print(importlib.import_module('2d_geometry_constraint_solver_b200').capi.load().gcs_b200_version().decode())" > gpurun_out/${P}_ncu_libversion.txt
ls -la gpurun_out/${P}_prof_*.ncu-rep; cat gpurun_out/${P}_ncu_libversion.txt
