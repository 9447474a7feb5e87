set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1r.json 2> gpurun_out/bench_r1r.err; tail -c 200 gpurun_out/bench_r1r.err
python bench.py --workload multistart8 --steps 20 --warmup 3 > gpurun_out/bench_multistart8_r1r.json 2>> gpurun_out/bench_r1r.err
python bench.py --workload sweep64m --steps 5 --warmup 3 > gpurun_out/bench_sweep64m_n1_r1r.json 2>> gpurun_out/bench_r1r.err
python scratch/soak_relaxed_guesses.py 262144 > gpurun_out/soak_relaxed_guesses2.log 2>&1; tail -1 gpurun_out/soak_relaxed_guesses2.log
python scratch/soak_ref.py 1048576 5 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1r.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll_r1r.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton_static -s 6 -c 2 -f -o gpurun_out/prof_contracted_r1r python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1r.log 2>&1
ls -la gpurun_out/prof_contracted_r1r.ncu-rep
