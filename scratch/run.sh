python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; cut -c1-300 gpurun_out/bench_r1g.json
python bench.py --workload multistart8 --steps 5 2>/dev/null | tail -1 > gpurun_out/bench_multistart8_r1g.json; cut -c1-200 gpurun_out/bench_multistart8_r1g.json
python bench.py --workload sweep64m --steps 5 2>/dev/null | tail -1 > gpurun_out/bench_sweep64m_n1_r1g.json; cut -c1-200 gpurun_out/bench_sweep64m_n1_r1g.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll_r1g.log 2>&1
tail -5 gpurun_out/launches_r1g.csv | cut -c1-250
