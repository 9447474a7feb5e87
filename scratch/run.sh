python bench.py > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; cut -c1-200 gpurun_out/bench_r1h.json
python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_r1h.json; cut -c1-200 gpurun_out/bench_ref_r1h.json
python bench.py --workload multistart8 --steps 5 2>/dev/null | tail -1 > gpurun_out/bench_multistart8_r1h.json
python bench.py --workload sweep64m --steps 5 2>/dev/null | tail -1 > gpurun_out/bench_sweep64m_n1_r1h.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll_r1h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton --launch-skip 3 -c 1 -o gpurun_out/prof_k1_r1h -f python scratch/kbench.py 3 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton --launch-skip 3 -c 1 -o gpurun_out/prof_k5_r1h -f python scratch/kbench.py 3 5 > /dev/null 2>&1
python profiles/sketch_bench.py 100000 2>/dev/null | tail -1 > gpurun_out/sketch_bench_r1h.json; cut -c1-400 gpurun_out/sketch_bench_r1h.json
