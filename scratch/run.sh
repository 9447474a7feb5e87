set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1p.json 2> gpurun_out/bench_r1p.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r1p.json 2>> gpurun_out/bench_r1p.err
python bench.py --workload multistart8 --steps 20 --warmup 3 > gpurun_out/bench_multistart8_r1p.json 2>> gpurun_out/bench_r1p.err
python bench.py --workload sweep64m --steps 5 --warmup 3 > gpurun_out/bench_sweep64m_n1_r1p.json 2>> gpurun_out/bench_r1p.err
python bench.py --variant 0 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bitidentical_r1p.json 2>> gpurun_out/bench_r1p.err
tail -c 400 gpurun_out/bench_r1p.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll_r1p.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:newton_static -s 6 -c 2 -f -o gpurun_out/prof_contracted_r1p python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1p.log 2>&1
ls -la gpurun_out/prof_contracted_r1p.ncu-rep
