set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; tail -c 300 gpurun_out/bench_r1n.err
python scratch/soak_relaxed.py 2097152 > gpurun_out/soak_relaxed.log 2>&1; tail -3 gpurun_out/soak_relaxed.log
