python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload multistart8 --steps 5 > gpurun_out/bench_multistart8.json; cut -c1-330 gpurun_out/bench_multistart8.json; echo
python bench.py --workload sweep64m --steps 5 > gpurun_out/bench_sweep64m_n1.json; cut -c1-330 gpurun_out/bench_sweep64m_n1.json; echo
