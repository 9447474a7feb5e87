python bench.py --workload sweep64m --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-900
python bench.py --workload multistart8 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-700
