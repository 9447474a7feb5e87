python -m pytest tests/test_gpu_host.py -m gpu -x -q 2>&1 | tail -5
python profiles/sketch_bench.py 100000 2>/dev/null | tail -1
