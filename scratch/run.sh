python -m pytest tests/test_gpu_host.py tests/test_dropin_client.py tests/test_decomposition.py -m gpu -x -q 2>&1 | tail -3
python profiles/sketch_bench.py 100000 2>&1 | tail -2
