python -m pytest tests/test_gpu_host.py tests/test_merge3.py tests/test_sketch_io.py -m gpu -x -q 2>&1 | tail -2
python profiles/sketch_bench.py 100000 2>/dev/null | tail -1 > gpurun_out/sketch_bench_r1j.json; cut -c1-420 gpurun_out/sketch_bench_r1j.json
