python profiles/merge3_bench.py 20000 | tail -1 > gpurun_out/merge3_bench.json; cat gpurun_out/merge3_bench.json
