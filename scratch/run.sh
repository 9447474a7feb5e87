set -x
mkdir -p gpurun_out
python -m pytest tests/test_merge3.py -m gpu -x -q > gpurun_out/r3o_merge3.log 2>&1; tail -5 gpurun_out/r3o_merge3.log
python profiles/merge3_node_bench.py 10 > gpurun_out/r3o_m3node.json 2> gpurun_out/r3o_m3node.err; tail -3 gpurun_out/r3o_m3node.err
python profiles/merge3_node_bench.py 20 > gpurun_out/r3o_m3node20.json 2>> gpurun_out/r3o_m3node.err
nproc
