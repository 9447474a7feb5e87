set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c_pytest.log 2>&1; tail -5 gpurun_out/r2c_pytest.log
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 1500 gpurun_out/r2c_bench.err; cut -c1-600 gpurun_out/r2c_bench.json
( time python bench.py --impl reference --steps 5 --warmup 2 ) > gpurun_out/r2c_ref.json 2> gpurun_out/r2c_ref.err; tail -c 600 gpurun_out/r2c_ref.err; cut -c1-300 gpurun_out/r2c_ref.json
