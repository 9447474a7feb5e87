set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2j_pytest.log 2>&1; tail -5 gpurun_out/r2j_pytest.log
python scratch/norerun.py > gpurun_out/r2j_norerun.log 2>&1; tail -30 gpurun_out/r2j_norerun.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; tail -3 gpurun_out/r2j_bench_n1.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2j_ref_n1.json 2> gpurun_out/r2j_ref_n1.err; cut -c1-200 gpurun_out/r2j_ref_n1.json
GCS_HOST_TRACE=1 python profiles/sketch_bench.py 100000 2>&1 | grep -v "wave launch" | tail -4 | cut -c1-900
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
