set -x
mkdir -p gpurun_out
for off in 0x1000 0x2000 0x3000; do python scratch/soak_relaxed.py 2097152 $off 2>&1 | tail -1; done > gpurun_out/r2l_soak.log 2>&1
python scratch/soak_relaxed_guesses.py 1048576 2>&1 | tail -1 >> gpurun_out/r2l_soak.log
python scratch/soak_relaxed_scaled.py 1048576 2>&1 | tail -1 >> gpurun_out/r2l_soak.log
python scratch/soak.py 2>&1 | tail -1 >> gpurun_out/r2l_soak.log
cat gpurun_out/r2l_soak.log
GCS_BENCH_TEST_VIOLATION=1 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/r2l_viol.err | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('violation run:', b['run']['variant'], b['run']['contract_violation'], b['value'], b['roofline']['kernel'], b['config']['gpu_kernel_class'][:40])"
tail -2 gpurun_out/r2l_viol.err
python -m pytest tests/test_capi_load.py tests/test_gpu_parity.py -q -x 2>&1 | tail -2
