set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_host.py tests/test_merge3.py -x -q > gpurun_out/r2w_host.log 2>&1; tail -5 gpurun_out/r2w_host.log
GCS_HOST_TRACE=1 python scratch/sketch_time.py 100000 6 > gpurun_out/r2w_trace.log 2>&1
grep -E "^rc" gpurun_out/r2w_trace.log | sed 's/.*decompose_us/decompose_us/'
grep -E "peel:|plan:" gpurun_out/r2w_trace.log | tail -8
