set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_host.py tests/test_merge3.py -x -q > gpurun_out/r2x_host.log 2>&1; tail -5 gpurun_out/r2x_host.log
for so in libgcs_host_prev.so libgcs_host.so; do
  echo "== $so"
  GCS_HOST_TRACE=1 GCS_HOST_SO=$so python scratch/sketch_time.py 100000 6 > gpurun_out/r2x_trace_$so.log 2>&1
  grep -E "^rc" gpurun_out/r2x_trace_$so.log | sed 's/.*decompose_us/decompose_us/'
  grep -E "peel:|plan:" gpurun_out/r2x_trace_$so.log | tail -6
done
