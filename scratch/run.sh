set -x
mkdir -p gpurun_out
GCS_E2E_TAIL_ROWS=0 GCS_B200_TRACE=1 python bench.py --no-cpu-baseline --no-extras --steps 5 > gpurun_out/r2z_trace.out 2> gpurun_out/r2z_trace.err
grep -n "trace" gpurun_out/r2z_trace.err | tail -60
