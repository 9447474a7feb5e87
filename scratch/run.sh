set -x
mkdir -p gpurun_out
nproc
GCS_HOST_TRACE=1 python scratch/sketch_time.py 100000 6 > gpurun_out/r2s_sketch.log 2>&1; grep -E "plan:|^rc|generated" gpurun_out/r2s_sketch.log
python -m pytest tests/test_gpu_host.py -x -q > gpurun_out/r2s_host.log 2>&1; tail -5 gpurun_out/r2s_host.log
