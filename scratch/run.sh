GCS_B200_TRACE=1 python scratch/trace.py 2> gpurun_out/trace_e2e.log; tail -40 gpurun_out/trace_e2e.log
python scratch/e2e_probe.py
for p in 3 6 8; do echo parts $p; GCS_B200_PARTS=$p python scratch/e2e_probe.py | grep both; done
