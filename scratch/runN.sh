set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
nproc >> gpurun_out/r2_topo.txt; numactl -H >> gpurun_out/r2_topo.txt 2>&1
PROBE_QUICK=1 python scratch/pcie_probe2.py > gpurun_out/r2_pcie_n1.jsonl 2> gpurun_out/r2_pcie.err
for N in 2 4 8; do
  PROBE_QUICK=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N scratch/pcie_probe2.py > gpurun_out/r2_pcie_n$N.jsonl 2>> gpurun_out/r2_pcie.err
done
cat gpurun_out/r2_pcie_n*.jsonl | cut -c1-330
python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -m gpu -q -k "sharded or index_ranges or several_devices" 2>&1 | tail -3
for N in 8 4 2; do
  ( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
  tail -4 gpurun_out/r2_bench_n$N.err
  python - <<PY
import json
b=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N value',b['value'],'e2e',b['e2e']['value'],b['e2e']['ms_per_step'],'pcie frac',b['e2e']['pcie']['frac_of_ceiling'],'ceil',b['e2e']['pcie']['ceiling_solves_per_s'],'agg GB/s',b['e2e']['pcie']['aggregate_gbs_at_ceiling'])
c=b['configs']; print(' sweep',c['sweep64m']['value'],c['sweep64m']['ms_per_step'],'ms8',c['multistart8']['value'],'sharded',c['sharded'])
PY
done
