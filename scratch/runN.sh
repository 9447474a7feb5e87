set -x
mkdir -p gpurun_out
NS="${NS:-2}"
for N in $NS; do
  if [ $N = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N"; fi
  ( time $L bench.py --impl reference --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2g_ref_n$N.json 2> gpurun_out/r2g_ref_n$N.err
  ( time $L bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err
  tail -4 gpurun_out/r2g_bench_n$N.err
  python - <<PY
import json
b=json.loads(open('gpurun_out/r2g_bench_n$N.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2g_ref_n$N.json').read().strip().splitlines()[-1])
e=b['e2e']
print('N=$N value %.4e e2e %.4e %.2f ms ceil %.2f typ %.2f | ref %.3e cores %d | value/ref %.0f e2e/ref %.1f same_config %s'%(b['value'],e['value'],e['ms_per_step'],e['pcie']['ceiling_ms_per_step'],e['pcie']['typical_copy_only_ms_per_step'],r['value'],r['cpu_baseline']['cores'],b['value']/r['value'],e['value']/r['value'],r['config']==b['config']))
c=b['configs']; print('  sweep %.4e %.3f ms  ms8 %.4e  K4 hbm %.3f at4m %.3f sharded %s  stale %s'%(c['sweep64m']['value'],c['sweep64m']['ms_per_step'],c['multistart8']['value'],c['kinds']['K4']['line_variant']['hbm_frac'],c['kinds']['K4']['at_4m']['hbm_frac'],c['sharded'],b['roofline']['traffic_capture_is_of_another_build']))
PY
done
