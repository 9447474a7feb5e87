set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2f_pytest.log 2>&1; tail -6 gpurun_out/r2f_pytest.log
python scratch/kbench.py 5 1,2,3,5 524288 > gpurun_out/r2f_kbench.log 2>&1
python scratch/kbench.py 7 1,5 524288 >> gpurun_out/r2f_kbench.log 2>&1
python scratch/kbench.py 5 1 1048576 8 >> gpurun_out/r2f_kbench.log 2>&1
python scratch/ksize.py 5 1 >> gpurun_out/r2f_kbench.log 2>&1
python scratch/k4_hbm.py >> gpurun_out/r2f_kbench.log 2>&1
cat gpurun_out/r2f_kbench.log
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 800 gpurun_out/r2f_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2f_bench.json'))
print('value',b['value'],'seq',b['sequential_launches']['value'],'k1',b['roofline']['launch_ms'],'frac',b['roofline']['frac'],'k5',b['roofline']['second_kernel']['launch_ms'])
print('e2e',b['e2e']['value'],b['e2e']['ms_per_step'],b['e2e']['pcie']['frac_of_ceiling'])
c=b['configs']
print('ms8',c['multistart8']['value'],c['multistart8']['roofline']['frac'],'sweep',c['sweep64m']['value'],c['sweep64m']['roofline']['frac'])
for k,v in c['kinds'].items(): print(k,{kk:(vv['launch_ms'],vv['hbm_frac'],vv['fp64_frac']) for kk,vv in v.items() if isinstance(vv,dict) and 'launch_ms' in vv})
print('sketch',{k:v for k,v in c['sketch100k'].items() if k!='workload'})
print('reruns',b['run']['literal_reruns'])
PY
GCS_HOST_TRACE=1 python profiles/sketch_bench.py 100000 2>&1 | grep -v "wave launch" | tail -5 | cut -c1-900
