set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2m_pytest.log 2>&1; tail -5 gpurun_out/r2m_pytest.log
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2m_ref_n1.json 2> gpurun_out/r2m_ref_n1.err; cut -c1-160 gpurun_out/r2m_ref_n1.json
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; tail -3 gpurun_out/r2m_bench_n1.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2m_bench_n1.json').read().strip().splitlines()[-1])
print('value',b['value'],'e2e',b['e2e']['value'],b['e2e']['pcie']['frac_of_ceiling'],'k1',b['roofline']['launch_ms'],b['roofline']['frac'],'traffic stale',b['roofline']['traffic_capture_is_of_another_build'])
print('sketch',{k:v for k,v in b['configs']['sketch100k'].items() if k!='workload'})
PY
