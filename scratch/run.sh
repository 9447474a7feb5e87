GCS_BENCH_TEST_VIOLATION=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/viol.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g'%d['value'], d['config']['variant'], d['config']['contract_violation'], d['roofline']['kernel'], 'frac %.3f'%d['roofline']['frac'], d['two_stream_step'], 'e2e %.4g'%d['e2e']['value'])"
grep CONTRACT gpurun_out/viol.err | cut -c1-200
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g'%d['value'], d['config']['variant'], d['config']['contract_violation'], d['roofline']['kernel'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.4g'%d['e2e']['value'])"
