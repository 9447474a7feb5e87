python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for p in 2 4 8; do
  GCS_B200_PARTS=$p python bench.py --no-cpu-baseline --steps 10 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graph parts', $p, 'e2e ms', round(d['e2e']['ms_per_step'],3), 'M/s', round(d['e2e']['value']/1e6,1), d['e2e']['gpu_launches'])"
done
GCS_B200_NOGRAPH=1 python bench.py --no-cpu-baseline --steps 10 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nograph parts 4 e2e ms', round(d['e2e']['ms_per_step'],3), 'M/s', round(d['e2e']['value']/1e6,1), d['e2e']['gpu_launches'])"
