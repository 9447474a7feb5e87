set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d_pytest.log 2>&1; tail -8 gpurun_out/r2d_pytest.log
python -m pytest tests/test_gpu_margins.py -m gpu -q -s 2>&1 | grep "decided by" > gpurun_out/r2d_margins.log; cat gpurun_out/r2d_margins.log
for mb in 4 8 16 32; do
  GCS_B200_STAGE_MB=$mb python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read()); e=b['e2e']; print('stage_mb $mb e2e', e['value'], e['ms_per_step'], 'frac', e['pcie']['frac_of_ceiling'], 'dense', e['dense_round1_format']['value'], 'k1', b['roofline']['launch_ms'], 'k5', b['roofline']['second_kernel']['launch_ms'], 'value', b['value'])"
done > gpurun_out/r2d_stage.log 2>&1; cat gpurun_out/r2d_stage.log
python bench.py --workload sweep64m --steps 5 --warmup 3 > gpurun_out/r2d_sweep.json 2>gpurun_out/r2d_sweep.err; cut -c1-400 gpurun_out/r2d_sweep.json
GCS_B200_LIB=$PWD/build/alt/libgcs_b200_nocareful.so python bench.py --workload sweep64m --steps 5 --warmup 3 2>/dev/null | cut -c1-300
