set -x
mkdir -p gpurun_out
python scratch/repro_k4.py 2>&1 | tail -5
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2h_pytest.log 2>&1; tail -6 gpurun_out/r2h_pytest.log
python scratch/kbench.py 5 1,2,3,5 524288 > gpurun_out/r2h_kbench.log 2>&1
python scratch/k4_hbm.py >> gpurun_out/r2h_kbench.log 2>&1
python scratch/soak_relaxed.py 1048576 0x200 >> gpurun_out/r2h_kbench.log 2>&1
python scratch/soak_relaxed_guesses.py 262144 >> gpurun_out/r2h_kbench.log 2>&1
grep -v "variant [13]:" gpurun_out/r2h_kbench.log | cut -c1-260
python bench.py --workload sweep64m --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sweep', b['value'], b['ms_per_step'], b['roofline']['frac'])"
