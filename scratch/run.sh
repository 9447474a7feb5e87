set -x
mkdir -p gpurun_out
P=r2h
python -m pytest tests -m gpu -x -q > gpurun_out/${P}_pytest.log 2>&1; tail -4 gpurun_out/${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; tail -2 gpurun_out/${P}_smoke.log
python bench.py --impl reference > gpurun_out/${P}_ref.json 2> gpurun_out/${P}_ref.err
python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; tail -3 gpurun_out/${P}_bench.err
( time timeout 150 python scratch/soak_relaxed_guesses.py 2097152 ) > gpurun_out/${P}_soak2.log 2>&1; tail -4 gpurun_out/${P}_soak2.log
( time timeout 150 python scratch/soak_relaxed_scaled.py 2097152 ) > gpurun_out/${P}_soak3.log 2>&1; tail -4 gpurun_out/${P}_soak3.log
