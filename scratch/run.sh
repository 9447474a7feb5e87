set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; tail -4 gpurun_out/r2k_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; tail -2 gpurun_out/r2k_smoke.log
