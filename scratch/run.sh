set -x
mkdir -p gpurun_out
nvidia-smi -L
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.log 2>&1; tail -5 gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
for v in 0 1 2 5 7; do python scratch/kbench.py $v 1,2,3,4,5 524288; done > gpurun_out/r2a_kbench.log 2>&1; cat gpurun_out/r2a_kbench.log
python scratch/k4_hbm.py > gpurun_out/r2a_k4.log 2>&1; tail -20 gpurun_out/r2a_k4.log
python scratch/pcie_probe2.py > gpurun_out/r2a_pcie_n1.jsonl 2> gpurun_out/r2a_pcie_n1.err; cat gpurun_out/r2a_pcie_n1.jsonl | cut -c1-400
GCS_B200_LIB=$PWD/build/alt/libgcs_b200_sharedrt.so python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.err; cut -c1-1500 gpurun_out/r2a_bench.json
