python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --workload sweep64m --steps 5 2>/dev/null | tail -1 > gpurun_out/bench_sweep64m_n2.json; cut -c1-260 gpurun_out/bench_sweep64m_n2.json
