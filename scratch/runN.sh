N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/bench_n${N}_r1g.err | tail -1 > gpurun_out/bench_n${N}_r1g.json; cut -c1-220 gpurun_out/bench_n${N}_r1g.json; grep -o '"e2e": {"value": [0-9.e+]*' gpurun_out/bench_n${N}_r1g.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --workload sweep64m --steps 10 2>/dev/null | tail -1 > gpurun_out/bench_sweep64m_n${N}_r1g.json; cut -c1-220 gpurun_out/bench_sweep64m_n${N}_r1g.json
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded" 2>&1 | tail -2
