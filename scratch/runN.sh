set -x
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_r1s.json 2> gpurun_out/bench_n${N}_r1s.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload sweep64m --steps 5 --warmup 3 > gpurun_out/bench_sweep64m_n${N}_r1s.json 2>> gpurun_out/bench_n${N}_r1s.err
tail -c 300 gpurun_out/bench_n${N}_r1s.err
