nvidia-smi topo -m 2>&1 | head -20
lscpu | grep -i "numa\|socket\|^CPU(s)\|Model name" | head; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; nproc
for per in 0.001 0.05; do
GCS_BENCH_CLOCK_PERIOD=$per python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('period $per', 'e2e ms', d['e2e']['ms_per_step'], 'e2e', d['e2e']['value']/1e9)"
done
