run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/ar_probe.py 2>&1 | grep allreduce; }
run
NCCL_MAX_CTAS=16 run
NCCL_ALGO=NVLS run
NCCL_ALGO=NVLS NCCL_MAX_CTAS=16 run
NCCL_ALGO=Ring run
NCCL_ALGO=Tree run
NCCL_ALGO=NVLS python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 200 --warmup 5 --nccl-ctas 16 --no-extras --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('bench n8 NVLS ctas16', round(d['ms_per_step'],4), round(d['value']))"
