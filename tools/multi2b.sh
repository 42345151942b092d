for cfg in "108 16" "92 16" "84 16" "92 24" "100 12"; do
  set -- $cfg
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 300 --warmup 5 --no-extras --no-cpu --pool-sms $1 --nccl-ctas $2 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('n2 pool_sms $1 ctas $2', round(d['ms_per_step'],4), round(d['value']))"
done
python -m pytest tests/test_gpu_ddp.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
