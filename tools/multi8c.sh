for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 200 --warmup 5 --no-extras --no-cpu 2>/dev/null > gpurun_out/r2_n${n}_final.json
python -c "
import sys,json
d=json.loads([l for l in open('gpurun_out/r2_n${n}_final.json') if l.startswith('{')][-1]); print('bench n$n', round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d.get('parts'), d.get('pool_sms'))"
done
