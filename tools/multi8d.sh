timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 200 --warmup 5 --no-extras --no-cpu > gpurun_out/r2_n8_final.json 2> gpurun_out/r2_n8_final.err
echo rc=$?
grep -v "^$" gpurun_out/r2_n8_final.err | grep -v "Warning\|warn" | tail -25 | cut -c1-300
python -c "
import sys,json
d=json.loads([l for l in open('gpurun_out/r2_n8_final.json') if l.startswith('{')][-1]); print('bench n8', round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d.get('parts'), d.get('pool_sms'))"
