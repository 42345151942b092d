set -x
python -m pytest tests/test_gpu_ddp.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_ddp2.log 2>&1; tail -3 gpurun_out/r2_ddp2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 --no-extras --no-cpu > gpurun_out/r2_n2_final.json 2> gpurun_out/r2_n2_final.err
python -c "
import json
l=[x for x in open('gpurun_out/r2_n2_final.json') if x.startswith('{')][-1]; d=json.loads(l); print('n2',d['ms_per_step'],d['value'],d['e2e']['value'],d.get('parts'))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
