# round-2 multi-GPU check: bench at N = 8 (and 4) with two NCCL CTA caps, plus the module-level end to end under DDP
for n in 8 4; do
  for ctas in 16 32; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 200 --warmup 5 --nccl-ctas $ctas > gpurun_out/r2_n${n}_ctas$ctas.json 2> gpurun_out/r2_n${n}_ctas$ctas.err
    python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_n${n}_ctas$ctas.json') if l.startswith('{')][-1]); print('n',$n,'ctas',$ctas,round(d['ms_per_step'],4),round(d['value']),round(d['e2e']['value']))"
  done
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 50 --warmup 5 --e2e-module > gpurun_out/r2_n8_module.json 2> gpurun_out/r2_n8_module.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_n8_module.json') if l.startswith('{')][-1]); print('module n8', d.get('e2e_module'))"
