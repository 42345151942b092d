set -x
python -m pytest tests/test_gpu_ddp.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_ddp.log 2>&1; tail -3 gpurun_out/r2_ddp.log
for ctas in 4 8 16 32; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 --no-extras --no-cpu --nccl-ctas $ctas > gpurun_out/r2_n2_ctas$ctas.json 2> gpurun_out/r2_n2_ctas$ctas.err
  python -c "import json;d=json.load(open('gpurun_out/r2_n2_ctas$ctas.json'));print('ctas',$ctas,d['ms_per_step'],d['value'],d['e2e']['value'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 --no-extras --no-cpu --part-sizes "" --parts 1 > gpurun_out/r2_n2_chain.json 2> gpurun_out/r2_n2_chain.err
python -c "import json;d=json.load(open('gpurun_out/r2_n2_chain.json'));print('chain',d['ms_per_step'],d['value'])"
python bench.py --steps 200 --warmup 5 --no-extras --no-cpu > gpurun_out/r2_n1_pipe.json 2>gpurun_out/r2_n1_pipe.err; python -c "import json;d=json.load(open('gpurun_out/r2_n1_pipe.json'));print('n1',d['ms_per_step'],d['value'],d['e2e']['value'],d['gpu_launches'])"
