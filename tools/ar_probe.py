"""Time the 44.97 MB fp32 gradient all-reduce alone (CUDA events, max over ranks) for the NCCL settings in the environment."""
import os
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
x = torch.zeros(11_176_512 + 65_536, device="cuda")
for _ in range(10):
    dist.all_reduce(x)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    dist.all_reduce(x)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 50], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("allreduce 44.97MB world=%d algo=%s ctas=%s: %.1f us" % (world, os.environ.get("NCCL_ALGO", "auto"), os.environ.get("NCCL_MAX_CTAS", "auto"), float(t) * 1e3))
dist.destroy_process_group()
