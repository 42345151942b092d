"""One large-graph walk fwd+bwd: python tools/prof_walk_large.py B N T [simt] [--kineto]
Plain: three steps (for ncu launch lists).  --kineto: per-kernel device times from torch.profiler (warm caches, in situ)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sapienza_video_contrastive_b200 import ops  # noqa: E402

B, N, T = (int(a) for a in sys.argv[1:4])
simt = "simt" in sys.argv[4:]
dev = torch.device("cuda", 0)
f = torch.randn(B, N, T, 128, device=dev, requires_grad=True)
ones = torch.ones(1, device=dev)


def step():
    f.grad = None
    q, loss, xent, acc = ops.walk(f, 0.07, 0.1, rng="philox", force_simt=simt)
    loss.backward(ones)
    return loss


for it in range(3):
    loss = step()
torch.cuda.synchronize()
print("loss", float(loss.detach()))
if "--kineto" in sys.argv:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for it in range(5):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=50))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("ms/step eager", e0.elapsed_time(e1) / 10)
