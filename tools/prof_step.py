"""Profiling target: a few eager steps of the headline hot path (bench.py's HotPath) and optionally one label-prop call.
Used under ncu (see profiles/README.md); prints nothing that is a bench number."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--lp", action="store_true")
ap.add_argument("--walk-general", action="store_true")
ap.add_argument("--chain", action="store_true", help="one chain of autograd operators instead of the micro-batch pipeline")
ap.add_argument("--kineto", action="store_true", help="per-kernel device times in situ from torch.profiler")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
if a.lp:
    bench.LP["n_tgt"] = 2
    print(bench.label_prop_bench(dev, cpu=False, gpu_baseline=False)["ms_per_frame"])
else:
    from sapienza_video_contrastive_b200 import ops
    ops.set_async_wgrad(True)
    hp = bench.HotPath(dev, 0, use_graph=False, sizes=None if a.chain else [4, 4, 4, 4, 4], pool_sms=0 if a.chain else 100,
                       head_splits=1 if a.chain else 4, pool_sms_bwd=0)
    for _ in range(a.steps):
        hp.step()
    torch.cuda.synchronize()
    if a.kineto:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10):
                hp.step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
print("ok")
