"""Profiling target: superpixel-graph steps (BASELINE configs[2] shape).  Plain: for ncu launch lists; --kineto: per-kernel
device times in situ from torch.profiler."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.sweep as sw  # noqa: E402

if "--kineto" in sys.argv:
    from torch.profiler import ProfilerActivity, profile
    sw.superpixel_point(B=8, T=8, SP=196)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        sw.superpixel_point(B=8, T=8, SP=196)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))
else:
    print(sw.superpixel_point(B=8, T=8, SP=196)["ms"])
