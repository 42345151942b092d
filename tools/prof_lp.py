"""Label propagation (BASELINE configs[3] shape): per-kernel device times in situ from torch.profiler."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
bench.label_prop_bench(dev, cpu=False, gpu_baseline=False)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    r = bench.label_prop_bench(dev, cpu=False, gpu_baseline=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
print(r["value"], "frames/s")
