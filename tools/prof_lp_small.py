"""One label-propagation call at the DAVIS shape with few target frames (one wave of CTAs): the launch to capture with
ncu --set full --import-source on -k regex:lp_topk_tc_kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sapienza_video_contrastive_b200 import LabelPropagator  # noqa: E402

n_tgt = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
c = bench.LP
feats, lbls = bench.lp_inputs(False, n_tgt=n_tgt)
lp = LabelPropagator(c["n_ctx"], [0], c["radius"], c["k"], c["tau"], normalize=True)
fd, ld = feats.to(dev), lbls.to(dev)
for _ in range(2):
    lp(fd, ld)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lp(fd, ld)
e1.record()
torch.cuda.synchronize()
print("targets", n_tgt, "ms", e0.elapsed_time(e1), lp.stats)
if "--kineto" in sys.argv:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            lp(fd, ld)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))
