"""Kernel timeline of one replay of the (split) step: name, stream, start, duration - is there any overlap?"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sapienza_video_contrastive_b200 import ops  # noqa: E402

sizes, sms = [int(x) for x in sys.argv[1].split(",")], int(sys.argv[2])
sizes = sizes if len(sizes) > 1 else None
graph = len(sys.argv) < 4 or sys.argv[3] != "eager"
dev = torch.device("cuda", 0)
ops.set_async_wgrad(True)
hp = bench.HotPath(dev, 0, use_graph=graph, parts=1, pool_sms=sms, sizes=sizes)
hp.prepare()
for _ in range(3):
    hp.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    hp.step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print("%8.1f %7.1f  %s" % ((e.time_range.start - t0), (e.time_range.end - e.time_range.start), e.name[:70]))
print("total", ev[-1].time_range.end - t0)
