"""Sweep of the stream-staggered step (pipeline.PatchWalkPipeline): micro-batches x SMs given to the pooling kernels.
Also the pooling kernels alone as a function of the SM limit.  CUDA-event medians over >= 50 ms."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sapienza_video_contrastive_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
ops.set_async_wgrad(True)
L = _lib.lib()
c = bench.CFG
rows, hw = c["B"] * c["N"] * c["T"] * c["Ce"], 64
maps = torch.randn(rows, hw, device=dev)
pooled = torch.empty(rows, device=dev)
st = torch.cuda.current_stream().cuda_stream
for sms in (0,):
    f, _, _ = bench.timed_median(lambda: L.crw_pool_patch_fwd_sm(maps.data_ptr(), pooled.data_ptr(), rows, hw, sms, st))
    b, _, _ = bench.timed_median(lambda: L.crw_pool_patch_bwd_sm(pooled.data_ptr(), maps.data_ptr(), rows, hw, sms, st))
    gb = (rows * hw * 4 + rows * 4) / 1e6
    print(json.dumps({"pool_sms": sms, "fwd_us": f * 1e3, "fwd_gbs": gb / f, "bwd_us": b * 1e3, "bwd_gbs": gb / b}), flush=True)
del maps, pooled
torch.cuda.empty_cache()
configs = [([4] * 5, 100, 4, 100), ([4] * 5, 100, 4, 116), ([4] * 5, 100, 4, 132), ([4] * 5, 100, 4, 0), ([4] * 5, 108, 4, 124), ([5] * 4, 100, 4, 124), ([4] * 5, 92, 4, 124)]
for sizes, sms, hs, sb in configs:
    res = []
    for trial in range(4):
        hp = bench.HotPath(dev, 0, use_graph=True, parts=1, pool_sms=sms, sizes=sizes, head_splits=hs, pool_sms_bwd=sb)
        hp.prepare()
        ms, n, tot = bench.timed_median(hp.step, min_ms=30.0)
        res.append(round(ms, 4))
        del hp
        torch.cuda.empty_cache()
    print(json.dumps({"sizes": sizes, "pool_sms": sms, "head_splits": hs, "pool_sms_bwd": sb, "ms_per_step_trials": res}), flush=True)
