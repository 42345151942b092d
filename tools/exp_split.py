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
configs = [(None, 0), ([20], 0), ([10, 10], 108), ([4, 4, 4, 4, 4], 116), ([5, 5, 5, 5], 116), ([5, 5, 4, 3, 3], 112), ([6, 5, 4, 3, 2], 112), ([13, 7], 108), ([14, 6], 116), ([8, 8, 4], 108), ([8, 8, 4], 116), ([6, 6, 6, 2], 108),
           ([6, 6, 4, 4], 116), ([5, 5, 5, 5], 108), ([4, 4, 4, 4, 4], 108), ([6, 6, 4, 2, 2], 116), ([4, 4, 4, 4, 2, 2], 120), ([8, 8, 4], 0),
           ([8, 8, 4], 128), ([8, 8, 4], 96)]
for sizes, sms in configs:
    try:
        hp = bench.HotPath(dev, 0, use_graph=True, parts=1, pool_sms=sms, sizes=sizes)
        hp.prepare()
        ms, n, tot = bench.timed_median(hp.step)
        print(json.dumps({"sizes": sizes, "pool_sms": sms, "ms_per_step": ms, "clips_per_s": c["B"] / ms * 1e3, "steps": n,
                          "loss": (hp.step(), hp.loss_value())[1]}), flush=True)
        del hp
        torch.cuda.empty_cache()
    except Exception as e:
        print(json.dumps({"sizes": sizes, "pool_sms": sms, "error": repr(e)[:300]}), flush=True)
