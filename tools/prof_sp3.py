"""A/B of the superpixel pooling forward at BASELINE configs[2]: tensor-core kernel (default) vs the SIMT scatter-reduce
(CRW_SEGMEAN_SIMT=1, set before the library is first used).  Prints one JSON line; run once per setting."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    from sapienza_video_contrastive_b200 import ops
    ops.check_device(dev)
    r = bench.superpixel_bench(dev)
    r["path"] = "simt" if os.environ.get("CRW_SEGMEAN_SIMT") else "tensor-core"
    # per-kernel times of one plain and one dilated forward (kineto)
    from torch.profiler import ProfilerActivity, profile
    B, T, C, SPn, size = bench.SP["B"], bench.SP["T"], bench.SP["C"], bench.SP["SP"], bench.SP["size"]
    g = torch.Generator(device=dev).manual_seed(0)
    lab = torch.randint(0, SPn, (B, T, size // 16, size // 16), generator=g, device=dev).repeat_interleave(16, -1).repeat_interleave(16, -2)
    maps = torch.randn(B, C, T, 32, 32, generator=g, device=dev)
    for _ in range(3):
        ops.segment_mean(maps, lab, SPn); ops.segment_mean_dilated(maps, lab, SPn, 51, "L1")
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            ops.segment_mean(maps, lab, SPn)
            ops.segment_mean_dilated(maps, lab, SPn, 51, "L1")
        torch.cuda.synchronize()
    r["kernels_us"] = {e.key[:60]: round(e.device_time_total / e.count, 1) for e in prof.key_averages() if e.device_time_total > 0}
    print(json.dumps(r))


if __name__ == "__main__":
    main()
