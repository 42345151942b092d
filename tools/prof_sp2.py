"""Profiling target: one forward + backward of the superpixel pooling (plain and dilated) at the BASELINE configs[2] shape and one
label-propagation call (2 target frames) - the launches the round-2 ncu summary still lacked."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sapienza_video_contrastive_b200 import LabelPropagator, ops  # noqa: E402

dev = torch.device("cuda", 0)
B, T, C, SPn, size = 8, 8, 512, 196, 256
g = torch.Generator(device=dev).manual_seed(0)
pts = torch.rand(B, T, SPn, 2, generator=g, device=dev) * size
yx = torch.stack(torch.meshgrid(torch.arange(size, device=dev), torch.arange(size, device=dev), indexing="ij"), -1).float()
lab = torch.stack([torch.cdist(yx.reshape(1, -1, 2).expand(T, -1, -1), pts[b]).argmin(-1) for b in range(B)]).reshape(B, T, size, size)
maps = torch.randn(B, C, T, 32, 32, generator=g, device=dev, requires_grad=True)
gout = torch.randn(B, SPn, T, C, generator=g, device=dev)
for _ in range(2):
    ops.segment_mean(maps, lab, SPn).backward(gout)
    ops.segment_mean_dilated(maps, lab, SPn, 51, "L1").backward(gout)
c = bench.LP
feats, lbls = bench.lp_inputs(False, n_tgt=2)
lp = LabelPropagator(c["n_ctx"], [0], c["radius"], c["k"], c["tau"], normalize=True)
for _ in range(2):
    lp(feats.to(dev), lbls.to(dev))
torch.cuda.synchronize()
print("ok")
