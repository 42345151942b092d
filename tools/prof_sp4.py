"""ncu target: the tensor-core superpixel pooling forward alone at BASELINE configs[2] (scan-ordered labels, then random ids)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sapienza_video_contrastive_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B, T, C, SPn, size = 8, 8, 512, 196, 256
g = torch.Generator(device=dev).manual_seed(0)
pts = torch.rand(B, T, SPn, 2, generator=g, device=dev) * size
band = (pts[..., 0] / (size / SPn ** 0.5)).floor()
order = (band * size + pts[..., 1]).argsort(-1)
pts = torch.gather(pts, 2, order[..., None].expand(-1, -1, -1, 2))
yx = torch.stack(torch.meshgrid(torch.arange(size, device=dev), torch.arange(size, device=dev), indexing="ij"), -1).float()
lab = torch.stack([torch.cdist(yx.reshape(1, -1, 2).expand(T, -1, -1), pts[b]).argmin(-1) for b in range(B)]).reshape(B, T, size, size)
maps = torch.randn(B, C, T, 32, 32, generator=g, device=dev)
for _ in range(3):
    ops.segment_mean(maps, lab, SPn)
perm = torch.randperm(SPn, generator=g, device=dev)
for _ in range(2):
    ops.segment_mean(maps, perm[lab], SPn)
torch.cuda.synchronize()
print("ok")
