"""ncu target: superpixel pooling at BASELINE configs[2] (tensor-core forward with scan-ordered labels, backward, dilated forward)
and the SLIC label-map producer on 8 frames of 256 x 256."""
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
from sapienza_video_contrastive_b200 import superpixels  # noqa: E402
maps.requires_grad_(True)
gout = torch.randn(B, SPn, T, C, generator=g, device=dev)
for _ in range(2):
    ops.segment_mean(maps, lab, SPn).backward(gout)
    ops.segment_mean_dilated(maps.detach(), lab, SPn, 51, "L1")
vid = torch.nn.functional.avg_pool2d(torch.randn(8, 3, size, size, generator=g, device=dev), 9, 1, 4)
for _ in range(2):
    superpixels.slic_frames(vid, 30, 200.0)
torch.cuda.synchronize()
print("ok")
