"""ncu target: the two data-loader producers (SLIC label maps on 8 frames of 256 x 256, patch grid on 8 frames)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sapienza_video_contrastive_b200 import augs, ops, superpixels  # noqa: E402

dev = torch.device("cuda", 0)
ops.check_device(dev)
g = torch.Generator(device=dev).manual_seed(0)
vid = torch.nn.functional.avg_pool2d(torch.randn(8, 3, 256, 256, generator=g, device=dev), 9, 1, 4)
frames = torch.randint(0, 256, (8, 256, 256, 3), dtype=torch.uint8, generator=g, device=dev)
torch.manual_seed(0)
boxes = augs.draw_patch_boxes(8, 49, 64).to(dev)
for _ in range(2):
    superpixels.slic_frames(vid, 30, 200.0)
    augs.patch_grid_frames(frames, boxes, 64, 32, 64)
torch.cuda.synchronize()
print("ok")
