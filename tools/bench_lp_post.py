"""Label-map post-processing at DAVIS 480p shape: device kernel vs the reference's CPU path (cv2.resize +
numpy argmax + colour table, utils/test_utils.py:96-103) on the box's host cores.  One JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sapienza_video_contrastive_b200 import ops  # noqa: E402

dev = "cuda"
n, h, w, L, H, W = 37, 60, 107, 4, 480, 854
torch.manual_seed(0)
pred = torch.softmax(torch.randn(n, h, w, L), -1)
pal = torch.randint(0, 256, (L, 3))
pd = pred.to(dev)
for _ in range(3):
    ops.lp_upsample_argmax(pd, (H, W), pal)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    ops.lp_upsample_argmax(pd, (H, W), pal)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
import cv2  # noqa: E402
import numpy as np  # noqa: E402
pal_np = np.array(pal, dtype=np.int32)
t0 = time.perf_counter()
for i in range(n):                                           # utils/test_utils.py:96-103, the reference's CPU path
    dist = cv2.resize(pred[i].numpy(), (W, H))
    lbl = pal_np[np.argmax(dist, axis=-1)]
cpu_ms = (time.perf_counter() - t0) * 1e3
alg = n * H * W * 4 + pred.numel() * 4                      # class byte + 3 colour bytes written, low-res maps read
print(json.dumps({"kind": "lp_upsample_argmax", "frames": n, "shape": [h, w, L, H, W], "gpu_ms": ms, "frames_per_s": n / ms * 1e3,
                  "algorithmic_gbs": alg / ms / 1e6, "cpu_reference_ms": cpu_ms, "cpu_frames_per_s": n / cpu_ms * 1e3,
                  "cpu_threads": torch.get_num_threads()}))
