"""Time the patch-grid producer on 80 frames of 256 x 256 (CUDA events, median)."""
import json, sys
import torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda:0")
from sapienza_video_contrastive_b200 import ops
ops.check_device(dev)
r = bench.producers_bench(dev, cpu=False)
print(json.dumps(r["patch_grid"]))
