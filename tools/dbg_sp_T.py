"""Does the 32 KB row stride (T = 8 frames x 4 KB) limit the tensor-core pooling's TMA stream?  Kernel time for T = 7, 8, 9."""
import sys
import torch
sys.path.insert(0, ".")
from torch.profiler import ProfilerActivity, profile
from sapienza_video_contrastive_b200 import ops
dev = torch.device("cuda:0")
ops.check_device(dev)
B, C, SPn, size = 8, 512, 196, 256
g = torch.Generator(device=dev).manual_seed(0)
for T in (7, 8, 9):
    lab = torch.randint(0, SPn, (B, T, size // 16, size // 16), generator=g, device=dev).repeat_interleave(16, -1).repeat_interleave(16, -2)
    maps = torch.randn(B, C, T, 32, 32, generator=g, device=dev)
    for _ in range(3):
        ops.segment_mean(maps, lab, SPn)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            ops.segment_mean(maps, lab, SPn)
        torch.cuda.synchronize()
    for e in prof.key_averages():
        if "mma" in e.key:
            us = e.device_time_total / e.count
            print("T=%d  %.1f us  %.2f TB/s" % (T, us, maps.numel() * 4 / us / 1e6))
