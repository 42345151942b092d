"""Times the plain and the dilated superpixel pooling (forward + backward) at the BASELINE config 3 shape.
usage: python tools/bench_dilated.py [B] [T] [C] [SP] [ksize]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sapienza_video_contrastive_b200 import ops  # noqa: E402
from tests.golden import cases  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    B, T, C, SP, ksize = [int(v) for v in sys.argv[1:6]] + [8, 8, 512, 196, 51][len(sys.argv) - 1:]
    g = torch.Generator().manual_seed(0)
    lab = cases.voronoi_labels(B, T, SP, 256, g, one_based=False).cuda()
    maps = torch.randn(B, C, T, 32, 32, device="cuda", requires_grad=True)
    gout = torch.randn(B, SP, T, C, device="cuda")

    def run(fn):
        out = fn()
        out.backward(gout)
        maps.grad = None

    print("shape B=%d T=%d C=%d SP=%d, labels 256x256, maps 32x32" % (B, T, C, SP))
    print("plain           fwd %.3f ms  fwd+bwd %.3f ms" % (timed(lambda: ops.segment_mean(maps.detach(), lab, SP)),
                                                          timed(lambda: run(lambda: ops.segment_mean(maps, lab, SP)))))
    for shape in ("L1", "circle", "cross"):
        f = timed(lambda: ops.segment_mean_dilated(maps.detach(), lab, SP, ksize, shape))
        fb = timed(lambda: run(lambda: ops.segment_mean_dilated(maps, lab, SP, ksize, shape)))
        print("dilated %-6s %d fwd %.3f ms  fwd+bwd %.3f ms" % (shape, ksize, f, fb))
    # per-kernel device times of one dilated forward + backward (torch.profiler, in situ)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            run(lambda: ops.segment_mean_dilated(maps, lab, SP, ksize, "L1"))
        torch.cuda.synchronize()
    for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
        print("  %-60s %8.1f us x %d" % (ev.key[:60], ev.device_time_total / max(ev.count, 1), ev.count))
    # what the reference does instead: one-hot masks, fp16 depthwise convolution, threshold (model.py:303-309)
    oh = (lab[:1, :, None] == torch.arange(SP, device="cuda")[None, None, :, None, None]).flatten(1, 2).half()
    k = torch.ones(T * SP, 1, ksize, ksize, device="cuda", dtype=torch.half)
    t = timed(lambda: torch.nn.functional.conv2d(oh, k, padding=ksize // 2, groups=T * SP) > 0, iters=3)
    print("library fp16 depthwise conv2d dilation alone, ONE clip: %.3f ms" % t)


if __name__ == "__main__":
    main()
