"""BASELINE configs[2] and configs[4]: superpixel-graph step and the (clip_len, nodes) sweep of the walk fwd+bwd, one GPU.
Writes one JSON object per line to stdout (copied into profiles/ by hand).  CUDA events, 5 warm-ups, median of 20."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sapienza_video_contrastive_b200 import ops  # noqa: E402
from tests.golden import cases  # noqa: E402

dev = torch.device("cuda", 0)


def timeit(fn, warm=5, reps=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def walk_point(B, N, T, D=128):
    f = torch.randn(B, N, T, D, device=dev, requires_grad=True)
    ones = torch.ones(1, device=dev)

    def step():
        f.grad = None
        q, loss, xent, acc = ops.walk(f, 0.07, 0.1, rng="philox")
        loss.backward(ones)

    ms = timeit(step)
    flops = 3 * (2 * (T - 1) * N * N * D + 2 * N ** 3 * 3 * (T - 2)) * B            # fwd + 2x bwd, minimal chain count
    return {"kind": "walk_fwd_bwd", "B": B, "N": N, "T": T, "D": D, "ms": ms, "clips_per_s": B / ms * 1e3,
            "algorithmic_tflops": flops / ms / 1e9,
            "path": "fused" if N <= 64 and T <= 4 else ("tcgen05 tf32" if 64 <= N < 512 and N % 4 == 0 else ("tcgen05 f16 split" if N >= 512 else "simt"))}


def superpixel_point(B=8, T=8, SP=196, C=512):
    g = torch.Generator().manual_seed(7)
    lab = cases.voronoi_labels(B, T, SP, 256, g, one_based=False).to(dev)
    lab3 = lab[:, :, None].repeat(1, 1, 3, 1, 1)
    maps = torch.randn(B, C, T, 32, 32, device=dev, requires_grad=True)
    torch.manual_seed(0)
    w = torch.nn.Linear(C, 128, bias=False).to(dev).weight
    ones = torch.ones(1, device=dev)
    parts = {}

    def step():
        maps.grad = None
        w.grad = None
        pooled = ops.segment_mean(maps, lab3[:, :, 0], SP)
        f = ops.head_linear(pooled, w)
        q, loss, xent, acc = ops.walk(f, 0.07, 0.1, rng="philox")
        loss.backward(ones)

    ms = timeit(step)
    parts["segmean_fwd_ms"] = timeit(lambda: ops.segment_mean(maps.detach(), lab3[:, :, 0], SP))
    bytes_alg = maps.numel() * 4 + lab.numel() * 8
    return {"kind": "superpixel_step", "B": B, "T": T, "SP": SP, "C": C, "ms": ms, "clips_per_s": B / ms * 1e3, **parts,
            "segmean_fwd_gbs": bytes_alg / parts["segmean_fwd_ms"] / 1e6}


def main():
    print(json.dumps(superpixel_point()), flush=True)
    for SP in (100, 256):
        print(json.dumps(superpixel_point(SP=SP)), flush=True)
    for T in (4, 8, 16):
        for N in (49, 64, 100, 128, 256, 512, 1024):
            B = max(1, min(64, int(2e9 / (N ** 3 * T))))
            if N >= 512 and T == 16:
                B = 1
            try:
                print(json.dumps(walk_point(B, N, T)), flush=True)
            except Exception as e:
                print(json.dumps({"kind": "walk_fwd_bwd", "N": N, "T": T, "error": repr(e)[:200]}), flush=True)


if __name__ == "__main__":
    main()
