"""Time the superpixel label-map producer (csrc/slic.cu) on a training-shaped batch: B clips x T frames of 256 x 256, and the
same work on the host cores with the oracle's algorithm is NOT timed here (pure numpy / Python, not a fair baseline)."""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from sapienza_video_contrastive_b200 import superpixels as SP  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=80)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--segments", type=int, default=30)
    ap.add_argument("--compactness", type=float, default=200.0)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(a.frames, 3, a.size, a.size, generator=g)
    x = torch.nn.functional.avg_pool2d(x, 9, 1, 4).to(dev)        # smooth blobs
    out = {}
    for conn in (False, True):
        SP.slic_frames(x, a.segments, a.compactness, enforce_connectivity=conn)
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lab = SP.slic_frames(x, a.segments, a.compactness, enforce_connectivity=conn)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out["connectivity" if conn else "clustering_only"] = {"ms": float(np.median(ts)), "frames_per_s": a.frames / (np.median(ts) * 1e-3),
                                                              "labels_max": int(lab.max())}
    print(json.dumps({"slic": out, "frames": a.frames, "size": a.size, "segments": a.segments, "compactness": a.compactness}))


if __name__ == "__main__":
    main()
