"""Where does the tensor-core path differ from the SIMT kernel on replicated frames (exact ties)?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sapienza_video_contrastive_b200 import LabelPropagator  # noqa: E402

C, h, w, n_ctx, n_tgt, k, radius = 192, 18, 21, 4, 5, 12, 4
g = torch.Generator().manual_seed(C + h + k)
feats = torch.nn.functional.normalize(torch.randn(1, C, n_ctx + n_tgt, h, w, generator=g), dim=1)
feats[:, :, : n_ctx + 1] = feats[:, :, :1]
out = {}
for mode in ("simt", "tc", "tc_exact_only"):
    lp = LabelPropagator(n_ctx, [0], radius, k, 0.07, normalize=False, force_simt=(mode == "simt"), exact_only=(mode == "tc_exact_only"))
    ki, Ws, Is = lp.affinity(feats.cuda())
    out[mode] = (Ws.cpu(), Is.cpu(), dict(lp.stats))
    print(mode, lp.stats)
hw = h * w
for mode in ("tc", "tc_exact_only"):
    d = out[mode][1] != out["simt"][1]
    print(mode, "differing picks", int(d.sum()), "of", d.numel(), "per target", d.sum((1, 2)).tolist())
    idx = d.nonzero()
    for n, r, q in idx[:6].tolist():
        a, b = out[mode][1][n, :, q], out["simt"][1][n, :, q]
        print(" target", n, "query", q, "rank", r)
        print("   tc  ", [(int(x) // hw, int(x) % hw) for x in a])
        print("   simt", [(int(x) // hw, int(x) % hw) for x in b])
        print("   Ws tc  ", [round(float(x), 5) for x in out[mode][0][n, :, q]])
        print("   Ws simt", [round(float(x), 5) for x in out["simt"][0][n, :, q]])
