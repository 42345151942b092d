"""One small invocation of every hot kernel family, for compute-sanitizer (memcheck / racecheck / synccheck): patch pooling
(plain and SM-limited), head GEMMs (gemm_tf32: tcgen05 + TMA + mbarrier pipeline), the small-graph walk (pair kernels, 4-CTA
cluster chain with DSMEM exchange), the batched walk on the tf32 tensor-core GEMM, superpixel pooling (TMA-pipelined accumulate),
label propagation (tcgen05 pre-ranking + fp32-faithful pass + exact re-ranking).  Shapes are tiny: the tools slow kernels 10-100x."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sapienza_video_contrastive_b200 import LabelPropagator, ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
which = sys.argv[1:] or ["pool", "head", "walk", "walk_tf32", "segmean", "lp"]
if "pool" in which:
    m = torch.randn(98, 4, 64, 8, 8, device=dev, requires_grad=True)
    ops.pool_patch(m).sum().backward()
    m2 = torch.randn(98, 4, 64, 8, 8, device=dev, requires_grad=True)
    ops.pool_patch(m2, sm_limit=16).sum().backward()
    print("pool ok")
if "head" in which:
    w = torch.randn(128, 512, device=dev, requires_grad=True)
    x = torch.randn(392, 512, device=dev, requires_grad=True)
    ops.head_linear(x, w).square().sum().backward()
    print("head ok", int(ops.tc_error_word(dev)[0]))
if "walk" in which:
    f = torch.randn(2, 49, 4, 128, device=dev, requires_grad=True)
    q, loss, xent, acc = ops.walk(f, 0.07, 0.1, rng="philox")
    loss.sum().backward()
    f2 = torch.randn(2, 49, 4, 128, device=dev, requires_grad=True)
    ops.walk(f2, 0.07, 0.1, rng="philox", no_cluster=True)[1].sum().backward()
    print("walk ok", float(loss))
if "walk_tf32" in which:
    f = torch.randn(1, 128, 3, 128, device=dev, requires_grad=True)
    q, loss, xent, acc = ops.walk(f, 0.07, 0.1, rng="philox")
    loss.sum().backward()
    print("walk tf32 ok", float(loss))
if "segmean" in which:
    g = torch.Generator().manual_seed(1)
    lab = torch.randint(0, 16, (1, 2, 32, 32), generator=g).repeat_interleave(8, 2).repeat_interleave(8, 3).to(dev)
    maps = torch.randn(1, 64, 2, 32, 32, device=dev, requires_grad=True)
    ops.segment_mean(maps, lab, 16).sum().backward()
    print("segmean ok")
if "lp" in which:
    C, h, w_, n_ctx, n_tgt = 64, 16, 24, 2, 2
    feats = torch.nn.functional.normalize(torch.randn(1, C, n_ctx + n_tgt, h, w_), dim=1)
    feats[:, :, 1] = feats[:, :, 0]                      # exact ties: the fp32-faithful pass runs too
    lbls = torch.zeros(n_ctx + n_tgt, h, w_, 3)
    lbls[: n_ctx + 1, :, :, 0] = 1
    lp = LabelPropagator(n_ctx, [0], 4, 10, 0.07, normalize=False)
    preds, _ = lp(feats.to(dev), lbls)
    print("lp ok", lp.stats)
torch.cuda.synchronize()
print("sanitize smoke done")
