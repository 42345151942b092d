"""Drivers that run the UNMODIFIED reference (oracle/ref_import.py: /root/reference/code, or the byte-identical files under the
git-ignored oracle/_ref/code) on the benchmark workloads.  BENCH / TEST INFRASTRUCTURE: used by bench.py's `--impl reference`
arm and its cpu_baseline / torch_gpu_baseline legs only - never by the product package.

  * walk_step(...)   code/model.py:334-415  CRW.forward + backward, the encoder swapped for precomputed maps (SURVEY 8d
                     "Timing method": only a1-a7 are timed), on any device;
  * lp_video(...)    code/test.py:67-160    the evaluator loop (feature normalisation aside): key bank, radius mask,
                     mem_efficient_batched_affinity (code/utils/test_utils.py:148-179) and the sequential label gather.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import tempfile
import time

import torch
import torch.nn as nn

from . import ref_import


class FakeEnc(nn.Module):
    """Returns prepared feature maps as a leaf (the reference's `encoder` slot).  Owns a parameter (model.py:42) and answers
    infer_dims' 256-px probe with a 32x32 map so that map_scale == 8 (model.py:40-45)."""

    def __init__(self, ce):
        super().__init__()
        self.dummy = nn.Parameter(torch.zeros(1))
        self.ce = ce
        self.maps = None

    def forward(self, x):
        if self.maps is None:
            return torch.zeros(1, self.ce, 1, 32, 32, device=self.dummy.device)
        return self.maps


def available() -> bool:
    return ref_import.available()


def make_walk_module(ce, head_weight, device, dropout, temp, flip=False):
    """The reference's CRW(args) on `device` with the FakeEnc encoder and the given head weights."""
    ref_model, ref_utils, _ = ref_import.load()
    enc = FakeEnc(ce)
    orig = ref_utils.make_encoder
    ref_utils.make_encoder = lambda a: enc
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            crw = ref_model.CRW(ref_import.namespace(device=str(device), dropout=dropout, temp=temp, flip=flip))
    finally:
        ref_utils.make_encoder = orig
    crw = crw.to(device)
    with torch.no_grad():
        crw.selfsim_fc[0].weight.copy_(head_weight)
    return crw, enc


def walk_step(maps, head_weight, B, N, T, tau, p, device):
    """-> step(i): one reference forward + backward on maps (B*N, Ce, T, H, W) (a leaf on `device`); returns the loss tensor.
    The dropout draws are the reference's own rand_like calls under torch.manual_seed(123 + i)."""
    crw, enc = make_walk_module(maps.shape[1], head_weight, device, p, tau)
    enc.maps = maps
    x = torch.zeros(B, T, N * 3, 8, 8, device=device)          # only its shape is read (model.py:336-349); FakeEnc ignores it

    def step(i):
        torch.manual_seed(123 + i)
        maps.grad = None
        crw.selfsim_fc[0].weight.grad = None
        q, loss, diags = crw(x, None, None)
        loss.mean().backward()
        return loss

    step.module = crw
    return step


def lp_video(feats, lbls, n_ctx, long_mem, radius, topk, tau, device, norm_mask=False):
    """Runs code/test.py's test() on one fake video whose encoder output is `feats` (1, C, Nf, h, w) (already normalised) and
    whose labels are `lbls` (Nf, h, w, L).  -> (Ws (Nt,k,hw), Is (Nt,k,hw), preds (Nt,h,w,L), seconds inside test())."""
    _, _, ref_tu = ref_import.load()
    import test as ref_test                     # the reference's test.py (ref_import put its directory first on sys.path)
    ref_test.vis = None                         # test.py:201 reads an undefined module global
    Nf = feats.shape[2]
    L = lbls.shape[-1]

    class Model:
        def encoder(self, x):                   # test.py:87, called on 5-frame chunks; imgs only carry the frame index
            b0 = int(x[0, 0, :, 0, 0].tolist()[0])
            return feats[:, :, b0:b0 + x.shape[2]]

    imgs = torch.arange(Nf).float()[None, :, None, None, None].repeat(1, 1, 3, 2, 2)
    loader = [(imgs, torch.zeros(1, Nf, 3, 4, 4), lbls[None].clone(), None, torch.zeros(1, L, 3), {})]
    captured, preds = {}, []
    orig_aff, orig_dump = ref_tu.mem_efficient_batched_affinity, ref_tu.dump_predictions

    def aff(*a, **k):
        Ws, Is = orig_aff(*a, **k)
        captured["Ws"], captured["Is"] = torch.stack(Ws), torch.stack(Is)
        return Ws, Is

    def dump(pred, *a, **k):
        preds.append(torch.from_numpy(pred).clone())
        return None, None, None

    ref_tu.mem_efficient_batched_affinity, ref_tu.dump_predictions = aff, dump
    args = argparse.Namespace(videoLen=n_ctx, long_mem=list(long_mem), radius=radius, temperature=tau, topk=topk,
                              device=str(device), no_l2=True, pca_vis=False, norm_mask=norm_mask, filelist="davis",
                              save_path=tempfile.mkdtemp(), visdom=False)
    try:
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter()
            ref_test.test(loader, Model(), args)
            if str(device).startswith("cuda"):
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
    finally:
        ref_tu.mem_efficient_batched_affinity, ref_tu.dump_predictions = orig_aff, orig_dump
    return captured["Ws"], captured["Is"], torch.stack(preds), dt
