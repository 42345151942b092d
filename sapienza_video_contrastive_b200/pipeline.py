"""Stream-staggered execution of the patch-walk hot path (model.py:92-123 + 366-413: pool -> head -> walk fwd+bwd -> head bwd ->
pool bwd) over micro-batches of clips.

Why: the step has two HBM-bound phases (the pooling forward and backward stream 0.5 GB each) around a latency-bound middle
(head GEMMs + the three walk kernels: 60-80 CTAs on 148 SMs, ~1 % of HBM).  Run back to back, each phase leaves the other
resource idle (0.57 of the step's HBM roofline although the pooling kernels themselves sit on the roof).
Here the clips of a step are split into micro-batches on their own CUDA streams, staggered by one pooling pass:

    stream 0:  pool_fwd(0)  head+walk(0)               head_bwd(0)  pool_bwd(0)
    stream 1:               pool_fwd(1)  head+walk(1)               head_bwd(1)  pool_bwd(1)

so that the pooling of one micro-batch runs beside the walk of another.  The pooling kernels are launched on a LIMITED number
of SMs (one 1024-thread CTA per SM, `pool_sms` of them - still enough bytes in flight for the HBM roof) because a walk CTA needs
a whole SM's register file: the remaining SMs stay free for the walk kernels of the other micro-batches.  The middle phase is a
chain of dependent launches whose length does not shrink with the batch, so the step ends one middle phase after the LAST
pooling pass: the micro-batches may be uneven (`sizes`), a small last one keeps that tail short.

Mathematically the step is unchanged: the loss is a mean over clips, so loss = sum_i (b_i / B) loss_i, and the gradients are
those of that sum (the weight b_i / B of a micro-batch is folded into its pooling backward and into the final reduction of the
head gradient).  Each micro-batch has its own device-side Philox state {seed, offset} for the edge dropout (the kernel advances
it, so a captured CUDA graph draws fresh masks on every replay).

The pipeline drives the C ABI directly (no autograd graph): every buffer is allocated once, so a step is a fixed set of launches
that can be captured in a CUDA graph.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib, ops


class PatchWalkPipeline:
    def __init__(self, head_weight: torch.Tensor, clips: int, nodes: int, frames: int, temperature: float, dropout: float,
                 n_parts: int = 2, pool_sms: int = 0, seed: int = 123, device=None, sizes: Optional[Sequence[int]] = None,
                 head_splits: int = 1, walk_flags: int = 0, pool_sms_bwd: Optional[int] = None):
        if sizes is None:
            if clips % n_parts:
                raise ValueError("clips (%d) must divide evenly into %d micro-batches (or pass `sizes`)" % (clips, n_parts))
            sizes = [clips // n_parts] * n_parts
        sizes = [int(x) for x in sizes]
        if sum(sizes) != clips or min(sizes) <= 0:
            raise ValueError("micro-batch sizes %s must be positive and add up to %d clips" % (sizes, clips))
        self.w = head_weight
        self.B, self.N, self.T, self.tau, self.p = clips, nodes, frames, float(temperature), float(dropout)
        self.sizes, self.n_parts, self.pool_sms = sizes, len(sizes), int(pool_sms)
        self.head_splits = int(head_splits)
        self.pool_sms_bwd = self.pool_sms if pool_sms_bwd is None else int(pool_sms_bwd)   # the store-bound backward wants more SMs
        self.walk_flags = int(walk_flags)               # e.g. _lib.WALK_NO_CLUSTER: one CTA per clip for the chain instead of a 4-CTA cluster
        dev = torch.device(device if device is not None else head_weight.device)
        ops.check_device(dev)
        self.dev = dev
        D, C = head_weight.shape
        self.D, self.C = D, C
        L = _lib.lib()
        f32 = dict(dtype=torch.float32, device=dev)
        nw = max(frames - 2, 0)
        self.nw = nw
        # one generator state per micro-batch: same seed, disjoint offset ranges far apart
        self.rng_states = [torch.tensor([seed, i << 40], dtype=torch.int64, device=dev) for i in range(self.n_parts)]
        self.streams: List[Optional[torch.cuda.Stream]] = [None] + [torch.cuda.Stream(device=dev) for _ in range(self.n_parts - 1)]
        self.side = torch.cuda.Stream(device=dev)                                            # weight gradients, one after the other
        self.err = ops.tc_error_word(dev)
        self.xent = torch.zeros(self.n_parts, nw + 1, **f32)                                 # [:, nw] = the micro-batch's loss
        self.acc = torch.zeros(self.n_parts, max(nw, 1), **f32)
        self.gw = torch.zeros(D, C, **f32)
        self.loss_parts = self.xent[:, nw]                                                   # view: one loss per micro-batch
        self.buf = []
        for b in sizes:
            R = b * nodes * frames
            wsb = L.crw_walk_workspace_bytes(b, nodes, frames, D, self.walk_flags)
            wgb = L.crw_head_wgrad_workspace_bytes(R, D, C)
            hfb = L.crw_head_fwd_splitk_workspace_bytes(R, D, self.head_splits)
            self.buf.append(dict(
                R=R, pooled=torch.empty(R, C, **f32), f=torch.empty(R, D, **f32), q=torch.empty(R, D, **f32),
                gf=torch.empty(R, D, **f32), gpooled=torch.empty(R, C, **f32),
                ws=torch.zeros(max(wsb, 256), dtype=torch.uint8, device=dev), wgws=torch.zeros(max(wgb, 256), dtype=torch.uint8, device=dev),
                hfws=torch.zeros(max(hfb, 256), dtype=torch.uint8, device=dev),
                thr=ops.torch_rand_threads(b * nodes * nodes, dev), gmaps=None))

    def step(self, parts: Sequence[torch.Tensor]):
        """parts: one tensor per micro-batch, (b_i * nodes, T, C, H, W) fp32 contiguous (the encoder's physical layout).
        -> (losses (n_parts,) - the step's loss is sum_i (b_i / B) losses[i], see `loss()` -, [d loss / d part_i],
        d loss / d head_weight (D, C)).  Everything is enqueued; on return the caller's stream has joined every side stream,
        so the results may be consumed on it, or the whole call captured in a CUDA graph.  The outputs are the pipeline's own
        buffers: they are overwritten by the next step."""
        L = _lib.lib()
        cur = torch.cuda.current_stream(self.dev)
        start = torch.cuda.Event()
        start.record(cur)
        w = self.w.detach()
        if not w.is_contiguous():
            w = w.contiguous()
        prev_pool = None
        sd = self.side
        for i, m in enumerate(parts):
            s = self.streams[i] or cur
            bf, b = self.buf[i], self.sizes[i]
            if tuple(m.shape[:3]) != (b * self.N, self.T, self.C) or m.dtype != torch.float32 or not m.is_contiguous():
                raise ValueError("micro-batch %d must be a contiguous fp32 (%d, %d, %d, H, W) tensor" % (i, b * self.N, self.T, self.C))
            hw = m.shape[-1] * m.shape[-2]
            if bf["gmaps"] is None or bf["gmaps"].shape != m.shape:
                bf["gmaps"] = torch.empty_like(m)
            if s is not cur:
                s.wait_event(start)
            if prev_pool is not None:
                s.wait_event(prev_pool)                          # stagger: this pooling pass starts when the previous one is done
            st = s.cuda_stream
            R = bf["R"]
            L.check(L.crw_pool_patch_fwd_sm(m.data_ptr(), bf["pooled"].data_ptr(), R * self.C, hw, self.pool_sms, st), "pool_patch_fwd")
            prev_pool = torch.cuda.Event()
            prev_pool.record(s)
            L.check(L.crw_head_fwd_splitk(bf["pooled"].data_ptr(), w.data_ptr(), bf["f"].data_ptr(), R, self.D, self.C, self.head_splits,
                                          bf["hfws"].data_ptr(), bf["hfws"].numel(), self.err.data_ptr(), st), "head_fwd")
            L.check(L.crw_walk_fwd_bwd(bf["f"].data_ptr(), b, self.N, self.T, self.D, self.tau, self.p, None, None, 0, 0, bf["thr"],
                                       self.rng_states[i].data_ptr() if self.p > 0 else None, self.walk_flags, bf["q"].data_ptr(),
                                       self.xent[i].data_ptr(), self.acc[i].data_ptr(), bf["gf"].data_ptr(), bf["ws"].data_ptr(),
                                       bf["ws"].numel(), st), "walk_fwd_bwd")
            fork = torch.cuda.Event()
            fork.record(s)
            sd.wait_event(fork)                                  # the weight gradient only needs gf and pooled: beside dgrad + pooling backward
            L.check(L.crw_head_wgrad_axpby(bf["gf"].data_ptr(), bf["pooled"].data_ptr(), self.gw.data_ptr(), R, self.D, self.C,
                                           b / self.B, 0.0 if i == 0 else 1.0, bf["wgws"].data_ptr(), bf["wgws"].numel(), sd.cuda_stream),
                    "head_wgrad")
            L.check(L.crw_head_dgrad(bf["gf"].data_ptr(), w.data_ptr(), bf["gpooled"].data_ptr(), R, self.D, self.C, self.err.data_ptr(), st), "head_dgrad")
            L.check(L.crw_pool_patch_bwd_scaled(bf["gpooled"].data_ptr(), bf["gmaps"].data_ptr(), R * self.C, hw, b / self.B, self.pool_sms_bwd, st),
                    "pool_patch_bwd")
        for s in self.streams[1:]:
            cur.wait_stream(s)
        cur.wait_stream(sd)
        return self.loss_parts, [bf["gmaps"] for bf in self.buf], self.gw

    def loss(self) -> torch.Tensor:
        """The step's loss [1] = sum_i (b_i / B) loss_i (model.py:413 over all clips), from the last step's per-micro-batch losses."""
        wts = torch.tensor([b / self.B for b in self.sizes], dtype=torch.float32, device=self.dev)
        return (self.loss_parts * wts).sum(0, keepdim=True)
