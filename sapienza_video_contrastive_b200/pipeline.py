"""Stream-staggered execution of the patch-walk hot path (model.py:92-123 + 366-413: pool -> head -> walk fwd+bwd -> head bwd ->
pool bwd) over micro-batches of clips.

Why: the step has two HBM-bound phases (the pooling forward and backward stream 0.5 GB each) around a latency-bound middle
(head GEMMs + the three walk kernels: 60-80 CTAs on 148 SMs, ~1 % of HBM).  Run back to back, each phase leaves the other
resource idle (0.57 of the step's HBM roofline although the pooling kernels themselves sit on the roof).
Here the clips of a step are split into `n_parts` micro-batches on their own CUDA streams, staggered by one pooling pass:

    stream 0:  pool_fwd(0)  head+walk(0)               head_bwd(0)  pool_bwd(0)
    stream 1:               pool_fwd(1)  head+walk(1)               head_bwd(1)  pool_bwd(1)

so that the pooling of one micro-batch runs beside the walk of the other.  The pooling kernels are launched on a LIMITED number
of SMs (one 1024-thread CTA per SM, `pool_sms` of them - still enough bytes in flight for the HBM roof) because a walk CTA needs
a whole SM's register file: the remaining SMs stay free for the walk kernels of the other micro-batch.

Mathematically the step is unchanged (the loss is a mean over clips and the micro-batches are equal: loss = mean of the parts'
losses; gradients add).  Each micro-batch has its own device-side Philox state {seed, offset} for the edge dropout (the
kernel advances it, so a captured CUDA graph draws fresh masks on every replay).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops


class PatchWalkPipeline:
    def __init__(self, head_weight: torch.Tensor, clips: int, nodes: int, frames: int, temperature: float, dropout: float,
                 n_parts: int = 2, pool_sms: int = 0, seed: int = 123, device=None):
        if clips % n_parts:
            raise ValueError("clips (%d) must divide evenly into %d micro-batches" % (clips, n_parts))
        self.w = head_weight
        self.B, self.N, self.T, self.tau, self.p = clips, nodes, frames, float(temperature), float(dropout)
        self.n_parts, self.pool_sms = int(n_parts), int(pool_sms)
        dev = torch.device(device if device is not None else head_weight.device)
        self.dev = dev
        # one generator state per micro-batch: same seed, disjoint offset ranges far apart
        self.rng_states = [torch.tensor([seed, i << 40], dtype=torch.int64, device=dev) for i in range(self.n_parts)]
        self.streams: List[Optional[torch.cuda.Stream]] = [None] + [torch.cuda.Stream(device=dev) for _ in range(self.n_parts - 1)]
        self.scale = torch.full((1,), 1.0 / self.n_parts, device=dev)

    def step(self, parts: Sequence[torch.Tensor]):
        """parts: n_parts tensors (clips/n_parts * nodes, T, C, H, W) (the encoder's physical layout), each requiring grad.
        -> (loss [1], [d loss / d part_i], d loss / d head_weight).  All work is enqueued; the caller's stream has joined every
        side stream on return, so the results may be consumed on it (or the whole call captured in a CUDA graph)."""
        cur = torch.cuda.current_stream(self.dev)
        start = torch.cuda.Event()
        start.record(cur)
        b = self.B // self.n_parts
        D = self.w.shape[0]
        prev_pool = None
        losses, gmaps, gws = [], [], []
        for i, m in enumerate(parts):
            s = self.streams[i] or cur
            if s is not cur:
                s.wait_event(start)
            if prev_pool is not None:
                s.wait_event(prev_pool)                          # stagger: this pooling pass starts when the previous one is done
            with torch.cuda.stream(s):
                pooled = ops.pool_patch(m, sm_limit=self.pool_sms)
                prev_pool = torch.cuda.Event()
                prev_pool.record(s)
                f = ops.head_linear(pooled, self.w).view(b, self.N, self.T, D)
                q, loss, xent, acc = ops.walk(f, self.tau, self.p, rng="device", rng_state=self.rng_states[i])
                gm, gw = torch.autograd.grad(loss, [m, self.w], grad_outputs=self.scale)
                ops.join_side_streams()
                losses.append(loss)
                gmaps.append(gm)
                gws.append(gw)
        for s in self.streams[1:]:
            cur.wait_stream(s)
        loss = losses[0] if self.n_parts == 1 else torch.stack(losses).sum(0) * self.scale
        gw = gws[0] if self.n_parts == 1 else torch.stack(gws).sum(0)
        return loss, gmaps, gw
