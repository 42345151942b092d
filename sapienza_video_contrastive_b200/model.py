"""Drop-in for the reference's `CRW` module (code/model.py:14-125,260-425) with the hot path on sm_100a kernels.

Same constructor (`CRW(args, vis=None)` reading the argparse Namespace fields of model.py:19-38), same parameter
names (`encoder.model.*`, `selfsim_fc.0.weight`: checkpoints interchange), same method surface and return values:

    forward(x, sp_mask=None, max_sp_num=None, just_feats=False, orig_unnorm=None) -> (q, loss[1], diags)
    affinity(x1, x2) / stoch_mat(A, zero_diagonal, do_dropout, do_sinkhorn) / pixels_to_nodes(x)
    image_to_nodes(x, sp_mask, max_sp_num) / xent_targets(A) / zeroout_diag(A)

What changed underneath: node pooling, the L2 normalisation, the frame-pair affinities, edge dropout, the
ZeroSoftmax transition matrices, the palindrome chains, the cross-entropy AND the whole backward of those run in
libcrw_b200.so (ops.py); the ResNet encoder and the `nn.Linear` head stay on stock PyTorch (cuDNN / cuBLAS).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .resnet import make_encoder

EPS = 1e-20


class ZeroSoftmax(nn.Module):
    """code/utils/__init__.py:414-422 (eps 1e-5), differentiable, same default `dim=0` as the reference; evaluated by the
    stoch_mat kernels along the requested dim."""

    def forward(self, x, dim=0, eps=1e-5):
        if eps != 1e-5:
            raise ValueError("the kernel implements the reference's eps = 1e-5")
        last = dim in (-1, x.dim() - 1)
        xt = x if last else x.transpose(dim, -1)
        if xt.dim() == 1:
            xt = xt[None]
            y = ops.stoch_mat(xt[None], 1.0)[0, 0]
        elif xt.dim() == 2:
            y = ops.stoch_mat(xt[None], 1.0)[0]
        else:
            y = ops.stoch_mat(xt, 1.0)
        return y if last else y.transpose(dim, -1)


class CRW(nn.Module):
    def __init__(self, args, vis=None):
        super().__init__()
        self.args = args
        self.edgedrop_rate = getattr(args, "dropout", 0)
        self.featdrop_rate = getattr(args, "featdrop", 0)
        self.temperature = getattr(args, "temp", getattr(args, "temperature", 0.07))

        self.encoder = make_encoder(args).to(self.args.device)
        self.infer_dims()
        self.selfsim_fc = self.make_head(depth=getattr(args, "head_depth", 0))
        self.zero_softmax = ZeroSoftmax()

        self.xent = nn.CrossEntropyLoss(reduction="none")
        self._xent_targets = dict()
        self.dropout = nn.Dropout(p=self.edgedrop_rate, inplace=False)
        self.featdrop = nn.Dropout(p=self.featdrop_rate, inplace=False)

        self.flip = getattr(args, "flip", False)
        self.sk_targets = getattr(args, "sk_targets", False)
        self.vis = vis
        # superpixel mask dilation (model.py:38, 303-309): the reference materialises a fp16 structuring element and runs a
        # depthwise convolution over the one-hot masks; here the element is only its (size, shape) and the pooling kernels
        # dilate label runs directly (ops.segment_mean_dilated)
        self.dilation = None
        if getattr(args, "dilate_superpixels", False):
            ksize = int(getattr(args, "dilation_kernel_size", 51))               # utils/arguments.py:209-210 defaults
            shape = getattr(args, "dilation_kernel_shape", "L1")
            assert ksize % 2 != 0, "Use an odd kernel size"                      # utils/__init__.py:591
            if shape not in ("L1", "circle", "cross"):
                raise ValueError("dilation_kernel_shape must be L1 | circle | cross, got %r" % (shape,))
            self.dilation = (ksize, shape)
        # 'philox': edge dropout is drawn inside the walk kernel from torch's CUDA generator stream (same numbers,
        # same generator advance as the reference's rand_like calls); 'torch': drawn by torch.rand and passed in.
        self.rng = getattr(args, "crw_rng", "auto")
        self.use_softmax = bool(getattr(args, "crw_softmax", False))

    # -- construction helpers (model.py:40-56) ----------------------------------------------------------------------
    def infer_dims(self):
        in_sz = 256
        dummy = torch.zeros(1, 3, 1, in_sz, in_sz).to(next(self.encoder.parameters()).device)
        dummy_out = self.encoder(dummy)
        self.enc_hid_dim = dummy_out.shape[1]
        self.map_scale = in_sz // dummy_out.shape[-1]

    def make_head(self, depth=1):
        head = []
        if depth >= 0:
            dims = [self.enc_hid_dim] + [self.enc_hid_dim] * depth + [128]
            for d1, d2 in zip(dims, dims[1:]):
                head += [nn.Linear(d1, d2, bias=False), nn.ReLU()]
            head = head[:-1]
        return nn.Sequential(*head)

    def _rng_mode(self, device):
        if self.rng == "auto":
            return "philox" if ops.philox_replay_ok(device) else "torch"
        return self.rng

    # -- operator surface -------------------------------------------------------------------------------------------
    def zeroout_diag(self, A, zero=0):
        mask = (torch.eye(A.shape[-1], device=A.device).unsqueeze(0).repeat(A.shape[0], 1, 1).bool() < 1).float()
        return A * mask

    def affinity(self, x1, x2):
        """model.py:63-72: (B,C,T,N),(B,C,T,M) -> (B,T,N,M); 3-D inputs get/lose a time dim."""
        in_t_dim = x1.ndim
        if in_t_dim < 4:
            x1, x2 = x1.unsqueeze(-2), x2.unsqueeze(-2)
        B, C, T, N = x1.shape
        M = x2.shape[-1]
        a = x1.permute(0, 2, 3, 1).reshape(B * T, N, C)
        b = x2.permute(0, 2, 3, 1).reshape(B * T, M, C)
        A = ops.affinity_nodes(a, b).view(B, T, N, M)
        return A.squeeze(1) if in_t_dim < 4 else A

    def stoch_mat(self, A, zero_diagonal=False, do_dropout=True, do_sinkhorn=False):
        """model.py:74-90.  Like the reference, dropout overwrites the caller's tensor (or view) with -1e20, and the result is
        differentiable with respect to `A` (dropped entries get no gradient)."""
        if zero_diagonal:
            A = self.zeroout_diag(A)
        u = None
        if do_dropout and self.edgedrop_rate > 0:
            u = torch.rand_like(A)                     # same draw, same layout rule as the reference
        if do_sinkhorn:
            # model.py:83-87: sinkhorn_knopp((A / tau).exp(), tol=0.01, max_iter=100) after the in-place edge drop.  Never
            # reached from the reference's forward; forward-only here (the result carries no gradient).
            if u is not None:
                with torch.no_grad():
                    A[u < self.edgedrop_rate] = -1e20
            flat = A.detach().reshape(-1, *A.shape[-2:]) if A.dim() != 3 else A.detach()
            out, _ = ops.sinkhorn_knopp(flat, tol=0.01, max_iter=100, exp_temperature=self.temperature)
            return out.reshape(A.shape)
        return ops.stoch_mat(A, self.temperature, self.edgedrop_rate if u is not None else 0.0, softmax=self.use_softmax,
                             uniform=u.contiguous() if u is not None else None)

    def pixels_to_nodes(self, x, featdrop=True):
        """model.py:92-123: x (B,N,C,T,h,w) -> feats (B,128,T,N) unit-norm, maps (B,N,C',T,H,W)."""
        f, maps, B, N = self._patch_nodes_prenorm(x, featdrop)
        q = ops.l2_normalize_last(f)                               # (B,N,T,D)
        return q.permute(0, 3, 2, 1), maps

    def image_to_nodes(self, x, sp_mask, max_sp_num):
        """model.py:260-332: x (B,T,c,h,w), sp_mask (B,T,c,h,w) int -> sp_feats (B,128,T,SP), maps (B,C,T,H,W)."""
        f, maps = self._superpixel_nodes_prenorm(x, sp_mask, max_sp_num)
        return ops.l2_normalize_last(f).permute(0, 3, 2, 1), maps

    def _patch_nodes_prenorm(self, x, featdrop=True):
        """`featdrop=False`: the teacher's variant (teacherstudent.py:453-455 comments the feature dropout out so that the
        teacher's targets stay deterministic)."""
        B, N, C, T, h, w = x.shape
        maps = self.encoder(x.flatten(0, 1))
        H, W = maps.shape[-2:]
        if featdrop and self.featdrop_rate > 0:
            maps = self.featdrop(maps)
        if N == 1:      # whole images: every feature-map position becomes a node (model.py:110-113)
            maps = maps.permute(0, 3, 4, 1, 2).contiguous()
            maps = maps.view(-1, *maps.shape[3:])[..., None, None]
            N, H, W = maps.shape[0] // B, 1, 1
        f = self._head(self._pool_nodes(maps))                                   # (BN, T, D), contiguous
        return f.reshape(B, N, T, f.shape[-1]), maps.view(B, N, *maps.shape[1:]), B, N

    def _head(self, pooled):
        """selfsim_fc (model.py:117).  The default single bias-free Linear goes through ops.head_linear (forward, input and
        split-K weight gradient on the fused tcgen05 tf32 GEMM; library GEMM for shapes TMA cannot address); deeper heads
        (head_depth > 0) run as the stock nn.Sequential."""
        if len(self.selfsim_fc) == 1 and pooled.is_cuda:
            return ops.head_linear(pooled, self.selfsim_fc[0].weight)
        return self.selfsim_fc(pooled)

    @staticmethod
    def _pool_nodes(maps):
        """(BN, C', T, H, W) -> (BN, T, C') spatial means (model.py:116), ready for the Linear head.  From3D leaves the
        maps physically as (BN, T, C', H, W); pooling that view directly avoids a 0.5 GB layout copy and lands the
        result in the layout the head consumes."""
        if maps.shape[-1] * maps.shape[-2] == 1:
            return maps[..., 0, 0].transpose(-1, -2)
        phys = maps.permute(0, 2, 1, 3, 4)
        if phys.is_contiguous():
            return ops.pool_patch(phys)                                          # (BN, T, C')
        return ops.pool_patch(maps).transpose(-1, -2)

    def _superpixel_nodes_prenorm(self, x, sp_mask, max_sp_num):
        B, T, c, h, w = x.shape
        maps = self.encoder(x.transpose(1, 2))                                   # (B, C', T, H, W)
        if self.featdrop_rate > 0:
            maps = self.featdrop(maps)
        labels = sp_mask[:, :, 0, :, :]                                          # strided view, no copy (model.py:298)
        if self.dilation is None:
            pooled = ops.segment_mean(maps, labels, int(max_sp_num))             # (B, SP, T, C')
        else:
            pooled = ops.segment_mean_dilated(maps, labels, int(max_sp_num), *self.dilation)
        return self._head(pooled), maps                                          # (B, SP, T, D)

    # -- forward (model.py:334-415) ----------------------------------------------------------------------------------
    def forward(self, x, sp_mask=None, max_sp_num=None, just_feats=False, orig_unnorm=None, walk_uniforms=None):
        """`walk_uniforms=(u12, u21p)` (each (T-1,B,N,N)) optionally supplies the edge-dropout draws explicitly."""
        B, T, C, H, W = x.shape
        _N = 1
        if sp_mask is None:
            _N, C = C // 3, 3
            x = x.transpose(1, 2).reshape(B, _N, C, T, H, W)
            f, mm, _, _ = self._patch_nodes_prenorm(x)
        else:
            f, mm = self._superpixel_nodes_prenorm(x, sp_mask, max_sp_num)
        if just_feats:
            q = ops.l2_normalize_last(f).permute(0, 3, 2, 1)
            if _N > 1 or sp_mask is not None:
                return q, mm
            h, w = np.ceil(np.array(x.shape[-2:]) / self.map_scale).astype(int)
            return q, q.view(*q.shape[:-1], h, w)

        qn, loss, xent, acc = ops.walk(f, self.temperature, self.edgedrop_rate, flip=self.flip, softmax=self.use_softmax,
                                       rng=self._rng_mode(f.device),
                                       u12=walk_uniforms[0] if walk_uniforms else None,
                                       u21p=walk_uniforms[1] if walk_uniforms else None)
        q = qn.permute(0, 3, 2, 1)                                               # (B, D, T, N)
        diags = dict()
        tag = "l" if self.flip else "r"
        for j in range(xent.shape[0]):
            diags["%s xent cyc %s%d" % (H, tag, j + 1)] = xent[j]
            diags["%s acc cyc %s%d" % (H, tag, j + 1)] = acc[j]
        return q, loss, diags

    def xent_targets(self, A):
        B, N = A.shape[:2]
        key = "%s:%sx%s" % (str(A.device), B, N)
        if key not in self._xent_targets:
            I = torch.arange(A.shape[-1])[None].repeat(B, 1)
            self._xent_targets[key] = I.view(-1).to(A.device)
        return self._xent_targets[key]
