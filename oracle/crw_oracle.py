"""CPU oracle for the CRW walk + label-propagation hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in closed form and in plain fp32 torch ops, the arithmetic of the
reference hot path (paths relative to /root/reference/):

    code/model.py:92-123    pixels_to_nodes      -> patch_nodes()
    code/model.py:260-332   image_to_nodes       -> segment_mean() / superpixel_nodes()
    code/model.py:63-72     affinity             -> affinity()
    code/model.py:74-90     stoch_mat            -> stoch_rows() (+ utils/__init__.py:414-422 ZeroSoftmax)
    code/model.py:366-415   walk + loss          -> walk_loss()
    code/utils/__init__.py:377-391 + code/test.py:118-122   radius mask -> radius_mask_additive()
    code/utils/test_utils.py:129-145  context_index_bank     -> context_index_bank()
    code/utils/test_utils.py:148-179  mem_efficient_batched_affinity -> lp_topk()
    code/test.py:141-160    label gather loop    -> lp_propagate()

It is NOT a copy of the reference: the reference builds the walk with in-place aliasing,
prefix recomputation and a python loop of bmm; here the same function is written out as
formulas.  Parity is PINNED: tests/test_oracle_golden.py checks
every function here against outputs of the unmodified reference executed in the authoring
container (oracle/gen_golden.py -> tests/golden/*.pt).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does (tests/test_host.py greps for it).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn.functional as F

EPS_LOG = 1e-20          # model.py:12
EPS_ZS = 1e-5            # utils/__init__.py:418
NEG_DROP = -1e20         # model.py:81
NEG_MASK = -1e10         # test.py:121


# ----------------------------------------------------------------------------------------------
# node formation
# ----------------------------------------------------------------------------------------------
def patch_pool(maps: torch.Tensor) -> torch.Tensor:
    """maps (BN, C, T, H, W) -> (BN, C, T) spatial mean.  model.py:116."""
    H, W = maps.shape[-2:]
    return maps.sum(-1).sum(-1) / (H * W)


def head_normalize(pooled: torch.Tensor, head_w: torch.Tensor) -> torch.Tensor:
    """pooled (R, C, T), head_w (D, C) -> unit-norm node vectors (R, T, D).  model.py:117-118
    (Linear without bias, then F.normalize over the channel dim, eps 1e-12)."""
    f = pooled.transpose(-1, -2) @ head_w.t()                    # (R, T, D)
    return f / f.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def patch_nodes(maps: torch.Tensor, head_w: torch.Tensor, B: int) -> torch.Tensor:
    """maps (B*N, C, T, H, W) -> q (B, D, T, N) as returned by pixels_to_nodes (model.py:120)."""
    BN, C, T = maps.shape[:3]
    N = BN // B
    f = head_normalize(patch_pool(maps), head_w)                 # (BN, T, D)
    return f.view(B, N, T, -1).permute(0, 3, 2, 1)


def segment_mean(maps: torch.Tensor, labels: torch.Tensor, SP: int) -> torch.Tensor:
    """Superpixel pooling stated as a segment mean (model.py:296-325).

    maps (B, C, T, Hm, Wm) fp32; labels (B, T, h, w) integer with h = s*Hm, w = s*Wm.
    out[b,t,s,:] = mean over pixels (y,x) with labels[b,t,y,x]==s of maps[b,:,t,y//s,x//s];
    labels outside [0,SP) are ignored; empty segments give zeros.  Returns (B, T, SP, C).
    """
    B, C, T, Hm, Wm = maps.shape
    h, w = labels.shape[-2:]
    sy, sx = h // Hm, w // Wm
    out = torch.zeros(B, T, SP, C, dtype=torch.float64)
    cnt = torch.zeros(B, T, SP, dtype=torch.float64)
    cell = (torch.arange(h)[:, None] // sy) * Wm + (torch.arange(w)[None, :] // sx)   # (h, w)
    cell = cell.reshape(-1)
    for b in range(B):
        for t in range(T):
            lab = labels[b, t].reshape(-1).long()
            ok = (lab >= 0) & (lab < SP)
            feats = maps[b, :, t].reshape(C, -1).t().double()            # (Hm*Wm, C)
            out[b, t].index_add_(0, lab[ok], feats[cell[ok]])
            cnt[b, t].index_add_(0, lab[ok], torch.ones(int(ok.sum()), dtype=torch.float64))
    out = out / (cnt[..., None] + EPS_LOG)
    return out.float()


def superpixel_nodes(maps, labels, SP, head_w) -> torch.Tensor:
    """-> sp_feats (B, D, T, SP).  model.py:328-330."""
    f = segment_mean(maps, labels, SP) @ head_w.t()              # (B, T, SP, D)
    f = f / f.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    return f.permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------------------------
# walk
# ----------------------------------------------------------------------------------------------
def affinity(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """model.py:63-72.  (B,C,T,N),(B,C,T,M) -> (B,T,N,M); 3-D inputs -> (B,N,M)."""
    if x1.ndim < 4:
        return torch.einsum("bcn,bcm->bnm", x1, x2)
    return torch.einsum("bctn,bctm->btnm", x1, x2)


def zero_softmax(x: torch.Tensor) -> torch.Tensor:
    """utils/__init__.py:414-422 along the last dim."""
    e = (torch.exp(x) - 1.0) ** 2
    return e / (e.sum(-1, keepdim=True) + EPS_ZS)


def stoch_rows(A: torch.Tensor, drop: Optional[torch.Tensor], tau: float, softmax: bool = False):
    """model.py:74-90 without the in-place side effect: rows of A -> transition rows."""
    if drop is not None:
        A = A.masked_fill(drop, NEG_DROP)
    x = A / tau
    return F.softmax(x, dim=-1) if softmax else zero_softmax(x)


def draw_uniforms(B: int, N: int, T: int, device="cpu", generator=None):
    """The reference's 2(T-1) rand_like draws in its order: first T-1 tensors for the
    forward matrices, then T-1 for the backward ones whose *physical* layout is the transpose of
    their logical one.  Returns (u12, u21p), each (T-1, B, N, N); u21p is physical, i.e. the
    backward matrix element (row m, col n) of pair i uses u21p[i, b, n, m]."""
    mk = lambda: torch.rand(B, N, N, device=device, generator=generator)
    u12 = torch.stack([mk() for _ in range(T - 1)]) if T > 1 else torch.empty(0, B, N, N)
    u21 = torch.stack([mk() for _ in range(T - 1)]) if T > 1 else torch.empty(0, B, N, N)
    return u12, u21


def walk_matrices(q, tau, p, u12, u21p, softmax=False):
    """q (B, D, T, N) unit-norm.  Returns lists A12[i], A21[i] (B,N,N), i = 0..T-2."""
    B, D, T, N = q.shape
    A = affinity(q[:, :, :-1], q[:, :, 1:])                      # (B, T-1, N, N)
    A12, A21 = [], []
    for i in range(T - 1):
        m1 = (u12[i] < p) if p > 0 else None
        A12.append(stoch_rows(A[:, i], m1, tau, softmax))
    for i in range(T - 1):
        # union mask: the forward drop was written in place into the shared buffer
        m = ((u12[i] < p) | (u21p[i] < p)).transpose(-1, -2) if p > 0 else None
        A21.append(stoch_rows(A[:, i].transpose(-1, -2), m, tau, softmax))
    return A12, A21


def walk_chain_products(q, tau, p, u12=None, u21p=None, flip=False, softmax=False):
    """model.py:376-382 / teacherstudent.py:507-515: the palindrome products `aar` (`aal` with flip) of walks 1..T-2."""
    B, D, T, N = q.shape
    A12, A21 = walk_matrices(q, tau, p, u12, u21p, softmax)
    out = []
    for i in range(1, T - 1):                                    # shortest cycle skipped (F7)
        chain = A12[: i + 1] + A21[: i + 1][::-1]
        if flip:
            chain = chain[::-1]
        W = chain[0]
        for M in chain[1:]:
            W = W @ M
        out.append(W)
    return out


def teacher_student_loss(q_s, q_t, tau, p, alpha, us12=None, us21p=None, ut12=None, ut21p=None, flip=False, softmax=True):
    """teacherstudent.py:494-577 (CRWTeacherStudent.forward from the node embeddings on).  q_s, q_t (B,D,T,N) unit-norm student /
    teacher nodes.  Edge dropout is applied to the TEACHER as well (stoch_mat's do_dropout defaults to True at :527-528,
    whatever the comment above those lines says), with its own draws made after the student's.
    -> (loss [1], walk xents list, teacher-student xents list)."""
    B, D, T, N = q_s.shape
    Ws = walk_chain_products(q_s, tau, p, us12, us21p, flip, softmax)
    Wt = walk_chain_products(q_t, tau, p, ut12, ut21p, flip, softmax)
    xents, ts = [], []
    for A, At in zip(Ws, Wt):
        diag = torch.diagonal(A, dim1=-2, dim2=-1)
        xents.append((-(torch.log(diag + EPS_LOG) - torch.log((A + EPS_LOG).sum(-1)))).mean())
        ts.append((-(At * torch.log_softmax(A, dim=-1)).sum(-1)).mean())          # SoftCrossEntropyLoss, :270-292
    zero = torch.zeros(1, device=q_s.device)
    loss = alpha * (sum(xents, zero) / max(1, len(xents))) + (1 - alpha) * (sum(ts, zero) / max(1, len(ts)))
    return loss, xents, ts


def walk_loss(q, tau, p, u12=None, u21p=None, flip=False, softmax=False):
    """model.py:366-413.  Returns (loss[1], xents list, accs list, names list)."""
    B, D, T, N = q.shape
    xents, accs, names = [], [], []
    tgt = torch.arange(N, device=q.device)
    for i, W in enumerate(walk_chain_products(q, tau, p, u12, u21p, flip, softmax), start=1):
        diag = torch.diagonal(W, dim1=-2, dim2=-1)
        rows = (W + EPS_LOG).sum(-1)
        xents.append((-(torch.log(diag + EPS_LOG) - torch.log(rows))).mean())
        accs.append((W.argmax(-1) == tgt[None]).float().mean())
        names.append(("l%d" if flip else "r%d") % i)
    loss = sum(xents, torch.zeros(1, device=q.device)) / max(1, len(xents))
    return loss, xents, accs, names


# ----------------------------------------------------------------------------------------------
# label propagation
# ----------------------------------------------------------------------------------------------
def context_index_bank(n_context: int, long_mem: Sequence[int], N: int) -> torch.Tensor:
    """test_utils.py:129-145, concatenated as in test.py:114 -> (N, len(long_mem)+n_context)."""
    cols = []
    for t in long_mem:
        assert 0 <= t < N, "context frame out of bounds"
        col = torch.full((N, 1), 0, dtype=torch.long)
        if t > 0:
            col += t + n_context + 1
            col[: n_context + t + 1] = 0
        cols.append(col)
    rows = torch.arange(N)[:, None]
    cols.append(rows + torch.arange(n_context)[None, :])
    return torch.cat(cols, dim=-1)


def radius_in(h: int, w: int, radius: float) -> torch.Tensor:
    """Boolean (h*w, h*w): True where float32 sqrt(dy^2+dx^2) < radius.  utils/__init__.py:381-385."""
    gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    gy, gx = gy.reshape(-1), gx.reshape(-1)
    d = ((gy[:, None] - gy[None, :]) ** 2 + (gx[:, None] - gx[None, :]) ** 2).float() ** 0.5
    return d < radius


def radius_mask_additive(h: int, w: int, radius: float) -> torch.Tensor:
    """(1,1,hw,hw) additive mask: 0 inside the radius, -1e10 outside.  test.py:118-122."""
    inside = radius_in(h, w, radius)
    return torch.where(inside, 0.0, NEG_MASK)[None, None].float()


def lp_topk(feats, key_indices, n_context, n_long, h, w, radius, tau, k):
    """test_utils.py:148-179 restated per target frame.

    feats (C, Nf, hw) unit-norm.  key_indices (Nt, n_long+n_context).  Returns Ws (Nt,k,hw) fp32,
    Is (Nt,k,hw) int64 with flat index slot*hw + key_pos, sorted by descending score.
    """
    C, Nf, hw = feats.shape
    Nt, S = key_indices.shape
    add = radius_mask_additive(h, w, radius)[0, 0] if radius is not None else None   # (key, query)
    Ws, Is = [], []
    for n in range(Nt):
        qf = feats[:, n + n_context]                                     # (C, hw)
        sc = []
        for c in range(S):
            kf = feats[:, key_indices[n, c]]                             # (C, hw)
            a = kf.t() @ qf                                              # (key, query)
            if c >= n_long and add is not None:
                a = a + add
            sc.append(a)
        sc = torch.cat(sc, 0) / tau                                      # (S*hw, hw)
        v, i = torch.topk(sc, k, dim=0)
        Ws.append(F.softmax(v, dim=0))
        Is.append(i)
    return torch.stack(Ws), torch.stack(Is)


def lp_propagate(lbls, key_indices, Ws, Is, n_context, norm_mask=False):
    """test.py:141-160.  lbls (Nf, h, w, L) soft labels (frames >= n_context are overwritten).
    Returns the soft maps handed to dump_predictions (Nt, h, w, L); frame 0 of the targets keeps lbls[0].
    norm_mask (test.py:162-164): every returned map is min-max-normalised over L; for t = 0 `pred` is a view of lbls[0], so
    the GROUND-TRUTH frame 0 is normalised in place AFTER it was copied to lbls[n_context], and later frames propagate from
    the normalised frame 0 (the propagated frames themselves are stored un-normalised)."""
    lbls = lbls.clone()
    lbls[n_context:] *= 0
    Nf, h, w, L = lbls.shape
    preds = []
    for t in range(key_indices.shape[0]):
        ctx = lbls[key_indices[t]].reshape(-1, L).t()                    # (L, S*hw)
        pred = (ctx[:, Is[t]] * Ws[t][None]).sum(1)                      # (L, hw)
        pred = pred.view(L, h, w).permute(1, 2, 0)
        if t == 0:
            pred = lbls[0]                                               # a view (test.py:159)
        lbls[t + n_context] = pred
        if norm_mask:
            pred[:, :, :] -= pred.min(-1)[0][:, :, None]
            pred[:, :, :] /= pred.max(-1)[0][:, :, None]
        preds.append(pred.clone())
    return torch.stack(preds)


def upsample_argmax(pred: torch.Tensor, lbl_set: torch.Tensor, size, norm_mask: bool = False):
    """Label-map post-processing: test.py:162-164 (--norm_mask) then utils/test_utils.py:96-103
    (dump_predictions): cv2.resize of the (h,w,L) soft label map to the image size (bilinear), numpy arg-max over L, colour
    table look-up.  -> (cls (H,W) int64, pred_lbl (H,W,3) int32, pred_dist (H,W,L) float32).  Needs OpenCV (test infra only)."""
    import cv2
    import numpy as np
    p = pred.clone().float()
    if norm_mask:
        p -= p.min(-1)[0][:, :, None]
        p /= p.max(-1)[0][:, :, None]
    H, W = int(size[0]), int(size[1])
    dist = cv2.resize(p.numpy(), (W, H))
    if dist.ndim == 2:
        dist = dist[..., None]
    cls = np.argmax(dist, axis=-1)
    lbl = np.array(lbl_set.cpu(), dtype=np.int32)[cls]
    return torch.from_numpy(cls), torch.from_numpy(lbl), torch.from_numpy(dist)


def dilation_halfwidths(ksize: int, shape: str):
    """Half-width w(dy) of the structuring element of utils/__init__.py:590-608 at vertical offset dy = -R..R (R = ksize//2):
    the element contains (dy, dx) iff |dx| <= w(dy).  'L1': diamond, 'circle': disc, 'cross': one row and one column."""
    R = ksize // 2
    out = []
    for dy in range(-R, R + 1):
        if shape == "L1":
            out.append(R - abs(dy))
        elif shape == "circle":
            w = 0
            while (w + 1) ** 2 + dy * dy <= R * R:
                w += 1
            out.append(w)
        elif shape == "cross":
            out.append(R if dy == 0 else 0)
        else:
            raise ValueError(shape)
    return out


def segment_mean_dilated(maps: torch.Tensor, labels: torch.Tensor, SP: int, ksize: int, shape: str = "L1") -> torch.Tensor:
    """Superpixel pooling with dilated masks (model.py:303-309 + utils/__init__.py:590-608): every label's
    binary mask is dilated by the structuring element (a pixel belongs to label s iff some pixel of s lies inside the element
    centred on it - the reference's depthwise conv2d(...) > 0), masks may overlap, then the same window counts / size
    normalisation / weighted sum as the plain path (model.py:311-325).  maps (B,C,T,Hm,Wm), labels (B,T,h,w) -> (B,T,SP,C)."""
    B, C, T, Hm, Wm = maps.shape
    h, w = labels.shape[-2:]
    sy, sx = h // Hm, w // Wm
    R = ksize // 2
    hw_ = dilation_halfwidths(ksize, shape)
    out = torch.zeros(B, T, SP, C, dtype=torch.float64)
    for b in range(B):
        for t in range(T):
            lab = labels[b, t].long()
            onehot = torch.zeros(SP, h, w, dtype=torch.bool)
            ok = (lab >= 0) & (lab < SP)
            ys, xs = torch.nonzero(ok, as_tuple=True)
            onehot[lab[ys, xs], ys, xs] = True
            dil = torch.zeros_like(onehot)
            for i, dy in enumerate(range(-R, R + 1)):
                if abs(dy) >= h:
                    continue
                # rows shifted by dy, dilated horizontally by hw_[i]
                src = torch.zeros_like(onehot)
                if dy >= 0:
                    src[:, : h - dy if dy else h] = onehot[:, dy:]
                else:
                    src[:, -dy:] = onehot[:, : h + dy]
                wv = hw_[i]
                if wv > 0:
                    src = torch.nn.functional.max_pool2d(src[None].float(), (1, 2 * wv + 1), stride=1, padding=(0, wv))[0] > 0
                dil |= src
            cnt = dil.view(SP, Hm, sy, Wm, sx).sum((2, 4)).double()                       # (SP, Hm, Wm) window counts
            size = dil.sum((1, 2)).double()
            wgt = cnt / (size + EPS_LOG)[:, None, None]
            out[b, t] = torch.einsum("sij,cij->sc", wgt, maps[b, :, t].double())
    return out.float()


def process_pose(pred: torch.Tensor, lbl_set, topk: int = 3):
    """utils/test_utils.py:60-84 restated: pred (h,w,L) soft maps, channel 0 = background -> (coords (2,L-1) float32 with
    -1 for channels that are zero everywhere, sharp (h,w,3) float64 image with lbl_set[c] at each key point).  Ties between
    equal values rank by position here (torch.topk leaves their order unspecified)."""
    import numpy as np
    h, w, L = pred.shape
    flat = pred[..., 1:].reshape(h * w, L - 1).float()
    k = min(h * w, topk)
    order = torch.argsort(flat, dim=0, descending=True, stable=True)[:k]              # (k, L-1), smaller position first on ties
    vals = torch.gather(flat, 0, order)
    vals = vals / vals.sum(0)[None]
    xx, yy = order % w, order // w
    coords = torch.stack([(xx * vals).sum(0), (yy * vals).sum(0)], dim=0)
    coords[:, flat.sum(0) == 0] = -1
    sharp = np.zeros((h, w, 3))
    lbl = np.asarray(lbl_set)
    for t in range(L - 1):
        x, y = int(coords[0, t]), int(coords[1, t])
        if x >= 0 and y >= 0:
            sharp[y, x, :] = lbl[t + 1]
    return coords, sharp


def sinkhorn_knopp(A: torch.Tensor, tol: float = 0.01, max_iter: int = 1000):
    """utils/__init__.py:615-641 restated (3-D input): -> (A2, number of sweeps)."""
    A = A / A.sum(-1).sum(-1)[:, None, None]
    A2 = A
    it = 0
    while (A2.sum(-2).std() > tol and it < max_iter) or it == 0:
        A1 = F.normalize(A2, p=1, dim=-2)
        A2 = F.normalize(A1, p=1, dim=-1)
        it += 1
    return A2, it


# ----------------------------------------------------------------------------------------------
# patch-grid producer: code/utils/augs.py:59-82
# ----------------------------------------------------------------------------------------------
def _pil_bilinear_coeffs(in_size: int, out_size: int):
    """Pillow 12.2 src/libImaging/Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR (triangle) filter:
    per output index the first tap and the fixed-point taps (PRECISION_BITS = 22).  Third-party arithmetic (Pillow is a
    dependency of the reference, requirements.txt), restated from its published algorithm and pinned against PIL itself."""
    import numpy as np
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    bounds, taps = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = []
        for x in range(xmax):
            v = abs((x + xmin - center + 0.5) / filterscale)
            w.append(1.0 - v if v < 1.0 else 0.0)
        ww = sum(w)
        w = [c / ww if ww != 0 else c for c in w]
        taps.append([int(0.5 + c * (1 << 22)) if c >= 0 else int(-0.5 + c * (1 << 22)) for c in w])
        bounds.append(xmin)
    return bounds, taps


def patch_grid(frame, boxes, win=64, stride=32, out_size=64, mean=(0.4914, 0.4822, 0.4465), std=(0.2023, 0.1994, 0.2010)):
    """augs.py:59-82 for one frame: frame (H, W, 3) uint8, boxes (P, 4) {top, left, height, width} per window (view_as_windows
    order) -> (P * 3, out_size, out_size) fp32 = cat over windows of Normalize(ToTensor(resize(crop(window)))).  Integer
    arithmetic of Pillow's two-pass 8-bit resampling, so the result is bit-exact."""
    import numpy as np
    frame = np.asarray(frame, dtype=np.uint8)
    H, W = frame.shape[:2]
    nwx, nwy = (W - win) // stride + 1, (H - win) // stride + 1
    outs = []
    for p in range(nwy * nwx):
        wy, wx = (p // nwx) * stride, (p % nwx) * stride
        i, j, h, w = (int(v) for v in boxes[p])
        crop = frame[wy + i: wy + i + h, wx + j: wx + j + w].astype(np.int64)
        bx, kx = _pil_bilinear_coeffs(w, out_size)
        tmp = np.empty((h, out_size, 3), np.int64)
        for xx in range(out_size):
            acc = np.full((h, 3), 1 << 21, np.int64)
            for t, c in enumerate(kx[xx]):
                acc += crop[:, bx[xx] + t] * c
            tmp[:, xx] = np.clip(acc >> 22, 0, 255)
        by, ky = _pil_bilinear_coeffs(h, out_size)
        res = np.empty((out_size, out_size, 3), np.int64)
        for yy in range(out_size):
            acc = np.full((out_size, 3), 1 << 21, np.int64)
            for t, c in enumerate(ky[yy]):
                acc += tmp[by[yy] + t] * c
            res[yy] = np.clip(acc >> 22, 0, 255)
        t = torch.from_numpy(res.astype(np.uint8)).permute(2, 0, 1).float().div(255)          # ToTensor
        t = (t - torch.tensor(mean)[:, None, None]) / torch.tensor(std)[:, None, None]          # Normalize
        outs.append(t)
    return torch.cat(outs, 0)
