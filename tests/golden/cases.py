"""Deterministic inputs for the golden fixtures (shared by oracle/gen_golden.py and the tests).

Inputs are regenerated from seeds with the CPU generator (bit-stable for a given torch build; the
GPU box runs the same image), so the committed fixtures only hold the reference's OUTPUTS.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# name -> dict(B, T, N, Ce, p, tau, flip, seed)           patch-walk cases (model.py:334-415)
WALK_CASES = {
    "w_cfg1like": dict(B=2, T=4, N=49, Ce=32, p=0.1, tau=0.07, flip=False, seed=11),
    "w_b3t5":     dict(B=3, T=5, N=49, Ce=32, p=0.1, tau=0.07, flip=False, seed=12),
    "w_nodrop":   dict(B=2, T=4, N=49, Ce=32, p=0.0, tau=0.07, flip=False, seed=13),
    "w_flip":     dict(B=2, T=4, N=25, Ce=16, p=0.2, tau=0.05, flip=True,  seed=14),
    "w_t3":       dict(B=2, T=3, N=49, Ce=16, p=0.1, tau=0.07, flip=False, seed=15),
    "w_n16t8":    dict(B=2, T=8, N=16, Ce=16, p=0.3, tau=0.07, flip=False, seed=16),
    "w_n100t4":   dict(B=1, T=4, N=100, Ce=16, p=0.1, tau=0.07, flip=False, seed=17),
}

# superpixel cases (model.py:260-332): SP labels on 256x256, maps 32x32
SP_CASES = {
    "sp_16":      dict(B=1, T=3, SP=16, Ce=32, one_based=False, p=0.1, tau=0.07, seed=21),
    "sp_40_empty": dict(B=2, T=3, SP=40, Ce=16, one_based=True, p=0.1, tau=0.07, seed=22),
}

# dilated superpixel cases (model.py:303-309, utils/__init__.py:590-608): all three structuring elements, an empty node
SPD_CASES = {
    "spd_l1":     dict(B=1, T=2, SP=14, Ce=16, one_based=False, ksize=11, shape="L1", seed=51),
    "spd_circle": dict(B=1, T=2, SP=12, Ce=16, one_based=True, ksize=15, shape="circle", seed=52),
    "spd_cross":  dict(B=2, T=1, SP=20, Ce=8, one_based=False, ksize=21, shape="cross", seed=53),
}

# teacher-student walk cases (teacherstudent.py:472-580): node vectors in, softmax transition matrices
TS_CASES = {
    "ts_basic":  dict(B=2, T=4, N=25, D=32, p=0.1, tau=0.07, alpha=0.5, flip=False, seed=61),
    "ts_flip":   dict(B=1, T=5, N=16, D=16, p=0.2, tau=0.05, alpha=0.25, flip=True, seed=62),
    "ts_nodrop": dict(B=2, T=3, N=36, D=32, p=0.0, tau=0.07, alpha=0.8, flip=False, seed=63),
}

# key-point cases (utils/test_utils.py:60-84, process_pose): distinct values, an all-zero channel, fewer positions than top-k
POSE_CASES = {
    "pose_jhmdb": dict(h=40, w=53, L=16, zero=[3], seed=71),
    "pose_small": dict(h=1, w=2, L=4, zero=[], seed=72),
    "pose_wide":  dict(h=7, w=300, L=3, zero=[1], seed=73),
}

# label-propagation cases (test_utils.py:148-179, test.py:141-160)
LP_CASES = {
    "lp_small":   dict(C=16, h=12, w=17, n_ctx=4, n_tgt=6, long_mem=[0], radius=3, k=5, tau=0.07, L=3,
                       dyadic=False, repeat_first=False, seed=31),
    "lp_dyadic":  dict(C=8, h=10, w=13, n_ctx=3, n_tgt=5, long_mem=[0], radius=4, k=4, tau=0.07, L=4,
                       dyadic=True, repeat_first=False, seed=32),
    "lp_repeat":  dict(C=16, h=9, w=11, n_ctx=5, n_tgt=7, long_mem=[0], radius=12, k=10, tau=0.05, L=2,
                       dyadic=False, repeat_first=True, seed=33),
    "lp_long2":   dict(C=16, h=8, w=9, n_ctx=3, n_tgt=8, long_mem=[0, 2], radius=2, k=3, tau=0.07, L=2,
                       dyadic=False, repeat_first=False, seed=34),
}

# label-propagation cases that reach the tensor-core kernel (C a multiple of 64, k <= 16, window <= 32 keys): generic
# unit-norm features at C = 64 and C = 256, a replicated first frame (exact ties are the norm there) and a dyadic grid
LP_TC_CASES = {
    "lp_tc64":      dict(C=64, h=20, w=27, n_ctx=3, n_tgt=4, long_mem=[0], radius=5, k=10, tau=0.07, L=3,
                         dyadic=False, repeat_first=False, seed=35),
    "lp_tc256":     dict(C=256, h=18, w=40, n_ctx=4, n_tgt=3, long_mem=[0], radius=12, k=10, tau=0.05, L=4,
                         dyadic=False, repeat_first=False, seed=36),
    "lp_tc_repeat": dict(C=64, h=17, w=19, n_ctx=5, n_tgt=7, long_mem=[0], radius=12, k=10, tau=0.07, L=2,
                         dyadic=False, repeat_first=True, seed=37),
    "lp_tc_dyadic": dict(C=128, h=16, w=24, n_ctx=3, n_tgt=4, long_mem=[0], radius=4, k=5, tau=0.07, L=4,
                         dyadic=True, repeat_first=False, seed=38),
}
# test.py:158-164 with --norm_mask: soft (non-one-hot) ground truth on frame 0, so the in-place normalisation of the view
# lbls[0] at t = 0 changes what every later frame propagates from.  Distinct frames: with replicated ones the tied keys
# would carry DIFFERENT labels here (normalised frame 0 vs its un-normalised copies) and torch.topk's unspecified tie order
# would decide the output
LP_NORM_CASE = dict(C=16, h=11, w=14, n_ctx=3, n_tgt=6, long_mem=[0], radius=4, k=5, tau=0.07, L=3,
                    dyadic=False, repeat_first=False, soft_first=True, seed=39)

# patch-grid producer (utils/augs.py:59-82): frame size, seeds of numpy's (stride draw) and torch's (crop draws) generators
PG_CASE = dict(H=160, W=128, np_seed=3, torch_seed=5, frame_seed=91)


def pg_frame(c):
    g = torch.Generator().manual_seed(c["frame_seed"])
    return torch.randint(0, 256, (c["H"], c["W"], 3), generator=g, dtype=torch.uint8)


# label-map post-processing cases (utils/test_utils.py:85-123): integer and non-integer scale factors, down-scaling, norm_mask
POST_CASES = {
    "post_x8":      dict(h=12, w=17, L=3, H=96, W=136, norm_mask=False, seed=41),
    "post_davis":   dict(h=30, w=54, L=4, H=240, W=427, norm_mask=False, seed=42),
    "post_norm":    dict(h=15, w=9, L=5, H=100, W=71, norm_mask=True, seed=43),
    "post_shrink":  dict(h=40, w=33, L=2, H=17, W=20, norm_mask=False, seed=44),
}


def post_inputs(c):
    """-> pred (h,w,L) soft label map (smooth blobs + noise, rows sum to 1), lbl_set (L,3) int64 colours, img (H,W,3)."""
    g = torch.Generator().manual_seed(c["seed"])
    ys, xs = torch.meshgrid(torch.arange(c["h"]).float(), torch.arange(c["w"]).float(), indexing="ij")
    logits = torch.stack([-((ys - torch.rand(1, generator=g) * c["h"]) ** 2 + (xs - torch.rand(1, generator=g) * c["w"]) ** 2)
                          / (0.1 * c["h"] * c["w"]) for _ in range(c["L"])], -1)
    pred = torch.softmax(logits + 0.3 * torch.randn(c["h"], c["w"], c["L"], generator=g), -1)
    lbl_set = torch.randint(0, 256, (c["L"], 3), generator=g)
    img = torch.rand(c["H"], c["W"], 3, generator=g) * 255
    return pred, lbl_set, img


def walk_inputs(c):
    """-> maps (B*N, Ce, T, 8, 8), head_w (128, Ce).  The walk RNG seed is c['seed'] + 1000."""
    g = torch.Generator().manual_seed(c["seed"])
    maps = torch.randn(c["B"] * c["N"], c["Ce"], c["T"], 8, 8, generator=g)
    head_w = torch.randn(128, c["Ce"], generator=g) / c["Ce"] ** 0.5
    return maps, head_w


def ts_inputs(c):
    """-> student, teacher node vectors before normalisation, each (B, N, T, D); the teacher is a perturbed student.
    The walk RNG seed is c['seed'] + 1000."""
    g = torch.Generator().manual_seed(c["seed"])
    fs = torch.randn(c["B"], c["N"], c["T"], c["D"], generator=g)
    ft = fs + 0.5 * torch.randn(c["B"], c["N"], c["T"], c["D"], generator=g)
    return fs, ft


def pose_inputs(c):
    """-> pred (h, w, L) non-negative soft maps with distinct values per channel, lbl_set (L, 3)."""
    g = torch.Generator().manual_seed(c["seed"])
    hw = c["h"] * c["w"]
    pred = torch.stack([torch.randperm(hw, generator=g).float() + 1 for _ in range(c["L"])], -1).reshape(c["h"], c["w"], c["L"])
    pred = pred / pred.sum((0, 1), keepdim=True) * torch.rand(c["L"], generator=g)
    for z in c["zero"]:
        pred[..., z] = 0
    lbl_set = torch.randint(0, 256, (c["L"], 3), generator=g)
    return pred, lbl_set


def voronoi_labels(B, T, SP, size, g, one_based):
    ys, xs = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
    out = torch.empty(B, T, size, size, dtype=torch.long)
    for b in range(B):
        for t in range(T):
            n_real = SP - 1 if one_based else SP
            pts = torch.rand(n_real, 2, generator=g) * size
            d = (ys[None] - pts[:, 0, None, None]) ** 2 + (xs[None] - pts[:, 1, None, None]) ** 2
            out[b, t] = d.argmin(0) + (1 if one_based else 0)
    return out


def sp_inputs(c):
    """-> maps (B, Ce, T, 32, 32), labels3 (B, T, 3, 256, 256) int64, head_w (128, Ce)."""
    g = torch.Generator().manual_seed(c["seed"])
    maps = torch.randn(c["B"], c["Ce"], c["T"], 32, 32, generator=g)
    head_w = torch.randn(128, c["Ce"], generator=g) / c["Ce"] ** 0.5
    lab = voronoi_labels(c["B"], c["T"], c["SP"], 256, g, c["one_based"])
    return maps, lab[:, :, None].repeat(1, 1, 3, 1, 1), head_w


def lp_inputs(c):
    """-> feats (1, C, Nf, h, w) unit-norm over C, lbls (Nf, h, w, L) with frame 0 one-hot blobs."""
    g = torch.Generator().manual_seed(c["seed"])
    Nf = c["n_ctx"] + c["n_tgt"]
    if c["dyadic"]:
        # multiples of 2^-6 in [-1,1]: every dot product is exact in fp32 in any summation order
        feats = torch.randint(-64, 65, (1, c["C"], Nf, c["h"], c["w"]), generator=g).float() / 64.0
    else:
        feats = F.normalize(torch.randn(1, c["C"], Nf, c["h"], c["w"], generator=g), dim=1)
    if c["repeat_first"]:
        feats[:, :, : c["n_ctx"] + 1] = feats[:, :, :1]        # vos.py:148-149 replicates frame 0
    lbls = torch.zeros(Nf, c["h"], c["w"], c["L"])
    seg = torch.randint(0, c["L"], (c["h"], c["w"]), generator=g)
    first = F.one_hot(seg, c["L"]).float()
    if c.get("soft_first"):                                     # bilinear-resized one-hots are soft (vos.py:262)
        first = torch.softmax(2.0 * first + torch.randn(c["h"], c["w"], c["L"], generator=g), -1)
    lbls[: c["n_ctx"] + 1] = first                              # replicated first frame carries GT
    lbls[c["n_ctx"] + 1:] = torch.rand(c["n_tgt"] - 1, c["h"], c["w"], c["L"], generator=g)  # junk, zeroed by test.py:142
    return feats, lbls
