"""Randomised shapes through the host simulator (tests/cusim) against the oracle: dilated superpixel pooling (forward, and the
backward through the adjoint identity <grad, d> = <gout, pool(d)>) and key-point extraction.  A bounded sample per run; the
generators are seeded, so a failure reproduces."""
import random

import pytest
import torch

from oracle import crw_oracle as O
from sapienza_video_contrastive_b200 import _lib
from tests.cusim import build_sim

DILATE = {"L1": 0, "circle": 1, "cross": 2}


@pytest.fixture(scope="module")
def sim():
    return _lib.CrwLib(build_sim.build())


def ptr(t):
    return None if t is None else t.data_ptr()


def test_dilated_pooling_random_shapes(sim):
    rnd = random.Random(11)
    done = 0
    while done < 24:
        sy, sx = rnd.choice([1, 2, 3, 4, 8]), rnd.choice([1, 2, 3, 4, 5, 8, 16, 32])
        if sy * sx > 64:
            continue
        Hm, Wm = rnd.randint(1, 6), rnd.randint(1, 40 if sx <= 8 else 6)
        SP, T, B, C = rnd.choice([1, 2, 7, 33, 64, 65, 200, 255]), rnd.randint(1, 2), rnd.randint(1, 2), rnd.choice([1, 3, 8])
        ksize, shape, bs = rnd.choice([1, 3, 5, 9, 15, 31]), rnd.choice(list(DILATE)), rnd.choice([1, 2, 3, 7])
        h, w = Hm * sy, Wm * sx
        g = torch.Generator().manual_seed(done)
        lab = torch.randint(-1, SP + 1, (B, T, (h + bs - 1) // bs, (w + bs - 1) // bs), generator=g)           # -1 and SP: ignored labels
        lab = lab.repeat_interleave(bs, 2).repeat_interleave(bs, 3)[..., :h, :w].contiguous()
        maps = torch.randn(B, C, T, Hm, Wm, generator=g)
        wsb = sim.crw_segmean_dilated_workspace_bytes(B, T, Hm, Wm, h, w, SP)
        ws = torch.zeros(wsb, dtype=torch.uint8)
        out = torch.empty(B, SP, T, C)
        sb, st, ssy, ssx = lab.stride()
        sim.check(sim.crw_segmean_dilated_fwd(ptr(maps), ptr(lab), sb, st, ssy, ssx, B, C, T, Hm, Wm, h, w, SP, ksize, DILATE[shape],
                                              ptr(out), ptr(ws), wsb, None))
        cfg = (B, T, C, Hm, Wm, sy, sx, SP, ksize, shape, bs)
        torch.testing.assert_close(out.transpose(1, 2), O.segment_mean_dilated(maps, lab, SP, ksize, shape), rtol=1e-5, atol=2e-6, msg=str(cfg))
        gout, d = torch.randn(B, SP, T, C, generator=g), torch.randn(B, C, T, Hm, Wm, generator=g)
        gm = torch.empty_like(maps)
        sim.check(sim.crw_segmean_dilated_bwd(ptr(gout), ptr(ws), wsb, B, C, T, Hm, Wm, h, w, SP, ptr(gm), None))
        lhs = float((gm * d).sum())
        rhs = float((gout.transpose(1, 2) * O.segment_mean_dilated(d, lab, SP, ksize, shape)).sum())
        assert abs(lhs - rhs) <= 1e-3 * max(1e-6, float((gm * d).abs().sum())), cfg
        done += 1


def test_pose_coords_random_shapes(sim):
    rnd = random.Random(3)
    for it in range(30):
        h, w, L, k, n = rnd.randint(1, 30), rnd.randint(1, 40), rnd.randint(2, 9), rnd.randint(1, 4), rnd.randint(1, 3)
        g = torch.Generator().manual_seed(it)
        pred = torch.rand(n, h, w, L, generator=g)
        if it % 3 == 0:
            pred = (pred * 4).floor() / 4                           # many equal values: the position tie rule
        if it % 5 == 0:
            pred[0, ..., 1] = 0                                     # an empty channel -> (-1, -1)
        coords = torch.empty(n, 2, L - 1)
        sim.check(sim.crw_lp_pose_coords(ptr(pred), n, h, w, L, k, ptr(coords), None))
        for f in range(n):
            ref = O.process_pose(pred[f], torch.zeros(L, 3).numpy(), topk=k)[0]
            assert torch.allclose(coords[f], ref, rtol=0, atol=0, equal_nan=True), (it, n, h, w, L, k)
