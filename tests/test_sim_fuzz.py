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


def test_walk_properties_hypothesis(sim):
    """SURVEY 4 "Property tests" (VERDICT r1 missing #7): hypothesis over clip length, node count, feature width, dropout rate,
    flip, softmax and EMPTY nodes (all-zero feature rows: zero embedding -> all-zero ZeroSoftmax row), both the fused
    small-graph kernels and the batched path of the walk (model.py:366-413) against the oracle: loss, per-walk cross-entropies,
    accuracy up to one near-tie, and the gradient with respect to the un-normalised node vectors."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import HealthCheck, given, settings, strategies as st
    from tests.test_sim_kernels import run_walk

    @settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(B=st.integers(1, 3), N=st.integers(2, 40), T=st.integers(3, 7), D=st.sampled_from([8, 32, 128]),
           p=st.sampled_from([0.0, 0.1, 0.5]), flip=st.booleans(), softmax=st.booleans(), general=st.booleans(),
           n_empty=st.integers(0, 2), seed=st.integers(0, 10 ** 6))
    def check(B, N, T, D, p, flip, softmax, general, n_empty, seed):
        g = torch.Generator().manual_seed(seed)
        f = torch.randn(B, N, T, D, generator=g)
        for e in range(min(n_empty, N - 1)):                       # empty superpixels: zero vectors in every frame of clip 0
            f[0, e] = 0
        u12, u21p = O.draw_uniforms(B, N, T, generator=g)
        fo = f.clone().requires_grad_(True)
        qo = (fo / fo.norm(dim=-1, keepdim=True).clamp_min(1e-12)).permute(0, 3, 2, 1)
        loss_o, xents, accs, _ = O.walk_loss(qo, 0.07, p, u12, u21p, flip=flip, softmax=softmax)
        loss_o.sum().backward()
        flags = (1 if softmax else 0) | (2 if flip else 0) | (4 if general else 0)
        q, xe, ac, gr = run_walk(sim, f, 0.07, p, u12, u21p, flags)
        cfg = dict(B=B, N=N, T=T, D=D, p=p, flip=flip, softmax=softmax, general=general, n_empty=n_empty, seed=seed)
        torch.testing.assert_close(q, qo.detach().permute(0, 3, 2, 1), rtol=1e-5, atol=1e-6, msg=str(cfg))
        torch.testing.assert_close(xe, torch.stack(xents).detach(), rtol=3e-5, atol=1e-6, msg=str(cfg))
        assert float((ac - torch.stack(accs)).abs().max()) <= 1.0 / (B * N) + 1e-6, cfg
        scale = float(fo.grad.abs().max())
        assert float((gr - fo.grad).abs().max()) <= 1e-4 * scale + 1e-9, cfg

    check()
