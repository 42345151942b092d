"""Kernel LOGIC tests without a GPU: the product's .cu sources compiled for the host by tests/cusim (every CUDA
thread a fiber) and driven through the same C ABI with CPU pointers, checked against the oracle and the golden
fixtures of the reference.  This library is test infrastructure; the product never loads it (see cuda_sim.h).
The real parity tests on sm_100a are the -m gpu tests.
"""
import os

import pytest
import torch

from oracle import crw_oracle as O
from sapienza_video_contrastive_b200 import _lib
from tests.cusim import build_sim
from tests.golden import cases

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def sim():
    return _lib.CrwLib(build_sim.build())


def ptr(t):
    return None if t is None else t.data_ptr()


def load(name):
    return torch.load(os.path.join(G, name + ".pt"), weights_only=False)


def run_walk(sim, f, tau, p, u12, u21p, flags=0, need_grad=True):
    B, N, T, D = f.shape
    wsb = sim.crw_walk_workspace_bytes(B, N, T, D, flags)
    ws = torch.zeros(wsb, dtype=torch.uint8)
    q = torch.empty_like(f)
    xe, ac = torch.zeros(max(T - 2, 0) + 1), torch.zeros(max(T - 2, 1))
    g = torch.empty_like(f) if need_grad else None
    for _ in range(2):       # twice: the workspace must be reusable without re-zeroing
        sim.check(sim.crw_walk_fwd_bwd(ptr(f), B, N, T, D, tau, p, ptr(u12) if p > 0 else None,
                                       ptr(u21p) if p > 0 else None, 0, 0, 0, None, flags, ptr(q), ptr(xe), ptr(ac), ptr(g),
                                       ptr(ws), wsb, None), "walk")
    if T >= 3:
        assert abs(float(xe[T - 2]) - float(xe[: T - 2].mean())) < 1e-6
    return q, xe[: max(T - 2, 0)], ac[: max(T - 2, 0)], g


def test_exports_and_version(sim):
    assert sim.crw_version() >= 100


@pytest.mark.parametrize("rows,hw", [(64, 64), (37, 64), (100, 32), (33, 16), (10, 49), (5, 1)])
def test_pool_patch(sim, rows, hw):
    torch.manual_seed(rows)
    x = torch.randn(rows, hw)
    out = torch.empty(rows)
    sim.check(sim.crw_pool_patch_fwd(ptr(x), ptr(out), rows, hw, None))
    torch.testing.assert_close(out, x.sum(-1) / hw, rtol=1e-6, atol=1e-6)
    g = torch.randn(rows)
    gx = torch.empty(rows, hw)
    sim.check(sim.crw_pool_patch_bwd(ptr(g), ptr(gx), rows, hw, None))
    assert torch.equal(gx, (g / hw)[:, None].expand(rows, hw))


@pytest.mark.parametrize("name", ["w_cfg1like", "w_flip", "w_t3", "w_n16t8", "w_n100t4", "w_nodrop"])
@pytest.mark.parametrize("force_general", [False, True])
def test_walk_vs_reference_golden(sim, name, force_general):
    """feature maps -> (oracle pooling + head, CPU) -> walk kernel; compared with the REFERENCE's own outputs."""
    c = cases.WALK_CASES[name]
    if force_general and name in ("w_n100t4",):
        pytest.skip("already the general path")
    fx = load(name)
    maps, head_w = cases.walk_inputs(c)
    pooled = O.patch_pool(maps)
    f = (pooled.transpose(-1, -2) @ head_w.t()).view(c["B"], c["N"], c["T"], 128).contiguous().requires_grad_(True)
    torch.manual_seed(c["seed"] + 1000)
    u12, u21p = O.draw_uniforms(c["B"], c["N"], c["T"]) if c["p"] > 0 else (None, None)
    flags = (_lib.WALK_FLIP if c["flip"] else 0) | (_lib.WALK_FORCE_GENERAL if force_general else 0)
    q, xe, ac, g = run_walk(sim, f.detach(), c["tau"], c["p"], u12, u21p, flags)
    torch.testing.assert_close(q.permute(0, 3, 2, 1), fx["q"], rtol=1e-5, atol=1e-6)
    names = [("l%d" if c["flip"] else "r%d") % i for i in range(1, c["T"] - 1)]
    for j, nm in enumerate(names):
        torch.testing.assert_close(xe[j], fx["diags"]["64 xent cyc %s" % nm], rtol=1e-5, atol=0)
        torch.testing.assert_close(ac[j], fx["diags"]["64 acc cyc %s" % nm], rtol=0, atol=1e-6)
    torch.testing.assert_close(xe.mean().reshape(1), fx["loss"], rtol=1e-5, atol=0)
    # gradient: chain the kernel's d loss / d feats through the (oracle) head and pooling
    f.backward(g)
    # grad wrt head weights of the reference = pooled^T-chain; recompute through autograd on the same graph
    pooled2 = pooled.clone()
    hw_ = head_w.clone().requires_grad_(True)
    f2 = (pooled2.transpose(-1, -2) @ hw_.t()).view(c["B"], c["N"], c["T"], 128)
    f2.backward(g)
    torch.testing.assert_close(hw_.grad, fx["grad_head"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("B,N,T,D,p,flags", [(2, 49, 4, 128, 0.1, 0), (1, 33, 4, 64, 0.3, 3), (1, 70, 4, 64, 0.1, 0),
                                             (1, 20, 6, 32, 0.2, 6), (1, 40, 3, 32, 0.1, 5), (2, 12, 2, 32, 0.1, 0)])
def test_walk_vs_oracle_grad(sim, B, N, T, D, p, flags):
    torch.manual_seed(B * 1000 + N)
    f = torch.randn(B, N, T, D)
    u12, u21p = O.draw_uniforms(B, N, T)
    fo = f.clone().requires_grad_(True)
    qo = (fo / fo.norm(dim=-1, keepdim=True).clamp_min(1e-12)).permute(0, 3, 2, 1)
    loss, xents, accs, _ = O.walk_loss(qo, 0.07, p, u12, u21p, flip=bool(flags & 2), softmax=bool(flags & 1))
    q, xe, ac, g = run_walk(sim, f, 0.07, p, u12, u21p, flags)
    torch.testing.assert_close(q, qo.detach().permute(0, 3, 2, 1), rtol=1e-5, atol=1e-6)
    if T >= 3:
        loss.sum().backward()
        torch.testing.assert_close(xe, torch.stack(xents), rtol=1e-5, atol=0)
        torch.testing.assert_close(ac, torch.stack(accs), rtol=0, atol=1e-6)
        assert ((g - fo.grad).abs().max() / fo.grad.abs().max()) < 1e-4
    q2, xe2, _, g2 = run_walk(sim, f, 0.07, p, u12, u21p, flags, need_grad=False)
    assert g2 is None and torch.equal(xe2, xe)


@pytest.mark.parametrize("name", list(cases.TS_CASES))
def test_teacher_student_walk_vs_reference_golden(sim, name):
    """crw_walk_ts_fwd_bwd: the teacher call hands out its chain products (checked against the oracle's), the student call
    blends the walk loss with the soft cross-entropy against them; loss, diags and the gradient of the student's node
    vectors against the reference's CRWTeacherStudent.forward."""
    c = cases.TS_CASES[name]
    fx = load(name)
    fs, ft = cases.ts_inputs(c)
    B, N, T, D = fs.shape
    J = T - 2
    torch.manual_seed(c["seed"] + 1000)
    us12, us21p = O.draw_uniforms(B, N, T)
    ut12, ut21p = O.draw_uniforms(B, N, T)
    flags = _lib.WALK_SOFTMAX | (_lib.WALK_FLIP if c["flip"] else 0)
    wsb = sim.crw_walk_workspace_bytes(B, N, T, D, flags | _lib.WALK_FORCE_GENERAL)
    ws = torch.zeros(wsb, dtype=torch.uint8)
    q = torch.empty_like(fs)
    xe, ac, tsx = torch.zeros(J + 1), torch.zeros(J), torch.zeros(J + 1)
    chains = torch.empty(B, J, N, N)
    p = c["p"]
    sim.check(sim.crw_walk_ts_fwd_bwd(ptr(ft), B, N, T, D, c["tau"], p, ptr(ut12) if p > 0 else None, ptr(ut21p) if p > 0 else None,
                                      0, 0, 0, None, flags, None, 1.0, ptr(chains), ptr(q), ptr(xe), None, ptr(ac), None,
                                      ptr(ws), wsb, None), "teacher")
    qt = torch.nn.functional.normalize(ft, dim=-1).permute(0, 3, 2, 1)
    Wt = torch.stack(O.walk_chain_products(qt, c["tau"], p, ut12, ut21p, flip=c["flip"], softmax=True), 1)
    torch.testing.assert_close(chains, Wt, rtol=1e-4, atol=1e-7)
    g = torch.empty_like(fs)
    for _ in range(2):                                           # the workspace is reusable as it is
        sim.check(sim.crw_walk_ts_fwd_bwd(ptr(fs), B, N, T, D, c["tau"], p, ptr(us12) if p > 0 else None, ptr(us21p) if p > 0 else None,
                                          0, 0, 0, None, flags, ptr(chains), c["alpha"], None, ptr(q), ptr(xe), ptr(tsx), ptr(ac), ptr(g),
                                          ptr(ws), wsb, None), "student")
    torch.testing.assert_close(q.permute(0, 3, 2, 1), fx["q"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(xe[J:], fx["loss"], rtol=1e-5, atol=0)
    names = [("l%d" if c["flip"] else "r%d") % i for i in range(1, T - 1)]
    for j, nm in enumerate(names):
        torch.testing.assert_close(xe[j], fx["diags"]["8 xent cyc %s" % nm], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(ac[j], fx["diags"]["8 acc cyc %s" % nm], rtol=0, atol=1e-6)
    _, _, ts_o = O.teacher_student_loss(fx["q"], qt, c["tau"], p, c["alpha"], us12, us21p, ut12, ut21p, flip=c["flip"])
    torch.testing.assert_close(tsx[:J], torch.stack(ts_o), rtol=1e-5, atol=0)
    assert abs(float(tsx[J]) - float(tsx[:J].mean())) < 1e-6
    assert float((g - fx["grad_feats"]).abs().max() / fx["grad_feats"].abs().max()) < 1e-4
    # alpha = 1 is the plain walk: same loss and gradient as crw_walk_fwd_bwd on the general path
    g1 = torch.empty_like(fs)
    sim.check(sim.crw_walk_ts_fwd_bwd(ptr(fs), B, N, T, D, c["tau"], p, ptr(us12) if p > 0 else None, ptr(us21p) if p > 0 else None,
                                      0, 0, 0, None, flags, ptr(chains), 1.0, None, ptr(q), ptr(xe), ptr(tsx), ptr(ac), ptr(g1),
                                      ptr(ws), wsb, None), "alpha 1")
    _, xe0, _, g0 = run_walk(sim, fs, c["tau"], p, us12, us21p, flags | _lib.WALK_FORCE_GENERAL)
    torch.testing.assert_close(xe[:J], xe0, rtol=1e-6, atol=0)
    torch.testing.assert_close(g1, g0, rtol=1e-5, atol=1e-9)
    # a teacher without somewhere to put its losses, or alpha outside [0, 1], is refused
    assert sim.crw_walk_ts_fwd_bwd(ptr(fs), B, N, T, D, c["tau"], 0.0, None, None, 0, 0, 0, None, flags, ptr(chains), 1.5, None, ptr(q),
                                   ptr(xe), ptr(tsx), ptr(ac), None, ptr(ws), wsb, None) != 0


@pytest.mark.parametrize("B,N,T,D,p,flags,alpha", [(2, 20, 4, 32, 0.2, 0, 0.3), (1, 70, 3, 64, 0.0, 2, 0.0), (1, 33, 5, 32, 0.1, 1, 1.0)])
def test_teacher_student_walk_vs_oracle(sim, B, N, T, D, p, flags, alpha):
    """ZeroSoftmax and softmax transition matrices, flip, the two ends of alpha, a graph beyond the fused kernels' size:
    losses and the gradient against autograd through the oracle."""
    g = torch.Generator().manual_seed(N * 10 + T)
    fs, ft = torch.randn(B, N, T, D, generator=g), torch.randn(B, N, T, D, generator=g)
    torch.manual_seed(N)
    us12, us21p = O.draw_uniforms(B, N, T)
    ut12, ut21p = O.draw_uniforms(B, N, T)
    J = T - 2
    nrm = lambda f: torch.nn.functional.normalize(f, dim=-1).permute(0, 3, 2, 1)
    fo = fs.clone().requires_grad_(True)
    loss_o, xents_o, ts_o = O.teacher_student_loss(nrm(fo), nrm(ft), 0.07, p, alpha, us12, us21p, ut12, ut21p, flip=bool(flags & 2),
                                                   softmax=bool(flags & 1))
    loss_o.sum().backward()
    wsb = sim.crw_walk_workspace_bytes(B, N, T, D, flags | _lib.WALK_FORCE_GENERAL)
    ws = torch.zeros(wsb, dtype=torch.uint8)
    q = torch.empty_like(fs)
    xe, ac, tsx = torch.zeros(J + 1), torch.zeros(J), torch.zeros(J + 1)
    chains = torch.empty(B, J, N, N)
    up = lambda u: ptr(u) if p > 0 else None
    sim.check(sim.crw_walk_ts_fwd_bwd(ptr(ft), B, N, T, D, 0.07, p, up(ut12), up(ut21p), 0, 0, 0, None, flags, None, 1.0, ptr(chains),
                                      ptr(q), ptr(xe), None, ptr(ac), None, ptr(ws), wsb, None), "teacher")
    gr = torch.empty_like(fs)
    sim.check(sim.crw_walk_ts_fwd_bwd(ptr(fs), B, N, T, D, 0.07, p, up(us12), up(us21p), 0, 0, 0, None, flags, ptr(chains), alpha, None,
                                      ptr(q), ptr(xe), ptr(tsx), ptr(ac), ptr(gr), ptr(ws), wsb, None), "student")
    torch.testing.assert_close(xe[J:], loss_o.detach(), rtol=1e-5, atol=0)
    torch.testing.assert_close(xe[:J], torch.stack(xents_o).detach(), rtol=1e-5, atol=0)
    torch.testing.assert_close(tsx[:J], torch.stack(ts_o).detach(), rtol=1e-5, atol=0)
    assert float((gr - fo.grad).abs().max() / fo.grad.abs().max()) < 1e-4


def test_walk_empty_nodes(sim):
    """all-zero node vectors (empty superpixels): zero ZeroSoftmax rows, loss row = log N, finite grads (F5)."""
    torch.manual_seed(3)
    f = torch.randn(1, 12, 4, 32)
    f[0, 0] = 0
    f[0, 5, 2] = 0
    u12, u21p = O.draw_uniforms(1, 12, 4)
    fo = f.clone().requires_grad_(True)
    qo = (fo / fo.norm(dim=-1, keepdim=True).clamp_min(1e-12)).permute(0, 3, 2, 1)
    loss, xents, accs, _ = O.walk_loss(qo, 0.07, 0.0, u12, u21p)
    for flags in (0, _lib.WALK_FORCE_GENERAL):
        q, xe, ac, g = run_walk(sim, f, 0.07, 0.0, None, None, flags)
        torch.testing.assert_close(xe, torch.stack(xents), rtol=1e-5, atol=0)
        assert torch.isfinite(g).all()


def test_affinity_and_stoch_mat(sim):
    torch.manual_seed(0)
    BT, N1, N2, D = 3, 19, 23, 40
    x1, x2 = torch.randn(BT, N1, D), torch.randn(BT, N2, D)
    out = torch.empty(BT, N1, N2)
    sim.check(sim.crw_affinity(ptr(x1), ptr(x2), BT, N1, N2, D, ptr(out), None))
    torch.testing.assert_close(out, x1 @ x2.transpose(-1, -2), rtol=1e-5, atol=1e-5)
    A = torch.randn(2, 7, 9)
    u = torch.rand(2, 7, 9)
    A0 = A.clone()
    y = torch.empty_like(A)
    sim.check(sim.crw_stoch_mat(ptr(A), ptr(u), 0.3, 0.07, 0, 2, 7, 9, ptr(y), None))
    assert torch.equal(A, A0.masked_fill(u < 0.3, -1e20))                 # in-place side effect (F4)
    torch.testing.assert_close(y, O.stoch_rows(A0, u < 0.3, 0.07), rtol=1e-5, atol=1e-8)
    y2 = torch.empty_like(A)
    A1 = A0.clone()
    sim.check(sim.crw_stoch_mat(ptr(A1), None, 0.0, 0.07, _lib.WALK_SOFTMAX, 2, 7, 9, ptr(y2), None))
    torch.testing.assert_close(y2, torch.softmax(A0 / 0.07, -1), rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("R,N,M,tol,max_iter,use_exp", [(3, 12, 12, 0.01, 100, True), (2, 49, 49, 1e-4, 1000, True), (1, 7, 20, 0.01, 3, False),
                                                        (4, 33, 33, 0.0, 5, True)])
def test_sinkhorn_knopp(sim, R, N, M, tol, max_iter, use_exp):
    """crw_sinkhorn_knopp against the reference's function (utils/__init__.py:615-641; imported when the reference tree is
    present, else its restatement in the oracle): same number of sweeps, same matrix."""
    import ctypes
    from oracle import ref_import
    g = torch.Generator().manual_seed(R * 100 + N)
    A = torch.randn(R, N, M, generator=g) * 0.3
    if use_exp:
        A[0, 1, 2] = -1e20                                       # a dropped edge (model.py:81): exp -> 0
    x = (A / 0.07).exp() if use_exp else A.abs() + 0.1
    ref, its = O.sinkhorn_knopp(x, tol, max_iter)
    if ref_import.available():
        _, ref_utils, _ = ref_import.load()
        torch.testing.assert_close(ref, ref_utils.sinkhorn_knopp(x, tol=tol, max_iter=max_iter), rtol=0, atol=0)
    work = A.clone() if use_exp else x.clone()
    wsb = sim.crw_sinkhorn_workspace_bytes(R, N, M)
    ws = torch.zeros(wsb, dtype=torch.uint8)
    n_it = ctypes.c_int(0)
    sim.check(sim.crw_sinkhorn_knopp(ptr(work), R, N, M, int(use_exp), 0.07, tol, max_iter, ctypes.addressof(n_it), ptr(ws), wsb, None))
    assert n_it.value == its
    torch.testing.assert_close(work, ref, rtol=2e-5, atol=1e-9)
    torch.testing.assert_close(work.sum(-1), torch.ones(R, N), rtol=1e-5, atol=1e-6)
    assert sim.crw_sinkhorn_knopp(ptr(work), R, N, M, 1, 0.0, tol, max_iter, None, ptr(ws), wsb, None) != 0


def test_philox_known_answers(sim):
    """Philox4x32-10 against the Random123 known-answer vectors, through the uniform mapping: element e of a draw with
    `threads` >= n uses counter (offset/4, 0, e, 0), component 0."""
    import ctypes
    n = 8
    out = torch.empty(n)
    sim.check(sim.crw_philox_uniform(ptr(out), n, 0, 0, 256, None))
    # counter = (0,0,0,0), key = (0,0) -> first word 0x6627e8d5
    expect0 = float(torch.tensor(0x6627e8d5, dtype=torch.float64) * 2.3283064e-10 + 1.1641532e-10)
    assert abs(out[0].item() - expect0) < 1e-7
    assert (out >= 0).all() and (out < 1).all() and out.unique().numel() == n
    # component selection: element e + threads*c reads component c of the same counter
    out2 = torch.empty(256 * 4)
    sim.check(sim.crw_philox_uniform(ptr(out2), 256 * 4, 0, 0, 256, None))
    assert out2[0] == out[0]
    expect1 = float(torch.tensor(0xe169c58d, dtype=torch.float64) * 2.3283064e-10 + 1.1641532e-10)
    assert abs(out2[256].item() - expect1) < 1e-7


@pytest.mark.parametrize("name", list(cases.SP_CASES))
def test_segmean_vs_reference_golden(sim, name):
    c = cases.SP_CASES[name]
    fx = load(name)
    maps, lab3, head_w = cases.sp_inputs(c)
    B, Ce, T = maps.shape[:3]
    SP = c["SP"]
    wsb = sim.crw_segmean_workspace_bytes(B, T, 32, 32, 256, 256, SP)
    ws = torch.zeros(wsb, dtype=torch.uint8)
    out = torch.empty(B, SP, T, Ce)
    lab = lab3[:, :, 0]                                         # strided view of channel 0, no copy
    assert not lab.is_contiguous()
    sb, st, sy, sx = lab.stride()
    sim.check(sim.crw_segmean_fwd(ptr(maps), ptr(lab), sb, st, sy, sx, B, Ce, T, 32, 32, 256, 256, SP, ptr(out), ptr(ws), wsb, None))
    torch.testing.assert_close(out.transpose(1, 2), O.segment_mean(maps, lab, SP), rtol=1e-5, atol=1e-6)
    f = out @ head_w.t()                                        # (B, SP, T, 128)
    q = (f / f.norm(dim=-1, keepdim=True).clamp_min(1e-12)).permute(0, 3, 2, 1)
    torch.testing.assert_close(q, fx["sp_feats"], rtol=1e-4, atol=2e-6)
    # backward against autograd through a differentiable restatement
    g = torch.randn(B, SP, T, Ce)
    gm = torch.empty_like(maps)
    sim.check(sim.crw_segmean_bwd(ptr(g), ptr(ws), wsb, B, Ce, T, 32, 32, 256, 256, SP, ptr(gm), None))
    m2 = maps.clone().requires_grad_(True)
    up = m2.repeat_interleave(8, -1).repeat_interleave(8, -2)                     # (B,C,T,256,256)
    oh = torch.nn.functional.one_hot(lab.clamp(0, SP - 1), SP).float() * ((lab >= 0) & (lab < SP))[..., None]
    ref = torch.einsum("bcthw,bthws->btsc", up, oh) / (oh.sum((2, 3))[..., None] + 1e-20)
    ref.backward(g.transpose(1, 2))
    torch.testing.assert_close(gm, m2.grad, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("B,T,C,Hm,scale,SP", [(1, 2, 64, 16, 4, 20), (2, 1, 128, 32, 2, 70), (1, 1, 64, 16, 8, 300)])
def test_segmean_bulk_copy_path_vs_oracle(sim, B, T, C, Hm, scale, SP):
    """Shapes the TMA kernel takes (cells % 256 == 0, C % 64 == 0): several chunks, several tiles, several items per CTA,
    labels out of range, more labels than cells; forward against the oracle, backward against autograd."""
    g = torch.Generator().manual_seed(B * 100 + SP)
    maps = torch.randn(B, C, T, Hm, Hm, generator=g)
    lab = cases.voronoi_labels(B, T, SP, Hm * scale, g, one_based=False)
    lab[:, :, :3, :5] = SP + 2                                  # ignored labels
    lab[:, :, -2:, :] = -1
    wsb = sim.crw_segmean_workspace_bytes(B, T, Hm, Hm, Hm * scale, Hm * scale, SP)
    ws = torch.zeros(wsb, dtype=torch.uint8)
    out = torch.empty(B, SP, T, C)
    sb, st, sy, sx = lab.stride()
    sim.check(sim.crw_segmean_fwd(ptr(maps), ptr(lab), sb, st, sy, sx, B, C, T, Hm, Hm, Hm * scale, Hm * scale, SP, ptr(out), ptr(ws), wsb, None))
    torch.testing.assert_close(out.transpose(1, 2), O.segment_mean(maps, lab, SP), rtol=1e-5, atol=1e-6)
    gout = torch.randn(B, SP, T, C, generator=g)
    gm = torch.empty_like(maps)
    sim.check(sim.crw_segmean_bwd(ptr(gout), ptr(ws), wsb, B, C, T, Hm, Hm, Hm * scale, Hm * scale, SP, ptr(gm), None))
    m2 = maps.clone().requires_grad_(True)
    up = m2.repeat_interleave(scale, -1).repeat_interleave(scale, -2)
    ok = (lab >= 0) & (lab < SP)
    oh = torch.nn.functional.one_hot(lab.clamp(0, SP - 1), SP).float() * ok[..., None]
    ref = torch.einsum("bcthw,bthws->btsc", up, oh) / (oh.sum((2, 3))[..., None] + 1e-20)
    ref.backward(gout.transpose(1, 2))
    torch.testing.assert_close(gm, m2.grad, rtol=1e-4, atol=1e-6)


DILATE = {"L1": 0, "circle": 1, "cross": 2}


@pytest.mark.parametrize("name", list(cases.SPD_CASES))
def test_segmean_dilated_vs_reference_golden(sim, name):
    """--dilate-superpixels (model.py:303-309): the run-dilating kernels against the reference's conv2d-dilated pooling,
    forward through the head to the golden node embeddings, backward to the golden gradient of the maps."""
    c = cases.SPD_CASES[name]
    fx = load(name)
    maps, lab3, head_w = cases.sp_inputs(c)
    B, Ce, T = maps.shape[:3]
    SP = c["SP"]
    wsb = sim.crw_segmean_dilated_workspace_bytes(B, T, 32, 32, 256, 256, SP)
    assert wsb > sim.crw_segmean_workspace_bytes(B, T, 32, 32, 256, 256, SP)
    ws = torch.zeros(wsb, dtype=torch.uint8)
    out = torch.empty(B, SP, T, Ce)
    lab = lab3[:, :, 0]
    sb, st, sy, sx = lab.stride()
    sim.check(sim.crw_segmean_dilated_fwd(ptr(maps), ptr(lab), sb, st, sy, sx, B, Ce, T, 32, 32, 256, 256, SP, c["ksize"],
                                          DILATE[c["shape"]], ptr(out), ptr(ws), wsb, None))
    torch.testing.assert_close(out.transpose(1, 2), O.segment_mean_dilated(maps, lab, SP, c["ksize"], c["shape"]), rtol=1e-5, atol=1e-6)
    pooled = out.clone().requires_grad_(True)
    f = pooled @ head_w.t()
    q = (f / f.norm(dim=-1, keepdim=True).clamp_min(1e-12)).permute(0, 3, 2, 1)
    torch.testing.assert_close(q.detach(), fx["sp_feats"], rtol=1e-4, atol=2e-6)
    g = torch.Generator().manual_seed(c["seed"] + 7)
    proj = torch.randn(q.shape, generator=g)                    # the functional gen_spd differentiated
    (q * proj).sum().backward()
    gm = torch.empty_like(maps)
    sim.check(sim.crw_segmean_dilated_bwd(ptr(pooled.grad.contiguous()), ptr(ws), wsb, B, Ce, T, 32, 32, 256, 256, SP, ptr(gm), None))
    torch.testing.assert_close(gm, fx["grad_maps"], rtol=1e-4, atol=1e-6)


def test_segmean_dilated_edges(sim):
    """ksize 1 is the undilated pooling; out-of-range labels dilate nothing; rectangular cells and a non-square image;
    arguments the kernels cannot take are refused."""
    g = torch.Generator().manual_seed(77)
    B, T, C, Hm, Wm, sy, sx, SP = 1, 2, 8, 6, 10, 4, 2, 9
    h, w = Hm * sy, Wm * sx
    maps = torch.randn(B, C, T, Hm, Wm, generator=g)
    lab = torch.randint(0, SP, (B, T, h // 3, w // 4 + 1), generator=g).repeat_interleave(3, 2).repeat_interleave(4, 3)[..., :h, :w].contiguous()
    lab[:, :, :2, :3] = SP + 1
    lab[:, :, -1, :] = -5
    sb, st, ss_y, ss_x = lab.stride()
    wsb = sim.crw_segmean_dilated_workspace_bytes(B, T, Hm, Wm, h, w, SP)
    for ksize, shape in [(1, "L1"), (5, "L1"), (7, "circle"), (9, "cross"), (2 * h + 1, "circle")]:
        ws = torch.zeros(wsb, dtype=torch.uint8)
        out = torch.empty(B, SP, T, C)
        sim.check(sim.crw_segmean_dilated_fwd(ptr(maps), ptr(lab), sb, st, ss_y, ss_x, B, C, T, Hm, Wm, h, w, SP, ksize, DILATE[shape],
                                              ptr(out), ptr(ws), wsb, None))
        ref = O.segment_mean(maps, lab, SP) if ksize == 1 else O.segment_mean_dilated(maps, lab, SP, ksize, shape)
        torch.testing.assert_close(out.transpose(1, 2), ref, rtol=1e-5, atol=1e-6)
    out = torch.empty(B, SP, T, C)
    ws = torch.zeros(wsb, dtype=torch.uint8)
    assert sim.crw_segmean_dilated_fwd(ptr(maps), ptr(lab), sb, st, ss_y, ss_x, B, C, T, Hm, Wm, h, w, SP, 4, 0, ptr(out), ptr(ws), wsb, None) != 0
    assert sim.crw_segmean_dilated_fwd(ptr(maps), ptr(lab), sb, st, ss_y, ss_x, B, C, T, Hm, Wm, h, w, SP, 5, 3, ptr(out), ptr(ws), wsb, None) != 0
    assert sim.crw_segmean_dilated_fwd(ptr(maps), ptr(lab), sb, st, ss_y, ss_x, B, C, T, Hm, Wm, h, w, SP, 129, 0, ptr(out), ptr(ws), wsb, None) != 0
    assert sim.crw_segmean_dilated_workspace_bytes(B, T, Hm, Wm, h, w, 256) == 0


@pytest.mark.parametrize("Hm,Wm,sy,sx,SP,ksize,shape", [(3, 40, 2, 8, 30, 13, "L1"),      # two strips of 32 cells
                                                         (4, 14, 3, 5, 11, 9, "circle"),    # cells straddling bitmap words
                                                         (2, 70, 1, 4, 40, 31, "cross"),    # 280 columns, one pixel row per cell
                                                         (5, 3, 2, 32, 7, 21, "L1"),        # 32-pixel-wide cells
                                                         (16, 16, 2, 2, 70, 15, "circle"),  # > 1024 CSR entries in a chunk (TMA path)
                                                         (16, 20, 2, 2, 40, 13, "L1"),      # the same on the cp.async path
                                                         (1, 100, 64, 1, 255, 127, "circle"), # tallest cells, most labels, widest element: bitmap capped by shared memory
                                                         (2, 3, 16, 4, 200, 9, "cross")])    # 16 x 4 cells
def test_segmean_dilated_strips(sim, Hm, Wm, sy, sx, SP, ksize, shape):
    g = torch.Generator().manual_seed(Hm * 1000 + Wm)
    B, T, C = 1, 2, (64 if Hm == 16 else 4)
    h, w = Hm * sy, Wm * sx
    maps = torch.randn(B, C, T, Hm, Wm, generator=g)
    lab = torch.randint(0, SP, (B, T, (h + 1) // 2, (w + 4) // 5), generator=g).repeat_interleave(2, 2).repeat_interleave(5, 3)[..., :h, :w].contiguous()
    sb, st, ss_y, ss_x = lab.stride()
    wsb = sim.crw_segmean_dilated_workspace_bytes(B, T, Hm, Wm, h, w, SP)
    ws = torch.zeros(wsb, dtype=torch.uint8)
    out = torch.empty(B, SP, T, C)
    sim.check(sim.crw_segmean_dilated_fwd(ptr(maps), ptr(lab), sb, st, ss_y, ss_x, B, C, T, Hm, Wm, h, w, SP, ksize, DILATE[shape],
                                          ptr(out), ptr(ws), wsb, None))
    torch.testing.assert_close(out.transpose(1, 2), O.segment_mean_dilated(maps, lab, SP, ksize, shape), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", list(cases.POSE_CASES))
def test_pose_coords_vs_reference_golden(sim, name):
    """crw_lp_pose_coords against the reference's process_pose (utils/test_utils.py:60-84): bit-equal coordinates, several
    frames in one launch, top-k 1..4, and the documented tie rule (equal values rank by position)."""
    c = cases.POSE_CASES[name]
    fx = load(name)
    pred, lbl_set = cases.pose_inputs(c)
    h, w, L = pred.shape
    batch = torch.stack([pred, pred.flip(0), pred * 0.5]).contiguous()
    coords = torch.full((3, 2, L - 1), 7.0)
    sim.check(sim.crw_lp_pose_coords(ptr(batch), 3, h, w, L, 3, ptr(coords), None))
    assert torch.equal(coords[0], fx["coords"])
    assert torch.equal(coords[1], O.process_pose(pred.flip(0), lbl_set.numpy())[0])
    assert torch.equal(coords[2], O.process_pose(pred * 0.5, lbl_set.numpy())[0])
    for k in (1, 2, 4):
        sim.check(sim.crw_lp_pose_coords(ptr(batch), 1, h, w, L, k, ptr(coords), None))
        assert torch.equal(coords[0], O.process_pose(pred, lbl_set.numpy(), topk=k)[0]), k
    tied = torch.zeros(h, w, L)
    tied[..., 1:] = 0.25
    sim.check(sim.crw_lp_pose_coords(ptr(tied), 1, h, w, L, 3, ptr(coords), None))
    assert torch.equal(coords[0], O.process_pose(tied, lbl_set.numpy())[0])
    assert sim.crw_lp_pose_coords(ptr(batch), 1, h, w, L, 5, ptr(coords), None) != 0 or h * w < 5


def check_label_images(cls, rgb, pred, lbl_set, c, fx):
    """Class map / label image of the kernel against the reference's dump_predictions output.  OpenCV's vectorised resize
    may fuse or reorder the two interpolation passes, so a pixel may legitimately differ only where the two best classes of
    the upsampled distribution are within float noise of each other (a documented exact-tie class, as for top-k)."""
    cls_o, lbl_o, dist = O.upsample_argmax(pred, lbl_set, (c["H"], c["W"]), c["norm_mask"])
    assert torch.equal(lbl_o.to(torch.uint8), fx["pred_lbl"])
    top2 = dist.topk(min(2, dist.shape[-1]), dim=-1).values
    gap = (top2[..., 0] - top2[..., -1]).abs() if dist.shape[-1] > 1 else torch.ones(dist.shape[:2])
    diff = cls.long() != cls_o
    assert int(diff.sum()) == 0 or float(gap[diff].max()) < 1e-5, (int(diff.sum()), float(gap[diff].max()))
    assert float(diff.float().mean()) < 1e-3
    same = ~diff
    assert torch.equal(rgb[same], fx["pred_lbl"][same])


@pytest.mark.parametrize("name", list(cases.POST_CASES))
def test_label_postprocessing_vs_reference_golden(sim, name):
    pytest.importorskip("cv2")
    c = cases.POST_CASES[name]
    fx = load(name)
    pred, lbl_set, img = cases.post_inputs(c)
    pal = lbl_set.to(torch.uint8).contiguous()
    cls = torch.empty(c["H"], c["W"], dtype=torch.uint8)
    rgb = torch.empty(c["H"], c["W"], 3, dtype=torch.uint8)
    p = pred.contiguous()
    sim.check(sim.crw_lp_upsample_argmax(ptr(p), 1, c["h"], c["w"], c["L"], c["H"], c["W"], int(c["norm_mask"]), ptr(pal), ptr(cls), ptr(rgb), None))
    check_label_images(cls, rgb, pred, lbl_set, c, fx)
    # several frames in one launch, class map only
    p2 = torch.stack([pred, pred.flip(0)]).contiguous()
    cls2 = torch.empty(2, c["H"], c["W"], dtype=torch.uint8)
    sim.check(sim.crw_lp_upsample_argmax(ptr(p2), 2, c["h"], c["w"], c["L"], c["H"], c["W"], int(c["norm_mask"]), None, ptr(cls2), None, None))
    assert torch.equal(cls2[0], cls)


def check_topk_indices(feats, ki, Is, Is_ref, c):
    """Top-k indices must be bit-exact apart from DOCUMENTED EXACT TIES.  torch.topk's order among equal scores is
    unspecified (ours: lowest flat index first), and equal scores are systematic here: target 0's long-memory frame 0
    is also its first short-term frame (test_utils.py:129-145), replicated first frames (vos.py:148-149) are identical
    keys, and the dyadic-grid features collide by construction.  So: wherever our index differs from the
    reference's, the two keys must have EXACTLY the same score under the oracle (for dyadic features that is exact
    arithmetic), and the index sets of a query must be duplicate-free; masked (out-of-radius) keys score -1e10/tau
    and can therefore never pass."""
    hw = c["h"] * c["w"]
    f = feats[0].flatten(-2)                                     # (C, Nf, hw)
    add = O.radius_mask_additive(c["h"], c["w"], c["radius"])[0, 0]
    n_long = len(c["long_mem"])
    n_diff = 0
    for n in range(Is.shape[0]):
        assert (Is[n] >= 0).all() and (Is[n] < ki.shape[1] * hw).all()
        sc = torch.cat([f[:, ki[n, s]].t() @ f[:, n + c["n_ctx"]] + (add if s >= n_long else 0) for s in range(ki.shape[1])], 0) / c["tau"]
        mine = torch.gather(sc, 0, Is[n])
        ref = torch.gather(sc, 0, Is_ref[n])
        assert torch.equal(mine, ref), "target %d: selected scores differ" % n
        assert (mine[:-1] >= mine[1:]).all(), "not sorted"
        srt = Is[n].sort(0).values
        assert (srt[1:] != srt[:-1]).all(), "duplicate index"
        neq = Is[n] != Is_ref[n]
        n_diff += int(neq.sum())
        if neq.any():     # every differing pick is a tie: some OTHER candidate has the identical score
            qs = neq.nonzero()[:, 1].unique()
            for q in qs.tolist():
                col = sc[:, q]
                for v in mine[neq[:, q], q].tolist():
                    assert int((col == v).sum()) >= 2
    return n_diff


def run_lp(sim, feats, c, dense=None):
    C, Nf = feats.shape[1], feats.shape[2]
    h, w = c["h"], c["w"]
    hw = h * w
    cf = feats[0].reshape(C, Nf, hw).contiguous()
    cl = torch.empty(Nf, hw, C)
    sim.check(sim.crw_lp_prepare(ptr(cf), C, Nf, hw, 0, ptr(cl), None))
    assert torch.equal(cl, cf.permute(1, 2, 0))
    ki = O.context_index_bank(c["n_ctx"], c["long_mem"], c["n_tgt"]).contiguous()
    qf = (torch.arange(c["n_tgt"]) + c["n_ctx"]).contiguous()
    Nt, S = ki.shape
    Ws = torch.empty(Nt, c["k"], hw)
    Is = torch.empty(Nt, c["k"], hw, dtype=torch.int64)
    ws = torch.zeros(256, dtype=torch.uint8)
    sim.check(sim.crw_lp_topk(ptr(cl), Nf, ptr(ki), ptr(qf), Nt, S, len(c["long_mem"]), h, w, C, float(c["radius"]), ptr(dense), c["tau"],
                              c["k"], 0, ptr(Ws), ptr(Is), ptr(ws), 256, None))
    return cl, ki, Ws, Is


@pytest.mark.parametrize("name", list(cases.LP_CASES))
def test_label_prop_vs_reference_golden(sim, name):
    c = cases.LP_CASES[name]
    fx = load(name)
    feats, lbls = cases.lp_inputs(c)
    cl, ki, Ws, Is = run_lp(sim, feats, c)
    hw = c["h"] * c["w"]
    check_topk_indices(feats, ki, Is, fx["Is"], c)
    torch.testing.assert_close(Ws, fx["Ws"], rtol=1e-5, atol=1e-7)
    # propagation (test.py:141-160)
    L = c["L"]
    lb = lbls.clone()
    lb[c["n_ctx"]:] *= 0
    lb = lb.reshape(-1, hw, L).contiguous()
    preds = []
    for t in range(ki.shape[0]):
        if t > 0:
            sim.check(sim.crw_lp_gather(ptr(lb), ptr(ki[t]), ptr(Ws[t]), ptr(Is[t]), hw, L, c["k"], t + c["n_ctx"], None))
        else:
            lb[c["n_ctx"]] = lb[0]
        preds.append(lb[t + c["n_ctx"]].clone())
    preds = torch.stack(preds).view(-1, c["h"], c["w"], L)
    torch.testing.assert_close(preds, fx["preds"], rtol=1e-5, atol=1e-6)


def test_label_prop_dense_mask_equals_radius_path(sim):
    """the reference's literal semantics (dense additive mask, every key visited) == the window-skipping path."""
    c = cases.LP_CASES["lp_small"]
    feats, _ = cases.lp_inputs(c)
    _, _, Ws, Is = run_lp(sim, feats, c)
    dense = O.radius_mask_additive(c["h"], c["w"], c["radius"])[0, 0].contiguous()
    _, _, Ws2, Is2 = run_lp(sim, feats, dict(c, radius=0), dense=dense)
    assert torch.equal(Is, Is2) and torch.equal(Ws, Ws2)


def test_l2norm(sim):
    torch.manual_seed(1)
    f = torch.randn(37, 48)
    f[3] = 0
    q, inv, nr = torch.empty_like(f), torch.empty(37), torch.empty(37)
    sim.check(sim.crw_l2norm_fwd(ptr(f), ptr(q), ptr(inv), ptr(nr), 37, 48, None))
    fo = f.clone().requires_grad_(True)
    qo = torch.nn.functional.normalize(fo, dim=1)
    torch.testing.assert_close(q, qo.detach(), rtol=1e-6, atol=1e-7)
    g = torch.randn(37, 48)
    qo.backward(g)
    gi = g.clone()
    sim.check(sim.crw_l2norm_bwd(ptr(q), ptr(gi), ptr(inv), ptr(nr), 37, 48, None))
    torch.testing.assert_close(gi[4:], fo.grad[4:], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("flags", [0, _lib.WALK_FORCE_GENERAL])
def test_walk_device_rng_state(sim, flags):
    """in-kernel Philox dropout: host-passed (seed, offset) == device-resident state, which the launch then advances."""
    torch.manual_seed(4)
    B, N, T, D = 2, 20, 4, 32
    f = torch.randn(B, N, T, D)
    wsb = sim.crw_walk_workspace_bytes(B, N, T, D, flags)
    ws = torch.zeros(wsb, dtype=torch.uint8)
    thr = 256 * ((B * N * N + 255) // 256)

    def run(seed, off, state):
        q, g = torch.empty_like(f), torch.empty_like(f)
        xe, ac = torch.zeros(T - 1), torch.zeros(T - 2)
        sim.check(sim.crw_walk_fwd_bwd(ptr(f), B, N, T, D, 0.07, 0.3, None, None, seed, off, thr, ptr(state), flags, ptr(q),
                                       ptr(xe), ptr(ac), ptr(g), ptr(ws), wsb, None))
        return xe[: T - 2].clone(), g

    xe1, g1 = run(42, 8, None)
    state = torch.tensor([42, 8], dtype=torch.int64)
    xe2, g2 = run(0, 0, state)
    assert torch.equal(xe1, xe2) and torch.equal(g1, g2)
    assert state.tolist() == [42, 8 + 4 * 2 * (T - 1)]
    xe3, _ = run(0, 0, state)                 # second replay: fresh masks
    assert not torch.equal(xe3, xe2)
    xe4, _ = run(42, 8 + 4 * 2 * (T - 1), None)
    assert torch.equal(xe3, xe4)
    # and the uniforms path agrees with an explicit draw of the same stream
    u = torch.empty(2 * (T - 1), B * N * N)
    for j in range(2 * (T - 1)):
        sim.check(sim.crw_philox_uniform(ptr(u[j]), B * N * N, 42, 8 + 4 * j, thr, None))
    u12 = u[: T - 1].view(T - 1, B, N, N).contiguous()
    u21p = u[T - 1:].view(T - 1, B, N, N).contiguous()
    q, xe5, _, g5 = run_walk(sim, f, 0.07, 0.3, u12, u21p, flags)
    assert torch.equal(xe5, xe1) and torch.equal(g5, g1)


def test_head_wgrad_splitk(sim):
    torch.manual_seed(0)
    R, D, C = 389, 128, 96
    g, x = torch.randn(R, D), torch.randn(R, C)
    nb = sim.crw_head_wgrad_workspace_bytes(R, D, C)
    ws = torch.empty(nb, dtype=torch.uint8)
    dW = torch.empty(D, C)
    sim.check(sim.crw_head_wgrad(ptr(g), ptr(x), ptr(dW), R, D, C, ptr(ws), nb, None))
    torch.testing.assert_close(dW, g.t() @ x, rtol=1e-4, atol=1e-4)


def test_patch_grid_kernel_matches_reference_golden(sim):
    """crw_patch_grid (csrc/patchgrid.cu) on the host simulator against the reference's own patch_grid output
    (utils/augs.py:59-82, tests/golden/pg_160x128.pt): bit-exact floats; several frames per launch."""
    import ctypes
    import numpy as np
    from sapienza_video_contrastive_b200.augs import IMG_MEAN, IMG_STD, draw_patch_boxes
    c = cases.PG_CASE
    fx = load("pg_160x128")
    frame = cases.pg_frame(c)
    np.random.seed(c["np_seed"])
    torch.manual_seed(c["torch_seed"])
    boxes = draw_patch_boxes(1, 12, 64)
    frames = torch.stack([frame, frame.flip(0)]).contiguous()
    bx = torch.cat([boxes, boxes]).contiguous()
    out = torch.empty(2, 36, 64, 64)
    m3, s3 = (ctypes.c_float * 3)(*IMG_MEAN), (ctypes.c_float * 3)(*IMG_STD)
    sim.check(sim.crw_patch_grid(ptr(frames), ptr(bx), 2, c["H"], c["W"], 64, 32, 64, ctypes.addressof(m3), ctypes.addressof(s3), ptr(out), None))
    mean, std = torch.tensor(IMG_MEAN)[:, None, None], torch.tensor(IMG_STD)[:, None, None]
    ref = ((fx["patches_u8"].float().div(255) - mean) / std).view(-1, 64, 64)
    assert torch.equal(out[0], ref)
    assert torch.equal(out[1], O.patch_grid(frame.flip(0).numpy(), boxes[0].numpy()))
    assert sim.crw_patch_grid(ptr(frames), ptr(bx), 2, c["H"], c["W"], 96, 32, 96, ctypes.addressof(m3), ctypes.addressof(s3), ptr(out), None) != 0


@pytest.mark.parametrize("softmax", [False, True])
def test_stoch_mat_backward_matches_autograd(sim, softmax):
    """crw_stoch_mat_bwd against torch autograd through the oracle's restatement of model.py:74-90 (dropped edges: no gradient)."""
    torch.manual_seed(3)
    R, N, M, tau, rate = 3, 9, 11, 0.07, 0.3
    A = torch.randn(R, N, M) * 0.3
    u = torch.rand(R, N, M)
    g = torch.randn(R, N, M)
    Ao = A.clone().requires_grad_(True)
    yo = O.stoch_rows(Ao, u < rate, tau, softmax)
    yo.backward(g)
    work, out, gA = A.clone(), torch.empty(R, N, M), torch.empty(R, N, M)
    flags = 1 if softmax else 0
    sim.check(sim.crw_stoch_mat(ptr(work), ptr(u), rate, tau, flags, R, N, M, ptr(out), None))
    torch.testing.assert_close(out, yo.detach(), rtol=1e-5, atol=1e-7)
    assert bool((work[u < rate] == -1e20).all())
    sim.check(sim.crw_stoch_mat_bwd(ptr(work), ptr(out), ptr(g), tau, flags, R, N, M, ptr(gA), None))
    assert float((gA - Ao.grad).abs().max() / Ao.grad.abs().max()) < 1e-5       # (cancellation in g - sum g y: compare at the scale of the row)
    assert float(gA[u < rate].abs().max()) == 0.0


def test_lp_gather_all_equals_frame_by_frame(sim):
    """crw_lp_gather_all (the whole loop of test.py:145-157 in one call) == crw_lp_gather applied target by target, bit for bit."""
    c = cases.LP_CASES["lp_small"]
    fx = load("lp_small")
    feats, lbls = cases.lp_inputs(c)
    hw, L, n_ctx, Nt = c["h"] * c["w"], c["L"], c["n_ctx"], c["n_tgt"]
    ki = O.context_index_bank(n_ctx, c["long_mem"], Nt)
    Ws, Is = fx["Ws"].contiguous(), fx["Is"].contiguous()
    res = []
    for mode in ("loop", "all"):
        lb = lbls.clone()
        lb[n_ctx:] = 0
        lb = lb.view(-1, hw, L).contiguous()
        lb[n_ctx] = lb[0]
        if mode == "loop":
            for t in range(1, Nt):
                sim.check(sim.crw_lp_gather(ptr(lb), ptr(ki[t].contiguous()), ptr(Ws[t]), ptr(Is[t]), hw, L, c["k"], t + n_ctx, None))
        else:
            sim.check(sim.crw_lp_gather_all(ptr(lb), ptr(ki.contiguous()), ptr(Ws), ptr(Is), Nt, ki.shape[1], hw, L, c["k"], 1, n_ctx, None))
        res.append(lb[n_ctx:].clone())
    assert torch.equal(res[0], res[1])
    torch.testing.assert_close(res[1].view(Nt, c["h"], c["w"], L), fx["preds"], rtol=1e-5, atol=1e-6)
