"""Parity tests proper: the sm_100a kernels, called through the C ABI (sapienza_video_contrastive_b200.ops ->
libcrw_b200.so), against the oracle on identical seeded inputs and against the golden outputs of the unmodified
reference (tests/golden).  Run on the B200 box: pytest -m gpu.

Tolerances (fp32 everywhere): loss / xent rel 1e-5 vs the reference, gradients rel 1e-4 of the largest entry
(BASELINE.json north_star: 1e-4), top-k indices bit-exact apart from exact score ties, label maps 1e-5.
"""
import argparse
import os

import pytest
import torch

from oracle import crw_oracle as O
from tests.golden import cases
from tests.test_sim_kernels import check_label_images, check_topk_indices, load

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from sapienza_video_contrastive_b200 import ops as _ops
    _ops.check_device(DEV)
    return _ops


def relmax(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_library_is_the_cuda_one(ops):
    from sapienza_video_contrastive_b200 import _lib
    L = _lib.lib()
    assert L.path.endswith("libcrw_b200.so") and L.crw_version() >= 100
    with pytest.raises(RuntimeError):
        ops.pool_patch(torch.zeros(2, 2, 8, 8))                     # CPU tensors are refused: no fallback


def test_philox_replay_matches_torch_rand(ops):
    assert ops.philox_replay_ok(DEV)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(1234)
    for numel in (7, 48020, 20 * 49 * 49, 303104, 303105, 2_000_003):
        seed, off = gen.initial_seed(), gen.get_offset()
        ref = torch.rand(numel, device=DEV)
        thr = ops.torch_rand_threads(numel, DEV)
        mine = ops.philox_uniform(numel, seed, off, thr, DEV)
        assert torch.equal(ref, mine), numel
        assert gen.get_offset() == off + ops.torch_rand_offset_increment(numel, thr)
    # rand_like of a transposed view keeps the transposed physical layout, also on CUDA
    torch.manual_seed(5)
    a = torch.rand(3, 5, 5, device=DEV)
    torch.manual_seed(5)
    base = torch.zeros(3, 2, 5, 5, device=DEV)
    b = torch.rand_like(base[:, 0].transpose(-1, -2))
    assert torch.equal(a.transpose(-1, -2), b)


@pytest.mark.parametrize("rows,hw", [(980 * 32 * 4, 64), (1001, 64), (777, 32), (50, 16), (33, 49)])
def test_pool_patch(ops, rows, hw):
    torch.manual_seed(rows)
    x = torch.randn(rows, hw)
    xd = x.to(DEV).view(rows, 1, hw).requires_grad_(True)
    out = ops.pool_patch(xd)
    torch.testing.assert_close(out.cpu().view(-1), O.patch_pool(x.view(rows, 1, 1, 1, hw)).view(-1), rtol=1e-6, atol=1e-6)
    g = torch.randn(rows, device=DEV)
    out.view(-1).backward(g)
    # (torch's CUDA `g / hw` multiplies by the reciprocal; the kernel divides like the reference's CPU path: 1 ulp)
    torch.testing.assert_close(xd.grad.view(rows, hw), (g / hw)[:, None].expand(rows, hw), rtol=2e-7, atol=0)


def test_pool_patch_full_size_linearity(ops):
    """BASELINE config 2 size (980,512,4,8,8): mean-pool is linear and reproduces constants exactly."""
    torch.manual_seed(0)
    a = torch.randn(980, 512, 4, 8, 8, device=DEV)
    pa = ops.pool_patch(a)
    assert pa.shape == (980, 512, 4)
    torch.testing.assert_close(pa, a.double().mean((-1, -2)).float(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ops.pool_patch(2 * a), 2 * pa, rtol=0, atol=0)
    assert torch.equal(ops.pool_patch(torch.full_like(a, 3.0)), torch.full_like(pa, 3.0))


def gpu_walk_from_case(ops, c, rng="explicit", force_general=False):
    maps, head_w = cases.walk_inputs(c)
    maps_d = maps.to(DEV).requires_grad_(True)
    head_d = head_w.to(DEV).requires_grad_(True)
    pooled = ops.pool_patch(maps_d)                                           # (BN, Ce, T)
    f = (pooled.transpose(-1, -2) @ head_d.t()).reshape(c["B"], c["N"], c["T"], 128)
    torch.manual_seed(c["seed"] + 1000)
    u12 = u21p = None
    if c["p"] > 0:
        u12, u21p = O.draw_uniforms(c["B"], c["N"], c["T"])
        u12, u21p = u12.to(DEV), u21p.to(DEV)
    q, loss, xent, acc = ops.walk(f, c["tau"], c["p"], flip=c["flip"], u12=u12, u21p=u21p, force_general=force_general)
    loss.sum().backward()
    return q, loss, xent, acc, maps_d.grad, head_d.grad


@pytest.mark.parametrize("name", list(cases.WALK_CASES))
@pytest.mark.parametrize("force_general", [False, True])
def test_walk_matches_reference_golden(ops, name, force_general):
    c = cases.WALK_CASES[name]
    fx = load(name)
    q, loss, xent, acc, gmaps, ghead = gpu_walk_from_case(ops, c, force_general=force_general)
    torch.testing.assert_close(q.permute(0, 3, 2, 1).cpu(), fx["q"], rtol=1e-5, atol=1e-6)
    assert loss.shape == (1,)
    torch.testing.assert_close(loss.cpu(), fx["loss"], rtol=1e-5, atol=0)
    tag = "l" if c["flip"] else "r"
    for j in range(c["T"] - 2):
        torch.testing.assert_close(xent[j].cpu(), fx["diags"]["64 xent cyc %s%d" % (tag, j + 1)], rtol=1e-5, atol=0)
        torch.testing.assert_close(acc[j].cpu(), fx["diags"]["64 acc cyc %s%d" % (tag, j + 1)], rtol=0, atol=1e-6)
    assert relmax(ghead.cpu(), fx["grad_head"]) < 1e-4
    assert relmax(gmaps[..., 0, 0].cpu(), fx["grad_maps00"]) < 1e-4
    assert float((gmaps - gmaps[..., :1, :1]).abs().max()) == 0


@pytest.mark.parametrize("B,N,T,D,p,flip,softmax", [(20, 49, 4, 128, 0.1, False, False), (4, 49, 8, 128, 0.1, False, False),
                                                    (3, 64, 4, 128, 0.1, True, False), (2, 128, 4, 128, 0.1, False, False),
                                                    (2, 100, 6, 128, 0.2, False, True), (1, 256, 5, 128, 0.1, False, False),
                                                    (2, 31, 3, 64, 0.5, False, False), (2, 17, 2, 32, 0.1, False, False),
                                                    (2, 131, 3, 128, 0.1, False, False), (1, 90, 4, 96, 0.1, True, False)])
def test_walk_matches_oracle(ops, B, N, T, D, p, flip, softmax):
    torch.manual_seed(B * 100 + N)
    f = torch.randn(B, N, T, D)
    u12, u21p = O.draw_uniforms(B, N, T)
    fo = f.clone().requires_grad_(True)
    qo = (fo / fo.norm(dim=-1, keepdim=True).clamp_min(1e-12)).permute(0, 3, 2, 1)
    loss_o, xents, accs, _ = O.walk_loss(qo, 0.07, p, u12, u21p, flip=flip, softmax=softmax)
    if T >= 3:
        loss_o.sum().backward()
    for fg in (False, True):
        fd = f.to(DEV).requires_grad_(True)
        q, loss, xent, acc = ops.walk(fd, 0.07, p, flip=flip, softmax=softmax, u12=u12.to(DEV), u21p=u21p.to(DEV), force_general=fg)
        torch.testing.assert_close(q.cpu(), qo.detach().permute(0, 3, 2, 1), rtol=1e-5, atol=1e-6)
        if T >= 3:
            torch.testing.assert_close(xent.cpu(), torch.stack(xents).detach(), rtol=2e-5, atol=0)
            # argmax accuracy can flip on a near-tie: allow one row
            assert float((acc.cpu() - torch.stack(accs)).abs().max()) <= 1.0 / (B * N) + 1e-6
            loss.sum().backward()
            assert relmax(fd.grad.cpu(), fo.grad) < 1e-4
        else:
            assert float(loss) == 0.0


@pytest.mark.parametrize("Z,M,N,K,ta,tb", [(3, 128, 128, 128, False, False), (2, 256, 128, 256, False, True), (2, 200, 72, 100, True, False),
                                            (1, 1024, 1024, 1024, True, True), (5, 130, 129, 67, False, False),
                                            (2, 512, 128, 512, True, False)])
def test_bmm_tc_is_fp32_faithful(ops, Z, M, N, K, ta, tb):
    """The tensor-core GEMM of the large-graph walk against an fp64 product: its error must be of the order of an fp32
    sgemm's, for plain operands, for rows of wildly different magnitude (per-row scaling), and when accumulating."""
    torch.manual_seed(Z * 1000 + M + N + K)
    A = torch.randn(Z, M, K, device=DEV)
    B = torch.randn(Z, K, N, device=DEV)
    A *= torch.exp(6 * torch.randn(Z, M, 1, device=DEV))              # per-row magnitudes over ~10 decades
    B *= torch.exp(6 * torch.randn(Z, 1, N, device=DEV))
    ref = A.double() @ B.double()
    bound = (A.double().abs() @ B.double().abs())                      # elementwise error scale |a|.|b|
    As = A.transpose(1, 2).contiguous() if ta else A
    Bs = B.transpose(1, 2).contiguous() if tb else B
    C = ops.bmm_tc(As, Bs, trans_a=ta, trans_b=tb)
    err_tc = ((C.double() - ref).abs() / bound).max().item()
    err_f32 = (((A @ B).double() - ref).abs() / bound).max().item()
    assert err_tc < 2e-6, (err_tc, err_f32)
    C2 = ops.bmm_tc(As, Bs, trans_a=ta, trans_b=tb, out=C.clone(), accumulate=True)
    assert ((C2.double() - 2 * ref).abs() / bound).max().item() < 4e-6
    # exactness on small integers (every partial product is representable)
    Ai = torch.randint(-8, 9, (Z, M, K), device=DEV).float()
    Bi = torch.randint(-8, 9, (Z, K, N), device=DEV).float()
    Ci = ops.bmm_tc(Ai.transpose(1, 2).contiguous() if ta else Ai, Bi.transpose(1, 2).contiguous() if tb else Bi, trans_a=ta, trans_b=tb)
    assert torch.equal(Ci, Ai @ Bi)


@pytest.mark.parametrize("Z,M,N,K,ta,tb", [(3, 128, 128, 128, False, False), (2, 256, 128, 256, False, True), (2, 200, 72, 100, True, False),
                                            (1, 384, 256, 196, True, True), (5, 132, 128, 68, False, False), (2, 196, 196, 196, True, False),
                                            (2, 196, 128, 392, False, False)])
def test_bmm_tf32_is_fp32_faithful(ops, Z, M, N, K, ta, tb):
    """The fused kind::tf32 GEMM (operands read in place, K-major or MN-major, split inside the kernel) against an fp64 product."""
    torch.manual_seed(Z * 1000 + M + N + K)
    A = torch.randn(Z, M, K, device=DEV)
    B = torch.randn(Z, K, N, device=DEV)
    A *= torch.exp(6 * torch.randn(Z, M, 1, device=DEV))
    B *= torch.exp(6 * torch.randn(Z, 1, N, device=DEV))
    ref = A.double() @ B.double()
    bound = (A.double().abs() @ B.double().abs())
    As = A.transpose(1, 2).contiguous() if ta else A
    Bs = B.transpose(1, 2).contiguous() if tb else B
    C = ops.bmm_tf32(As, Bs, trans_a=ta, trans_b=tb)
    err = ((C.double() - ref).abs() / bound).max().item()
    assert err < 2e-6, err
    C2 = ops.bmm_tf32(As, Bs, trans_a=ta, trans_b=tb, out=C.clone(), accumulate=True)
    assert ((C2.double() - 2 * ref).abs() / bound).max().item() < 4e-6
    Ai = torch.randint(-8, 9, (Z, M, K), device=DEV).float()
    Bi = torch.randint(-8, 9, (Z, K, N), device=DEV).float()
    Ci = ops.bmm_tf32(Ai.transpose(1, 2).contiguous() if ta else Ai, Bi.transpose(1, 2).contiguous() if tb else Bi, trans_a=ta, trans_b=tb)
    assert torch.equal(Ci, Ai @ Bi)


@pytest.mark.parametrize("B,N,T,flip", [(2, 192, 4, False), (1, 320, 5, True), (2, 200, 3, False), (1, 515, 4, False)])
def test_walk_large_graph_tensor_core_vs_simt(ops, B, N, T, flip):
    """Large graphs: the tcgen05 GEMM path and the exact-fp32 SIMT path of the same walk agree to fp32 noise."""
    torch.manual_seed(N)
    f = torch.randn(B, N, T, 128, device=DEV)
    u12, u21p = O.draw_uniforms(B, N, T)
    res = []
    for simt in (True, False):
        fd = f.clone().requires_grad_(True)
        q, loss, xent, acc = ops.walk(fd, 0.07, 0.1, flip=flip, u12=u12.to(DEV), u21p=u21p.to(DEV), force_simt=simt)
        loss.sum().backward()
        res.append((loss.detach(), xent.detach(), fd.grad))
    torch.testing.assert_close(res[0][1], res[1][1], rtol=1e-5, atol=0)
    assert relmax(res[0][2], res[1][2]) < 5e-5


@pytest.mark.parametrize("B,N,T,flip", [(20, 49, 4, False), (3, 64, 4, True), (5, 33, 6, False), (2, 8, 3, False), (7, 49, 5, True),
                                        (37, 49, 4, False), (40, 49, 4, False)])
def test_walk_chain_cluster_vs_single_cta(ops, B, N, T, flip):
    """Small graphs: the chain split over a 4-CTA cluster (distributed shared memory) against the one-CTA-per-clip chain."""
    torch.manual_seed(N + T)
    f = torch.randn(B, N, T, 128, device=DEV)
    u12, u21p = O.draw_uniforms(B, N, T)
    res = []
    for no_cluster in (True, False):
        fd = f.clone().requires_grad_(True)
        q, loss, xent, acc = ops.walk(fd, 0.07, 0.1, flip=flip, u12=u12.to(DEV), u21p=u21p.to(DEV), no_cluster=no_cluster)
        loss.sum().backward()
        res.append((loss.detach(), xent.detach(), acc.detach(), fd.grad))
    torch.testing.assert_close(res[0][1], res[1][1], rtol=2e-6, atol=0)
    torch.testing.assert_close(res[0][2], res[1][2], rtol=0, atol=1e-6)
    assert relmax(res[0][3], res[1][3]) < 2e-6


def test_walk_in_kernel_dropout_equals_torch_draws(ops):
    """rng='philox' (drawn inside the kernel) must give exactly the run that rng='torch' (torch.rand draws handed to
    the kernel) gives from the same generator state, and leave the generator in the same state."""
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    for (B, N, T, fg) in [(20, 49, 4, False), (3, 100, 5, True), (200, 49, 4, False)]:
        f = torch.randn(B, N, T, 128, device=DEV)
        outs = []
        for rng in ("torch", "philox"):
            torch.manual_seed(77)
            fd = f.clone().requires_grad_(True)
            q, loss, xent, acc = ops.walk(fd, 0.07, 0.1, rng=rng, force_general=fg)
            loss.sum().backward()
            outs.append((loss.clone(), xent.clone(), fd.grad.clone(), gen.get_offset(), torch.rand(3, device=DEV)))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        assert torch.equal(outs[0][2], outs[1][2])
        assert outs[0][3] == outs[1][3] and torch.equal(outs[0][4], outs[1][4])


def test_walk_full_size_config2_fused_vs_general(ops):
    """BASELINE config 2 shape (B=20,N=49,T=4,D=128): the two implementations agree; loss near log N at random init."""
    torch.manual_seed(2)
    f = torch.randn(20, 49, 4, 128, device=DEV)
    res = []
    for fg in (False, True):
        torch.manual_seed(9)
        fd = f.clone().requires_grad_(True)
        q, loss, xent, acc = ops.walk(fd, 0.07, 0.1, rng="philox", force_general=fg)
        loss.sum().backward()
        res.append((loss, xent, fd.grad))
    torch.testing.assert_close(res[0][0], res[1][0], rtol=1e-5, atol=0)
    assert relmax(res[0][2], res[1][2]) < 1e-4
    assert 1.0 < float(res[0][0]) < 12.0


def make_args(**kw):
    d = dict(device=DEV, dropout=0.1, featdrop=0.0, temp=0.07, head_depth=0, model_type="scratch", remove_layers=[],
             dilate_superpixels=False, flip=False, sk_targets=False)
    d.update(kw)
    return argparse.Namespace(**d)


def test_crw_dropin_config1_matches_reference(ops):
    """BASELINE config 1 verbatim: CRW with a random ResNet-18 under seed 0, x ~ N(0,1) (2,4,147,64,64), dropout draws
    under seed 123.  The encoder runs on cuDNN here and on CPU in the reference, so tolerances are those of conv."""
    from sapienza_video_contrastive_b200 import CRW
    fx = load("cfg1_resnet18")
    torch.manual_seed(0)
    crw = CRW(make_args(device="cpu"))             # build on CPU: identical init stream to the reference
    assert list(crw.state_dict().keys()) == fx["state_dict_keys"]
    assert abs(float(sum(p.double().sum() for p in crw.parameters())) - fx["param_checksum"]) < 1e-6
    x = torch.randn(2, 4, 147, 64, 64)
    torch.manual_seed(123)
    u12, u21p = O.draw_uniforms(2, 49, 4)
    crw = crw.to(DEV)
    crw.args.device = DEV
    q, loss, diags = crw(x.to(DEV), None, None, walk_uniforms=(u12.to(DEV), u21p.to(DEV)))
    assert q.shape == (2, 128, 4, 49) and loss.shape == (1,)
    assert set(diags) == set(fx["diags"])
    torch.testing.assert_close(q.cpu(), fx["q"], rtol=1e-2, atol=2e-3)        # cuDNN vs CPU convolutions upstream
    torch.testing.assert_close(loss.cpu(), fx["loss"], rtol=1e-4, atol=0)
    for k in fx["diags"]:
        if "xent" in k:
            torch.testing.assert_close(diags[k].cpu(), fx["diags"][k], rtol=1e-4, atol=0)
    loss.mean().backward()
    g = crw.selfsim_fc[0].weight.grad.cpu()
    assert relmax(g, fx["grad_head"]) < 2e-2          # encoder on cuDNN vs CPU upstream; the walk itself is pinned at 1e-4 above
    assert relmax(crw.encoder.model.conv1.weight.grad.cpu(), fx["grad_conv1"]) < 0.3   # 20 conv/BN layers on cuDNN vs CPU at random init: not the path under test


def test_crw_module_dilated_superpixels(ops):
    """CRW(args) with --dilate-superpixels (model.py:38, 303-309): image_to_nodes pools with the dilated masks (checked
    against the oracle on the module's own feature maps) and the full forward / backward runs."""
    from sapienza_video_contrastive_b200 import CRW
    torch.manual_seed(0)
    crw = CRW(make_args(dilate_superpixels=True, dilation_kernel_size=11, dilation_kernel_shape="circle", dropout=0.0)).to(DEV)
    assert crw.dilation == (11, "circle")
    B, T, SP = 1, 3, 12
    g = torch.Generator().manual_seed(4)
    lab = cases.voronoi_labels(B, T, SP, 256, g, one_based=False)
    sp_mask = lab[:, :, None].repeat(1, 1, 3, 1, 1).to(DEV)
    x = torch.randn(B, T, 3, 256, 256, generator=g).to(DEV)
    with torch.no_grad():
        feats, maps = crw.image_to_nodes(x, sp_mask, SP)
    assert feats.shape == (B, 128, T, SP) and maps.shape == (B, 512, T, 32, 32)
    pooled = O.segment_mean_dilated(maps.cpu(), lab, SP, 11, "circle")                     # (B,T,SP,C)
    f = pooled @ crw.selfsim_fc[0].weight.detach().cpu().t()
    ref = (f / f.norm(dim=-1, keepdim=True).clamp_min(1e-12)).permute(0, 3, 1, 2)
    torch.testing.assert_close(feats.cpu(), ref, rtol=1e-3, atol=1e-4)                     # tf32-split head vs fp32 matmul
    plain = CRW(make_args(dropout=0.0)).to(DEV)
    plain.load_state_dict(crw.state_dict())
    with torch.no_grad():
        feats0, _ = plain.image_to_nodes(x, sp_mask, SP)
    assert float((feats0 - feats).abs().max()) > 1e-3                                     # the dilation does change the nodes
    q, loss, diags = crw(x, sp_mask, SP)
    loss.mean().backward()
    assert q.shape == (B, 128, T, SP) and torch.isfinite(loss).all()
    assert crw.selfsim_fc[0].weight.grad is not None and torch.isfinite(crw.encoder.model.conv1.weight.grad).all()


def test_teacher_student_module_forward(ops):
    """CRWTeacherStudent.forward (teacherstudent.py:472-580) through the module: loss against the oracle evaluated on the
    module's own node vectors, gradients reach the student only, the reference's diag names."""
    from sapienza_video_contrastive_b200 import CRWBase, CRWTeacherStudent
    torch.manual_seed(0)
    args = make_args(alpha_teacher_student=0.4, dropout=0.1)
    teacher = CRWBase(args)
    ts = CRWTeacherStudent(args, teacher=teacher).to(DEV)
    B, T, N = 2, 4, 9
    g = torch.Generator().manual_seed(8)
    x = torch.randn(B, T, N * 3, 64, 64, generator=g).to(DEV)
    torch.manual_seed(21)
    uni = O.draw_uniforms(B, N, T) + O.draw_uniforms(B, N, T)
    q, loss, diags = ts(x, walk_uniforms=tuple(u.to(DEV) for u in uni))
    assert q.shape == (B, 128, T, N) and loss.shape == (1,)
    assert sorted(diags) == ["64 acc cyc r1", "64 acc cyc r2", "64 xent cyc r1", "64 xent cyc r2"]
    with torch.no_grad():
        xr = x.transpose(1, 2).reshape(B, N, 3, T, 64, 64)
        fs = ts._patch_nodes_prenorm(xr)[0].cpu()
        ft = ts.teacher._patch_nodes_prenorm(xr)[0].cpu()
    nrm = lambda f: torch.nn.functional.normalize(f, dim=-1).permute(0, 3, 2, 1)
    loss_o, xents_o, _ = O.teacher_student_loss(nrm(fs), nrm(ft), 0.07, 0.1, 0.4, *uni)
    torch.testing.assert_close(loss.detach().cpu(), loss_o, rtol=1e-4, atol=0)
    torch.testing.assert_close(diags["64 xent cyc r1"].cpu(), xents_o[0], rtol=1e-4, atol=0)
    loss.mean().backward()
    assert ts.selfsim_fc[0].weight.grad is not None and ts.selfsim_fc[0].bias.grad is not None
    assert all(p.grad is None for p in ts.teacher.parameters())


def test_crw_dropin_api_surface(ops):
    from sapienza_video_contrastive_b200 import CRW
    torch.manual_seed(0)
    crw = CRW(make_args()).to(DEV)
    # affinity: 4-D and 3-D forms (model.py:63-72)
    x1, x2 = torch.randn(2, 16, 3, 7, device=DEV), torch.randn(2, 16, 3, 9, device=DEV)
    torch.testing.assert_close(crw.affinity(x1, x2), torch.einsum("bctn,bctm->btnm", x1, x2), rtol=1e-5, atol=1e-5)
    assert crw.affinity(x1[:, :, 0], x2[:, :, 0]).shape == (2, 7, 9)
    # stoch_mat mutates its argument through a transposed view, like the reference (F4)
    As = torch.randn(2, 3, 6, 6, device=DEV)
    A0 = As.clone()
    torch.manual_seed(3)
    view = As[:, 1].transpose(-1, -2)
    out = crw.stoch_mat(view, do_dropout=True)
    torch.manual_seed(3)
    u = torch.rand_like(A0[:, 1].transpose(-1, -2))
    exp_view = A0[:, 1].transpose(-1, -2).masked_fill(u < 0.1, -1e20)
    assert torch.equal(As[:, 1].transpose(-1, -2), exp_view) and torch.equal(As[:, 0], A0[:, 0])
    torch.testing.assert_close(out.cpu(), O.stoch_rows(A0[:, 1].transpose(-1, -2).cpu(), (u < 0.1).cpu(), 0.07), rtol=1e-5, atol=1e-8)
    # pixels_to_nodes returns unit-norm (B,128,T,N) and the maps
    x = torch.randn(1, 4, 3, 2, 64, 64, device=DEV)
    feats, maps = crw.pixels_to_nodes(x)
    assert feats.shape == (1, 128, 2, 4) and maps.shape == (1, 4, 512, 2, 8, 8)
    torch.testing.assert_close(feats.norm(dim=1), torch.ones(1, 2, 4, device=DEV), rtol=1e-5, atol=1e-5)
    assert torch.equal(crw.xent_targets(torch.zeros(2, 5, 5, device=DEV)).cpu(), torch.arange(5).repeat(2))
    # default rng mode: a full forward/backward under the module API
    v = torch.randn(2, 4, 49 * 3, 64, 64, device=DEV)
    q, loss, diags = crw(v)
    loss.mean().backward()
    assert torch.isfinite(loss).all() and crw.selfsim_fc[0].weight.grad is not None
    assert sorted(diags) == ["64 acc cyc r1", "64 acc cyc r2", "64 xent cyc r1", "64 xent cyc r2"]


@pytest.mark.parametrize("name", list(cases.SP_CASES))
def test_superpixel_nodes_and_walk_match_reference(ops, name):
    c = cases.SP_CASES[name]
    fx = load(name)
    maps, lab3, head_w = cases.sp_inputs(c)
    md = maps.to(DEV).requires_grad_(True)
    hd = head_w.to(DEV).requires_grad_(True)
    lab = lab3.to(DEV)[:, :, 0]
    pooled = ops.segment_mean(md, lab, c["SP"])                               # (B,SP,T,Ce)
    torch.testing.assert_close(pooled.detach().cpu().transpose(1, 2), O.segment_mean(maps, lab3[:, :, 0], c["SP"]), rtol=1e-5, atol=1e-6)
    f = pooled @ hd.t()
    torch.manual_seed(c["seed"] + 1000)
    u12, u21p = O.draw_uniforms(c["B"], c["SP"], c["T"])
    q, loss, xent, acc = ops.walk(f, c["tau"], c["p"], u12=u12.to(DEV), u21p=u21p.to(DEV))
    torch.testing.assert_close(q.permute(0, 3, 2, 1).cpu(), fx["q"], rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(loss.cpu(), fx["loss"], rtol=1e-5, atol=0)
    loss.sum().backward()
    assert relmax(hd.grad.cpu(), fx["grad_head"]) < 1e-4
    assert relmax(md.grad.cpu(), fx["grad_maps"]) < 1e-4


@pytest.mark.parametrize("name", list(cases.POSE_CASES))
def test_pose_coords_match_reference(ops, name):
    """utils/test_utils.py:60-84 (process_pose, the JHMDB branch of test.py:171-172): key-point coordinates bit-equal to the
    reference's, through the operator (several frames per launch) and through the reference-signature mirror."""
    from sapienza_video_contrastive_b200 import test_utils as TU
    c = cases.POSE_CASES[name]
    fx = load(name)
    pred, lbl_set = cases.pose_inputs(c)
    batch = torch.stack([pred, pred.flip(0), pred * 0.5]).to(DEV)
    coords = ops.lp_pose_coords(batch).cpu()
    assert torch.equal(coords[0], fx["coords"])
    assert torch.equal(coords[1], O.process_pose(pred.flip(0), lbl_set.numpy())[0])
    assert torch.equal(coords[2], O.process_pose(pred * 0.5, lbl_set.numpy())[0])
    cc, sharp = TU.process_pose(pred, lbl_set.numpy())
    assert torch.equal(cc, fx["coords"]) and torch.equal(torch.from_numpy(sharp), fx["sharp"])
    for k in (1, 4):
        assert torch.equal(ops.lp_pose_coords(pred.to(DEV), k)[0].cpu(), O.process_pose(pred, lbl_set.numpy(), topk=k)[0])


@pytest.mark.parametrize("name", list(cases.TS_CASES))
def test_teacher_student_walk_matches_reference(ops, name):
    """teacherstudent.py:472-580 from the node vectors on (crw_walk_ts_fwd_bwd): loss, walk diags and the gradient of the
    student's node vectors against the reference's CRWTeacherStudent.forward; teacher-student terms against the oracle."""
    c = cases.TS_CASES[name]
    fx = load(name)
    fs, ft = cases.ts_inputs(c)
    B, N, T, D = fs.shape
    torch.manual_seed(c["seed"] + 1000)
    us12, us21p = O.draw_uniforms(B, N, T)
    ut12, ut21p = O.draw_uniforms(B, N, T)
    fd = fs.to(DEV).requires_grad_(True)
    uni = tuple(u.to(DEV) for u in (us12, us21p, ut12, ut21p))
    q, loss, xent, acc, tsx = ops.walk_teacher_student(fd, ft.to(DEV), c["tau"], c["p"], c["alpha"], flip=c["flip"], uniforms=uni)
    torch.testing.assert_close(q.permute(0, 3, 2, 1).cpu(), fx["q"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(loss.cpu(), fx["loss"], rtol=1e-5, atol=0)
    names = [("l%d" if c["flip"] else "r%d") % i for i in range(1, T - 1)]
    for j, nm in enumerate(names):
        torch.testing.assert_close(xent[j].cpu(), fx["diags"]["8 xent cyc %s" % nm], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(acc[j].cpu(), fx["diags"]["8 acc cyc %s" % nm], rtol=0, atol=1e-6)
    qt = torch.nn.functional.normalize(ft, dim=-1).permute(0, 3, 2, 1)
    _, _, ts_o = O.teacher_student_loss(fx["q"], qt, c["tau"], c["p"], c["alpha"], us12, us21p, ut12, ut21p, flip=c["flip"])
    torch.testing.assert_close(tsx.cpu(), torch.stack(ts_o), rtol=1e-5, atol=0)
    loss.sum().backward()
    assert relmax(fd.grad.cpu(), fx["grad_feats"]) < 1e-4
    chains = ops.walk_chains(ft.to(DEV), c["tau"], c["p"], c["flip"], True, u12=uni[2], u21p=uni[3])
    Wt = torch.stack(O.walk_chain_products(qt, c["tau"], c["p"], ut12, ut21p, flip=c["flip"], softmax=True), 1)
    torch.testing.assert_close(chains.cpu(), Wt, rtol=1e-4, atol=1e-7)


def test_teacher_student_walk_generator_order(ops):
    """Without explicit draws the student's 2(T-1) dropout draws come first in the generator stream and the teacher's second,
    as in the reference - although the teacher's chains are computed first: 'philox' (in-kernel replay) and 'torch'
    (torch.rand) give the same numbers from the same seed and leave the generator in the same state."""
    if not ops.philox_replay_ok(DEV):
        pytest.skip("torch's CUDA rand launch geometry is not the one the replay assumes")
    B, N, T, D = 3, 70, 4, 64                                     # a large-graph shape: tensor-core GEMMs inside
    g = torch.Generator().manual_seed(5)
    fs, ft = torch.randn(B, N, T, D, generator=g).to(DEV), torch.randn(B, N, T, D, generator=g).to(DEV)
    outs = []
    for rng in ("torch", "philox"):
        torch.manual_seed(77)
        f = fs.clone().requires_grad_(True)
        q, loss, xent, acc, tsx = ops.walk_teacher_student(f, ft, 0.07, 0.2, 0.3, rng=rng)
        loss.sum().backward()
        outs.append((loss.detach(), tsx, f.grad, torch.rand(4, device=DEV)))
    for a, b in zip(*outs):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-9)
    assert torch.equal(outs[0][3], outs[1][3])


@pytest.mark.parametrize("name", list(cases.SPD_CASES))
def test_dilated_superpixel_nodes_match_reference(ops, name):
    """--dilate-superpixels (model.py:303-309): pooled features against the oracle, node embeddings and the gradients of
    the maps and the head against the reference's fixtures (its fp16 depthwise-convolution dilation)."""
    c = cases.SPD_CASES[name]
    fx = load(name)
    maps, lab3, head_w = cases.sp_inputs(c)
    md = maps.to(DEV).requires_grad_(True)
    hd = head_w.to(DEV).requires_grad_(True)
    lab = lab3.to(DEV)[:, :, 0]
    pooled = ops.segment_mean_dilated(md, lab, c["SP"], c["ksize"], c["shape"])          # (B,SP,T,Ce)
    ref = O.segment_mean_dilated(maps, lab3[:, :, 0], c["SP"], c["ksize"], c["shape"])
    torch.testing.assert_close(pooled.detach().cpu().transpose(1, 2), ref, rtol=1e-5, atol=1e-6)
    q = ops.l2_normalize_last(pooled @ hd.t()).permute(0, 3, 2, 1)                      # (B,D,T,SP)
    torch.testing.assert_close(q.detach().cpu(), fx["sp_feats"], rtol=1e-4, atol=2e-6)
    proj = torch.randn(q.shape, generator=torch.Generator().manual_seed(c["seed"] + 7))
    (q * proj.to(DEV)).sum().backward()
    assert relmax(md.grad.cpu(), fx["grad_maps"]) < 1e-4
    assert relmax(hd.grad.cpu(), fx["grad_head"]) < 1e-4


def test_dilated_superpixels_reference_default_element(ops):
    """The reference's default element (51 x 51 diamond, utils/arguments.py:209-210) at the BASELINE config 3 shape: a
    frame against the oracle, a 1 x 1 element is the undilated pooling, constants are reproduced for every label present,
    dilated sizes never shrink, and the module routes --dilate-superpixels here."""
    B, C, T, SP = 2, 128, 4, 196
    g = torch.Generator().manual_seed(11)
    lab = cases.voronoi_labels(B, T, SP, 256, g, one_based=True)
    maps = torch.randn(B, C, T, 32, 32, generator=g)
    md, ld = maps.to(DEV), lab.to(DEV)
    out = ops.segment_mean_dilated(md, ld, SP, 51, "L1")
    ref = O.segment_mean_dilated(maps[:1, :, :1], lab[:1, :1], SP, 51, "L1")
    torch.testing.assert_close(out[:1, :, :1].cpu().transpose(1, 2), ref, rtol=1e-5, atol=1e-6)
    assert torch.equal(ops.segment_mean_dilated(md, ld, SP, 1, "circle"), ops.segment_mean(md, ld, SP))
    ones = ops.segment_mean_dilated(torch.ones_like(md), ld, SP, 51, "cross")
    present = torch.zeros(B, T, SP, device=DEV).scatter_(2, ld.flatten(2), 1.0).transpose(1, 2)
    torch.testing.assert_close(ones[..., 0], present, rtol=1e-6, atol=1e-6)
    with pytest.raises(Exception):
        ops.segment_mean_dilated(md, ld, 300, 51, "L1")                                  # SP <= 255 in the dilated path
    with pytest.raises(ValueError):
        ops.segment_mean_dilated(md, ld, SP, 51, "square")


def test_segmean_config3_shape_properties(ops):
    """BASELINE config 3 shape: constants are reproduced, sizes partition the image, empty labels give zero rows."""
    B, C, T, SP = 2, 512, 8, 196
    g = torch.Generator().manual_seed(7)
    lab = cases.voronoi_labels(B, T, SP, 256, g, one_based=True).to(DEV)
    maps = torch.randn(B, C, T, 32, 32, device=DEV)
    out = ops.segment_mean(maps, lab, SP)
    assert out.shape == (B, SP, T, C) and float(out[:, 0].abs().max()) == 0            # label 0 unused -> empty node
    ones = ops.segment_mean(torch.ones_like(maps), lab, SP)
    present = torch.zeros(B, T, SP, device=DEV).scatter_(2, lab.flatten(2), 1.0).transpose(1, 2)
    torch.testing.assert_close(ones[..., 0], present, rtol=1e-6, atol=1e-6)
    ref = O.segment_mean(maps[:1, :64].cpu(), lab[:1].cpu(), SP)
    torch.testing.assert_close(out[:1, :, :, :64].cpu().transpose(1, 2), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B,T,C,Hm,scale,SP", [(2, 3, 128, 32, 8, 100), (1, 2, 64, 32, 4, 256), (1, 2, 64, 16, 8, 300), (1, 1, 64, 32, 2, 600),
                                               (2, 2, 96, 32, 8, 196), (1, 2, 64, 24, 4, 50)])
def test_segmean_forward_backward_vs_oracle(ops, B, T, C, Hm, scale, SP):
    """Every accumulator tier of the forward (SP up to 128 / 224 / 256 / 512 / 1024 labels), the TMA path (cells % 256 == 0,
    C % 64 == 0) and the generic cp.async path (C = 96, 24x24 maps), labels out of range, more labels than cells: forward
    against the oracle, backward against autograd through a one-hot restatement."""
    g = torch.Generator().manual_seed(B * 100 + SP)
    maps = torch.randn(B, C, T, Hm, Hm, generator=g)
    lab = cases.voronoi_labels(B, T, SP, Hm * scale, g, one_based=False)
    lab[:, :, :3, :5] = SP + 2
    lab[:, :, -2:, :] = -1
    md = maps.to(DEV).requires_grad_(True)
    out = ops.segment_mean(md, lab.to(DEV), SP)                                   # (B, SP, T, C)
    torch.testing.assert_close(out.detach().cpu().transpose(1, 2), O.segment_mean(maps, lab, SP), rtol=1e-5, atol=1e-6)
    gout = torch.randn(B, SP, T, C, generator=g)
    out.backward(gout.to(DEV))
    m2 = maps.clone().requires_grad_(True)
    up = m2.repeat_interleave(scale, -1).repeat_interleave(scale, -2)
    ok = (lab >= 0) & (lab < SP)
    oh = torch.nn.functional.one_hot(lab.clamp(0, SP - 1), SP).float() * ok[..., None]
    ref = torch.einsum("bcthw,bthws->btsc", up, oh) / (oh.sum((2, 3))[..., None] + 1e-20)
    ref.backward(gout.transpose(1, 2))
    torch.testing.assert_close(md.grad.cpu(), m2.grad, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", list(cases.LP_CASES))
def test_label_prop_matches_reference_golden(ops, name):
    from sapienza_video_contrastive_b200 import LabelPropagator, context_index_bank
    c = cases.LP_CASES[name]
    fx = load(name)
    feats, lbls = cases.lp_inputs(c)
    lp = LabelPropagator(c["n_ctx"], c["long_mem"], c["radius"], c["k"], c["tau"], normalize=False)
    preds, (Ws, Is) = lp(feats.to(DEV), lbls)
    ki = torch.cat(context_index_bank(c["n_ctx"], c["long_mem"], c["n_tgt"]), -1)
    check_topk_indices(feats, ki, Is.cpu(), fx["Is"], c)
    torch.testing.assert_close(Ws.cpu(), fx["Ws"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(preds.cpu(), fx["preds"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["lp_small", "lp_long2"])
@pytest.mark.parametrize("mask_kind", ["dense", "radius_object", "arbitrary_dense"])
def test_reference_signature_mem_efficient_batched_affinity(ops, name, mask_kind):
    """The reference's own call (test.py:112-129): materialised key bank + dense additive mask, lists of CPU tensors back."""
    from sapienza_video_contrastive_b200 import RadiusMask, context_index_bank, mem_efficient_batched_affinity
    c = cases.LP_CASES[name]
    fx = load(name)
    feats, lbls = cases.lp_inputs(c)
    ki = torch.cat(context_index_bank(c["n_ctx"], c["long_mem"], c["n_tgt"]), -1)
    keys, query = feats[:, :, ki].flatten(-2), feats[:, :, c["n_ctx"]:].flatten(-2)
    D = O.radius_mask_additive(c["h"], c["w"], c["radius"])
    if mask_kind == "radius_object":
        mask = RadiusMask(c["radius"], c["h"], c["w"])
    elif mask_kind == "arbitrary_dense":
        mask = D.clone()
        mask[0, 0, 0, 1] = -1e10                      # no longer a radius mask -> literal dense path
        mask[0, 0, 1, 0] = -1e10
    else:
        mask = D
    Ws, Is = mem_efficient_batched_affinity(query, keys, mask, c["tau"], c["k"], c["long_mem"], DEV, chunk=4)
    assert isinstance(Ws, list) and len(Ws) == c["n_tgt"] and not Ws[0].is_cuda and Is[0].dtype == torch.int64
    if mask_kind == "arbitrary_dense":
        Wo, Io = [], []
        f = feats[0].flatten(-2)
        for n in range(c["n_tgt"]):
            sc = torch.cat([f[:, ki[n, s]].t() @ f[:, n + c["n_ctx"]] + (mask[0, 0] if s >= len(c["long_mem"]) else 0)
                            for s in range(ki.shape[1])], 0) / c["tau"]
            v, i = torch.topk(sc, c["k"], dim=0)
            torch.testing.assert_close(torch.gather(sc, 0, Is[n]), v, rtol=0, atol=0)
    else:
        check_topk_indices(feats, ki, torch.stack(Is), fx["Is"], c)
        torch.testing.assert_close(torch.stack(Ws), fx["Ws"], rtol=1e-5, atol=1e-7)


def test_label_prop_davis_shape_properties(ops):
    """BASELINE config 4 shape (C=256, 60x107, 20 context frames + long memory, radius 12, k=10) on a short video:
    structural properties that do not need the oracle, plus an oracle spot check of two targets."""
    from sapienza_video_contrastive_b200 import LabelPropagator
    C, h, w, n_ctx, n_tgt, k, r = 256, 60, 107, 20, 3, 10, 12
    torch.manual_seed(0)
    feats = torch.nn.functional.normalize(torch.randn(1, C, n_ctx + n_tgt, h, w), dim=1)
    lp = LabelPropagator(n_ctx, [0], r, k, 0.07, normalize=True)
    ki, Ws, Is = lp.affinity(feats.to(DEV))
    hw = h * w
    assert Ws.shape == (n_tgt, k, hw) and Is.shape == (n_tgt, k, hw)
    torch.testing.assert_close(Ws.sum(1), torch.ones(n_tgt, hw, device=DEV), rtol=1e-5, atol=1e-5)
    assert bool((Ws[:, :-1] >= Ws[:, 1:]).all())
    assert int(Is.min()) >= 0 and int(Is.max()) < (n_ctx + 1) * hw
    slot, pos = Is // hw, Is % hw
    qpos = torch.arange(hw, device=DEV)[None, None]
    d2 = (pos // w - qpos // w) ** 2 + (pos % w - qpos % w) ** 2
    assert bool(((slot == 0) | (d2 < r * r)).all())                                  # radius respected outside long memory
    # oracle spot check on a band of queries of target 1
    f = feats[0].flatten(-2)
    for n in (1,):
        qs = torch.arange(3000, 3000 + 64)
        sc = torch.cat([f[:, ki[n, s].item()].t() @ f[:, n + n_ctx][:, qs] +
                        (O.radius_mask_additive(h, w, r)[0, 0][:, qs] if s >= 1 else 0) for s in range(ki.shape[1])], 0) / 0.07
        v, i = torch.topk(sc, k, dim=0)
        mine = torch.gather(sc, 0, Is[n][:, qs].cpu())
        torch.testing.assert_close(mine, v, rtol=1e-5, atol=1e-5)
        assert float((Is[n][:, qs].cpu() == i).float().mean()) > 0.97


def _lp_scores(feats, ki, n, n_ctx, h, w, radius, n_long, tau):
    f = feats[0].flatten(-2).double()
    add = O.radius_mask_additive(h, w, radius)[0, 0].double()
    return torch.cat([f[:, ki[n, s]].t() @ f[:, n + n_ctx] + (add if s >= n_long else 0) for s in range(ki.shape[1])], 0) / tau


@pytest.mark.parametrize("C,h,w,n_ctx,n_tgt,k,radius,dyadic", [(64, 20, 27, 3, 3, 10, 5, False), (128, 33, 19, 2, 2, 5, 12, False),
                                                              (256, 17, 40, 4, 2, 10, 12, False), (64, 18, 21, 2, 3, 10, 4, True),
                                                              (256, 24, 24, 2, 2, 16, 3, True)])
def test_label_prop_tensor_core_path(ops, C, h, w, n_ctx, n_tgt, k, radius, dyadic):
    """tcgen05 path (fp16 hi/lo split, fp32 accumulate) vs the exact-fp32 SIMT kernel and the float64 scores:
    identical index sets up to near-ties (|score difference| < 1e-5 / tau, the documented tolerance), exact on the
    dyadic grid where every product is exact in both."""
    from sapienza_video_contrastive_b200 import LabelPropagator, context_index_bank
    g = torch.Generator().manual_seed(C + h)
    Nf = n_ctx + n_tgt
    if dyadic:
        feats = torch.randint(-64, 65, (1, C, Nf, h, w), generator=g).float() / 64.0
    else:
        feats = torch.nn.functional.normalize(torch.randn(1, C, Nf, h, w, generator=g), dim=1)
    tau = 0.07
    res = {}
    for simt in (False, True):
        lp = LabelPropagator(n_ctx, [0], radius, k, tau, normalize=False, force_simt=simt)
        ki, Ws, Is = lp.affinity(feats.to(DEV))
        res[simt] = (Ws.cpu(), Is.cpu())
    ki = ki.cpu()
    hw = h * w
    (Wt, It), (Ws_, Is_) = res[False], res[True]
    assert int(It.min()) >= 0 and int(It.max()) < ki.shape[1] * hw
    for n in range(n_tgt):
        sc = _lp_scores(feats, ki, n, n_ctx, h, w, radius, 1, tau)
        vt, vs = torch.gather(sc, 0, It[n]), torch.gather(sc, 0, Is_[n])
        vref, iref = torch.topk(sc, k, dim=0)
        if dyadic:
            assert torch.equal(vt, vref) and torch.equal(vs, vref)           # exact arithmetic: identical score multisets
        else:
            assert float((vt - vref).abs().max()) < 1e-5 / tau
            assert float((vs - vref).abs().max()) < 1e-5 / tau
        srt = It[n].sort(0).values
        assert bool((srt[1:] != srt[:-1]).all())
        if n >= 1:
            assert float((It[n] == iref).float().mean()) > 0.995
    torch.testing.assert_close(Wt, Ws_, rtol=2e-4, atol=1e-6)


def test_head_linear_matches_nn_linear(ops):
    torch.manual_seed(0)
    lin = torch.nn.Linear(512, 128, bias=False).to(DEV)
    x = torch.randn(980, 4, 512, device=DEV, requires_grad=True)
    g = torch.randn(980, 4, 128, device=DEV)
    y0 = lin(x)
    y0.backward(g)
    gx0, gw0 = x.grad.clone(), lin.weight.grad.clone()
    x.grad = None
    lin.weight.grad = None
    y1 = ops.head_linear(x, lin.weight)
    y1.backward(g)
    torch.testing.assert_close(y1, y0, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(x.grad, gx0, rtol=1e-5, atol=1e-5)
    assert relmax(lin.weight.grad, gw0) < 1e-5
    assert int(ops.tc_error_word(DEV)[0]) == 0
    # against fp64: operands enter at full fp32 precision (tf32 big + small parts); what remains is the tensor core's own
    # fp32 accumulation, which truncates (measured 5e-6 of the largest output at K = 512; an fp32 sgemm gives ~5e-7)
    y64 = x.detach().double() @ lin.weight.detach().double().t()
    assert relmax(y1.detach().double(), y64) < 2e-5
    gw64 = g.double().reshape(-1, 128).t() @ x.detach().double().reshape(-1, 512)
    assert relmax(lin.weight.grad.double(), gw64) < 2e-5
    # a row count with no usable divisor takes the exact-fp32 SIMT split-K weight gradient
    x3 = torch.randn(977, 512, device=DEV, requires_grad=True)
    w3 = torch.randn(128, 512, device=DEV, requires_grad=True)
    g3 = torch.randn(977, 128, device=DEV)
    ops.head_linear(x3, w3).backward(g3)
    assert relmax(w3.grad.double(), g3.double().t() @ x3.detach().double()) < 2e-5
    assert relmax(x3.grad.double(), g3.double() @ w3.detach().double()) < 2e-5
    # a shape TMA cannot address (C not a multiple of 4) falls back to the library GEMM
    w2 = torch.randn(128, 510, device=DEV, requires_grad=True)
    x2 = torch.randn(77, 510, device=DEV, requires_grad=True)
    y2 = ops.head_linear(x2, w2)
    y2.sum().backward()
    torch.testing.assert_close(y2, x2 @ w2.t(), rtol=1e-5, atol=1e-5)


def test_async_wgrad_overlap_gives_identical_gradients(ops):
    """ops.set_async_wgrad(True) only changes scheduling (side stream + explicit join), never values."""
    torch.manual_seed(1)
    w = torch.randn(128, 512, device=DEV, requires_grad=True)
    maps = torch.randn(98, 4, 512, 8, 8, device=DEV)
    outs = []
    for flag in (False, True):
        ops.set_async_wgrad(flag)
        try:
            w.grad = None
            m = maps.clone().requires_grad_(True)
            f = ops.head_linear(ops.pool_patch(m), w).view(2, 49, 4, 128)
            torch.manual_seed(5)
            q, loss, xent, acc = ops.walk(f, 0.07, 0.1)
            loss.backward()
            ops.join_side_streams()
            torch.cuda.synchronize()
            outs.append((w.grad.clone(), m.grad.clone()))
        finally:
            ops.set_async_wgrad(False)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("name", list(cases.POST_CASES))
def test_label_postprocessing_matches_reference_golden(ops, name):
    """dump_predictions' resize + arg-max + colour table in one kernel, against the reference's own output."""
    pytest.importorskip("cv2")
    from sapienza_video_contrastive_b200 import test_utils as TU
    c = cases.POST_CASES[name]
    fx = load(name)
    pred, lbl_set, img = cases.post_inputs(c)
    cls, rgb = ops.lp_upsample_argmax(pred.to(DEV), (c["H"], c["W"]), lbl_set, norm_mask=c["norm_mask"])
    check_label_images(cls[0].cpu(), rgb[0].cpu(), pred, lbl_set, c, fx)
    blend, lbl, _ = TU.dump_predictions(pred.numpy(), lbl_set, img.numpy(), None, norm_mask=c["norm_mask"])        # the reference's signature
    assert lbl.shape == (c["H"], c["W"], 3) and lbl.dtype == torch.int32 and torch.equal(lbl.cpu().to(torch.uint8), rgb[0].cpu())
    torch.testing.assert_close(blend.cpu(), img * 0.5 + lbl.cpu().float() * 0.5, rtol=0, atol=1e-4)


def test_label_postprocessing_davis_shape_properties(ops):
    """DAVIS 480p shape (60x107 -> 480x854, 37 frames): one-hot maps reproduce nearest labels away from boundaries, the
    result is invariant to a positive rescaling of the maps, frames are independent."""
    torch.manual_seed(3)
    n, h, w, Lb, H, W = 37, 60, 107, 4, 480, 854
    hard = torch.randint(0, Lb, (n, h // 6, w // 6 + 1), device=DEV).repeat_interleave(6, 1).repeat_interleave(6, 2)[:, :h, :w]
    pred = torch.nn.functional.one_hot(hard, Lb).float()
    pal = torch.tensor([[0, 0, 0], [255, 0, 0], [0, 255, 0], [0, 0, 255]], device=DEV)
    cls, rgb = ops.lp_upsample_argmax(pred, (H, W), pal)
    assert cls.shape == (n, H, W) and rgb.shape == (n, H, W, 3)
    assert torch.equal(rgb, pal.to(torch.uint8)[cls.long()])
    # inside a 6x6 block of constant label (more than one source pixel from its border) the upsampled label is that label
    ys = ((torch.arange(H, device=DEV) + 0.5) * h / H - 0.5).round().clamp(0, h - 1).long()
    xs = ((torch.arange(W, device=DEV) + 0.5) * w / W - 0.5).round().clamp(0, w - 1).long()
    inner_y = ((ys % 6) >= 2) & ((ys % 6) <= 3)
    inner_x = ((xs % 6) >= 2) & ((xs % 6) <= 3)
    near = hard[:, ys][:, :, xs]
    m = inner_y[:, None] & inner_x[None, :]
    assert torch.equal(cls[:, m].long(), near[:, m])
    cls2, _ = ops.lp_upsample_argmax(pred * 3.0, (H, W), pal, want_rgb=False)
    assert torch.equal(cls2, cls)
    cls3, _ = ops.lp_upsample_argmax(pred[5:6], (H, W), pal, want_rgb=False)
    assert torch.equal(cls3[0], cls[5])
