"""The oracle (oracle/crw_oracle.py) is pinned here against outputs of the UNMODIFIED reference
(tests/golden/*.pt, produced by oracle/gen_golden.py in the authoring container)."""
import os

import pytest
import torch

from oracle import crw_oracle as O
from tests.golden import cases

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(G, name + ".pt"), weights_only=False)


def oracle_walk_from_maps(c, maps, head_w):
    maps = maps.clone().requires_grad_(True)
    head_w = head_w.clone().requires_grad_(True)
    q = O.patch_nodes(maps, head_w, c["B"])
    torch.manual_seed(c["seed"] + 1000)
    u12, u21p = O.draw_uniforms(c["B"], c["N"], c["T"]) if c["p"] > 0 else (None, None)
    loss, xents, accs, names = O.walk_loss(q, c["tau"], c["p"], u12, u21p, flip=c["flip"])
    loss.mean().backward()
    return q, loss, xents, accs, names, maps.grad, head_w.grad, (u12, u21p)


@pytest.mark.parametrize("name", list(cases.WALK_CASES))
def test_walk_matches_reference(name):
    c = cases.WALK_CASES[name]
    fx = load(name)
    maps, head_w = cases.walk_inputs(c)
    q, loss, xents, accs, names, gmaps, ghead, (u12, u21p) = oracle_walk_from_maps(c, maps, head_w)
    assert q.shape == fx["q"].shape
    torch.testing.assert_close(q, fx["q"], rtol=1e-5, atol=1e-6)
    assert loss.shape == fx["loss"].shape == (1,)
    torch.testing.assert_close(loss, fx["loss"], rtol=2e-6, atol=0)
    for n, xe, ac in zip(names, xents, accs):
        torch.testing.assert_close(xe, fx["diags"]["64 xent cyc %s" % n], rtol=2e-6, atol=0)
        torch.testing.assert_close(ac, fx["diags"]["64 acc cyc %s" % n], rtol=0, atol=1e-6)
    assert len(fx["diags"]) == 2 * len(names)
    torch.testing.assert_close(ghead, fx["grad_head"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(gmaps[..., 0, 0], fx["grad_maps00"], rtol=1e-4, atol=1e-8)
    # the transition matrices themselves, including the in-place union-mask artefact (F4) and the
    # transposed physical layout of the backward draws (F6)
    A12, A21 = O.walk_matrices(q.detach(), c["tau"], c["p"], u12, u21p)
    torch.testing.assert_close(torch.stack(A12), fx["A12"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(torch.stack(A21), fx["A21"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", list(cases.POSE_CASES))
def test_process_pose_matches_reference(name):
    """utils/test_utils.py:60-84 run unmodified: key-point coordinates (bit-equal: same float operations in the same order)
    and the sharp key-point image."""
    c = cases.POSE_CASES[name]
    fx = load(name)
    pred, lbl_set = cases.pose_inputs(c)
    coords, sharp = O.process_pose(pred, lbl_set.numpy())
    assert torch.equal(coords, fx["coords"])
    assert torch.equal(torch.from_numpy(sharp), fx["sharp"])
    for z in c["zero"]:
        assert coords[0, z - 1] == -1 and coords[1, z - 1] == -1


@pytest.mark.parametrize("name", list(cases.TS_CASES))
def test_teacher_student_loss_matches_reference(name):
    """CRWTeacherStudent.forward (teacherstudent.py:472-580) run unmodified from the node embeddings on: loss, walk diags and
    the gradient of the student's node vectors."""
    c = cases.TS_CASES[name]
    fx = load(name)
    fs, ft = cases.ts_inputs(c)
    fs = fs.clone().requires_grad_(True)
    nrm = lambda f: torch.nn.functional.normalize(f, p=2, dim=-1).permute(0, 3, 2, 1)          # (B,D,T,N)
    torch.manual_seed(c["seed"] + 1000)
    us12, us21p = O.draw_uniforms(c["B"], c["N"], c["T"])
    ut12, ut21p = O.draw_uniforms(c["B"], c["N"], c["T"])        # the teacher's matrices are dropped out too, with later draws
    loss, xents, ts = O.teacher_student_loss(nrm(fs), nrm(ft), c["tau"], c["p"], c["alpha"], us12, us21p, ut12, ut21p, flip=c["flip"])
    torch.testing.assert_close(loss, fx["loss"], rtol=1e-5, atol=0)
    names = [("l%d" if c["flip"] else "r%d") % i for i in range(1, c["T"] - 1)]
    for n, xe in zip(names, xents):
        torch.testing.assert_close(xe, fx["diags"]["%d xent cyc %s" % (8, n)], rtol=1e-5, atol=1e-7)
    loss.sum().backward()
    assert float((fs.grad - fx["grad_feats"]).abs().max() / fx["grad_feats"].abs().max()) < 1e-4


@pytest.mark.parametrize("name", list(cases.SPD_CASES))
def test_dilated_superpixel_matches_reference(name):
    """--dilate-superpixels (model.py:303-309): the oracle's run-free restatement against the reference's fp16 depthwise
    convolution, for the three structuring elements of utils/__init__.py:590-608."""
    c = cases.SPD_CASES[name]
    fx = load(name)
    maps, lab3, head_w = cases.sp_inputs(c)
    pooled = O.segment_mean_dilated(maps, lab3[:, :, 0], c["SP"], c["ksize"], c["shape"])       # (B,T,SP,C)
    f = pooled @ head_w.t()
    q = (f / f.norm(dim=-1, keepdim=True).clamp_min(1e-12)).permute(0, 3, 1, 2)
    torch.testing.assert_close(q, fx["sp_feats"], rtol=1e-4, atol=2e-6)
    if c["one_based"]:
        assert q[:, :, :, 0].abs().max() == 0
    # the element is symmetric, R wide on its centre row and a single pixel on its first and last rows
    hw = O.dilation_halfwidths(c["ksize"], c["shape"])
    R = c["ksize"] // 2
    assert len(hw) == c["ksize"] and hw[R] == R and hw == hw[::-1] and hw[0] == 0
    assert c["shape"] != "cross" or set(hw[:R]) == {0}


@pytest.mark.parametrize("name", list(cases.SP_CASES))
def test_superpixel_matches_reference(name):
    c = cases.SP_CASES[name]
    fx = load(name)
    maps, lab3, head_w = cases.sp_inputs(c)
    maps = maps.clone().requires_grad_(True)
    head_w = head_w.clone().requires_grad_(True)
    # segment_mean runs in float64 without autograd; rebuild it differentiably for the grad check
    q = O.superpixel_nodes(maps.detach(), lab3[:, :, 0], c["SP"], head_w.detach())
    torch.testing.assert_close(q, fx["sp_feats"], rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(q, fx["q"], rtol=1e-4, atol=2e-6)
    if c["one_based"]:
        assert q[:, :, :, 0].abs().max() == 0          # empty node -> zero embedding (F5)
    torch.manual_seed(c["seed"] + 1000)
    u12, u21p = O.draw_uniforms(c["B"], c["SP"], c["T"])
    loss, xents, accs, names = O.walk_loss(q, c["tau"], c["p"], u12, u21p)
    torch.testing.assert_close(loss, fx["loss"], rtol=1e-5, atol=0)
    for n, xe in zip(names, xents):
        torch.testing.assert_close(xe, fx["diags"]["256 xent cyc %s" % n], rtol=1e-5, atol=0)


ALL_LP = dict(cases.LP_CASES, **cases.LP_TC_CASES)


@pytest.mark.parametrize("name", list(ALL_LP))
def test_label_prop_matches_reference(name):
    c = ALL_LP[name]
    fx = load(name)
    feats, lbls = cases.lp_inputs(c)
    n_tgt = c["n_tgt"]
    ki = O.context_index_bank(c["n_ctx"], c["long_mem"], n_tgt)
    f = feats[0].flatten(-2)                                   # (C, Nf, hw)
    Ws, Is = O.lp_topk(f, ki, c["n_ctx"], len(c["long_mem"]), c["h"], c["w"], c["radius"], c["tau"], c["k"])
    assert Is.dtype == torch.int64 and Is.shape == fx["Is"].shape
    if not c["repeat_first"] and name != "lp_tc_dyadic":         # (exact ties: replicated frames, dyadic collisions at C=128)
        assert torch.equal(Is, fx["Is"])
    torch.testing.assert_close(Ws, fx["Ws"], rtol=1e-5, atol=1e-7)
    preds = O.lp_propagate(lbls, ki, Ws, Is, c["n_ctx"])
    torch.testing.assert_close(preds, fx["preds"], rtol=1e-5, atol=1e-6)


def test_label_prop_norm_mask_side_effect_matches_reference():
    """test.py:158-164: --norm_mask normalises the ground-truth frame 0 in place through the view `pred = lbls[0]`."""
    c = cases.LP_NORM_CASE
    fx = load("lp_normmask")
    feats, lbls = cases.lp_inputs(c)
    ki = O.context_index_bank(c["n_ctx"], c["long_mem"], c["n_tgt"])
    Ws, Is = O.lp_topk(feats[0].flatten(-2), ki, c["n_ctx"], 1, c["h"], c["w"], c["radius"], c["tau"], c["k"])
    torch.testing.assert_close(O.lp_propagate(lbls, ki, Ws, Is, c["n_ctx"], norm_mask=True), fx["preds"], rtol=1e-5, atol=1e-6)
    plain = O.lp_propagate(lbls, ki, Ws, Is, c["n_ctx"])
    plain = (plain - plain.min(-1, keepdim=True)[0])
    plain = plain / plain.max(-1, keepdim=True)[0]
    assert float((plain - fx["preds"]).abs().max()) > 1e-3       # normalising the outputs alone does not reproduce it


@pytest.mark.parametrize("name", list(cases.POST_CASES))
def test_label_postprocessing_matches_reference(name):
    """oracle.upsample_argmax == the reference's dump_predictions (utils/test_utils.py:85-123), label image bit for bit."""
    cv2 = pytest.importorskip("cv2")  # noqa: F841
    c = cases.POST_CASES[name]
    fx = load(name)
    pred, lbl_set, img = cases.post_inputs(c)
    cls, lbl, dist = O.upsample_argmax(pred, lbl_set, (c["H"], c["W"]), c["norm_mask"])
    assert torch.equal(lbl.to(torch.uint8), fx["pred_lbl"])
    assert cls.shape == (c["H"], c["W"]) and int(cls.max()) < c["L"]
    if "blend" in fx:
        torch.testing.assert_close(img * 0.5 + lbl.float() * 0.5, fx["blend"], rtol=0, atol=1e-4)


def test_misc_known_answers():
    fx = load("misc")
    torch.testing.assert_close(O.zero_softmax(fx["zs_in"]), fx["zs_out"], rtol=1e-6, atol=0)
    assert torch.equal(O.radius_mask_additive(5, 6, 3), fx["mask_5x6_r3"])
    for (nc, lm, N), bank in fx["banks"].items():
        assert torch.equal(O.context_index_bank(nc, list(lm), N), bank)
    # ZeroSoftmax artefacts the kernels must reproduce (F5): dropped edge -> numerator 1, zero logit -> 0
    x = torch.tensor([[-1e20 / 0.07, 0.0, 0.5 / 0.07]])
    e = (torch.exp(x) - 1) ** 2
    assert e[0, 0] == 1 and e[0, 1] == 0


def test_patch_grid_matches_reference_and_pil():
    """utils/augs.py:59-82 run unmodified (tests/golden/pg_160x128.pt) == oracle.patch_grid fed with the crop boxes drawn by the
    package's host mirror from the same seeds (pins the generator order), and the oracle's restatement of Pillow's BILINEAR
    resampling == PIL itself on random crops."""
    import numpy as np
    from PIL import Image
    from sapienza_video_contrastive_b200.augs import IMG_MEAN, IMG_STD, draw_patch_boxes
    c = cases.PG_CASE
    fx = load("pg_160x128")
    frame = cases.pg_frame(c)
    np.random.seed(c["np_seed"])
    torch.manual_seed(c["torch_seed"])
    assert np.random.random() * 0.0 + 0.5 == 0.5                     # the reference's stride draw (augs.py:60) consumes one number
    boxes = draw_patch_boxes(1, 12, 64)[0]
    out = O.patch_grid(frame.numpy(), boxes.numpy())
    mean, std = torch.tensor(IMG_MEAN)[:, None, None], torch.tensor(IMG_STD)[:, None, None]
    ref = ((fx["patches_u8"].float().div(255) - mean) / std).view(-1, 64, 64)
    assert torch.equal(out, ref)
    g = np.random.default_rng(1)
    img = (g.random((64, 64, 3)) * 255).astype(np.uint8)
    for (i, j, h, w) in [(5, 0, 53, 62), (0, 3, 61, 50), (2, 2, 60, 60), (0, 0, 64, 64), (10, 12, 47, 52)]:
        pil = np.asarray(Image.fromarray(img).crop((j, i, j + w, i + h)).resize((64, 64), Image.BILINEAR))
        mine = O.patch_grid(img, np.array([[i, j, h, w]]), 64, 32, 64, (0, 0, 0), (1, 1, 1))
        assert torch.equal(mine, torch.from_numpy(pil.copy()).permute(2, 0, 1).float().div(255))
