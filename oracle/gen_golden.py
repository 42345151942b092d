"""Generate tests/golden/*.pt by EXECUTING the unmodified reference (authoring container only).

    python -m oracle.gen_golden            # rewrites every fixture

Each fixture stores only the reference's outputs; inputs are regenerated from seeds by
tests/golden/cases.py.  TEST INFRASTRUCTURE - never imported by the product package.
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.golden import cases            # noqa: E402
from oracle import ref_import             # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class FakeEnc(nn.Module):
    """Stands in for the ResNet so the fixture exercises only the hot path: returns the prepared
    feature maps as a leaf.  Owns a parameter (model.py:42) and answers the 256-px probe of
    infer_dims with a 32x32 map so map_scale == 8 (model.py:40-45)."""

    def __init__(self, ce):
        super().__init__()
        self.dummy = nn.Parameter(torch.zeros(1))
        self.ce = ce
        self.maps = None

    def forward(self, x):
        if self.maps is None:
            return torch.zeros(1, self.ce, 1, 32, 32)
        return self.maps


def make_crw(ref_model, ref_utils, ce, head_w, **ns):
    enc = FakeEnc(ce)
    orig = ref_utils.make_encoder
    ref_utils.make_encoder = lambda a: enc
    try:
        crw = ref_model.CRW(ref_import.namespace(**ns))
    finally:
        ref_utils.make_encoder = orig
    with torch.no_grad():
        crw.selfsim_fc[0].weight.copy_(head_w)
    return crw, enc


def record_stoch(crw):
    rec = []
    orig = crw.stoch_mat

    def wrapped(A, *a, **k):
        out = orig(A, *a, **k)
        rec.append(out.detach().clone())
        return out

    crw.stoch_mat = wrapped
    return rec


def gen_walk(ref_model, ref_utils):
    for name, c in cases.WALK_CASES.items():
        maps, head_w = cases.walk_inputs(c)
        crw, enc = make_crw(ref_model, ref_utils, c["Ce"], head_w, dropout=c["p"], temp=c["tau"], flip=c["flip"])
        rec = record_stoch(crw)
        enc.maps = maps.clone().requires_grad_(True)
        x = torch.zeros(c["B"], c["T"], c["N"] * 3, 64, 64)
        torch.manual_seed(c["seed"] + 1000)
        q, loss, diags = crw(x, None, None)
        loss.mean().backward()
        gm = enc.maps.grad
        assert (gm - gm[..., :1, :1]).abs().max() == 0          # mean-pool grad is a spatial broadcast
        T = c["T"]
        fx = dict(q=q.detach().clone(), loss=loss.detach().clone(),
                  diags={k: v.detach().clone() for k, v in diags.items()},
                  grad_head=crw.selfsim_fc[0].weight.grad.clone(), grad_maps00=gm[..., 0, 0].clone(),
                  A12=torch.stack(rec[: T - 1]), A21=torch.stack(rec[T - 1:]))
        torch.save(fx, os.path.join(OUT, name + ".pt"))
        print(name, float(loss), {k: round(float(v), 5) for k, v in diags.items()})


def gen_cfg1(ref_model, ref_utils):
    """BASELINE config 1 verbatim: random ResNet-18 'scratch' under seed 0, x ~ N(0,1) (2,4,147,64,64),
    seed 123 right before the call."""
    torch.manual_seed(0)
    crw = ref_model.CRW(ref_import.namespace(dropout=0.1, temp=0.07))
    x = torch.randn(2, 4, 147, 64, 64)
    torch.manual_seed(123)
    q, loss, diags = crw(x, None, None)
    loss.mean().backward()
    sd_keys = list(crw.state_dict().keys())
    fx = dict(q=q.detach().clone(), loss=loss.detach().clone(),
              diags={k: v.detach().clone() for k, v in diags.items()},
              grad_head=crw.selfsim_fc[0].weight.grad.clone(),
              grad_conv1=crw.encoder.model.conv1.weight.grad.clone(),
              state_dict_keys=sd_keys,
              param_checksum=float(sum(p.double().sum() for p in crw.parameters())))
    torch.save(fx, os.path.join(OUT, "cfg1_resnet18.pt"))
    print("cfg1", float(loss), {k: round(float(v), 5) for k, v in diags.items()})


def gen_sp(ref_model, ref_utils):
    for name, c in cases.SP_CASES.items():
        maps, lab3, head_w = cases.sp_inputs(c)
        crw, enc = make_crw(ref_model, ref_utils, c["Ce"], head_w, dropout=c["p"], temp=c["tau"])
        enc.maps = maps.clone().requires_grad_(True)
        x = torch.zeros(c["B"], c["T"], 3, 256, 256)
        # nodes only
        with torch.no_grad():
            sp_feats, _ = crw.image_to_nodes(x, lab3, c["SP"])
        torch.manual_seed(c["seed"] + 1000)
        q, loss, diags = crw(x, lab3, c["SP"])
        loss.mean().backward()
        fx = dict(sp_feats=sp_feats.clone(), q=q.detach().clone(), loss=loss.detach().clone(),
                  diags={k: v.detach().clone() for k, v in diags.items()},
                  grad_head=crw.selfsim_fc[0].weight.grad.clone(), grad_maps=enc.maps.grad.clone())
        torch.save(fx, os.path.join(OUT, name + ".pt"))
        print(name, float(loss), tuple(sp_feats.shape))


def gen_spd(ref_model, ref_utils):
    """CRW.image_to_nodes with --dilate-superpixels (model.py:303-309), all three kernel shapes, plus the gradient of a
    fixed linear functional of the node embeddings with respect to the maps and the head."""
    for name, c in cases.SPD_CASES.items():
        maps, lab3, head_w = cases.sp_inputs(c)
        crw, enc = make_crw(ref_model, ref_utils, c["Ce"], head_w, dropout=0.0, temp=0.07, dilate_superpixels=True,
                            dilation_kernel_size=c["ksize"], dilation_kernel_shape=c["shape"])
        enc.maps = maps.clone().requires_grad_(True)
        x = torch.zeros(c["B"], c["T"], 3, 256, 256)
        sp_feats, _ = crw.image_to_nodes(x, lab3, c["SP"])
        g = torch.Generator().manual_seed(c["seed"] + 7)
        proj = torch.randn(sp_feats.shape, generator=g)
        (sp_feats * proj).sum().backward()
        fx = dict(sp_feats=sp_feats.detach().clone(), grad_maps=enc.maps.grad.clone(), grad_head=crw.selfsim_fc[0].weight.grad.clone())
        torch.save(fx, os.path.join(OUT, name + ".pt"))
        print("golden", name, tuple(sp_feats.shape))


def gen_ts(ref_model, ref_utils):
    """CRWTeacherStudent.forward (teacherstudent.py:472-580) from the node embeddings on.  The class needs a pretrained
    checkpoint to construct, so an instance is made without __init__ and given exactly the attributes forward reads; the two
    pixels_to_nodes methods return fixed unit-norm embeddings (B,C,T,N)."""
    import teacherstudent as ts_mod
    import torch.nn.functional as F
    for name, c in cases.TS_CASES.items():
        fs, ft = cases.ts_inputs(c)
        obj = ts_mod.CRWTeacherStudent.__new__(ts_mod.CRWTeacherStudent)
        nn.Module.__init__(obj)
        obj.args = argparse.Namespace(device="cpu")
        obj.edgedrop_rate, obj.temperature, obj.flip, obj.alpha, obj.vis = c["p"], c["tau"], c["flip"], c["alpha"], None
        obj._xent_targets = dict()
        obj.xent = nn.CrossEntropyLoss(reduction="none")
        obj.soft_xent = ts_mod.SoftCrossEntropyLoss()
        fs_req = fs.clone().requires_grad_(True)
        obj.pixels_to_nodes = lambda x: (F.normalize(fs_req, p=2, dim=-1).permute(0, 3, 2, 1), None)
        obj.pixels_to_nodes_tchr = lambda x: (F.normalize(ft, p=2, dim=-1).permute(0, 3, 2, 1), None)
        x = torch.zeros(c["B"], c["T"], c["N"] * 3, 8, 8)
        torch.manual_seed(c["seed"] + 1000)
        q, loss, diags = obj(x)
        loss.sum().backward()
        fx = dict(q=q.detach().clone(), loss=loss.detach().clone(), diags={k: v.detach().clone() for k, v in diags.items()},
                  grad_feats=fs_req.grad.clone())
        torch.save(fx, os.path.join(OUT, name + ".pt"))
        print("golden", name, float(loss))


def gen_pose(ref_model, ref_utils, ref_tu):
    """utils/test_utils.py:60-84 (process_pose), the JHMDB branch of test.py:171-172."""
    for name, c in cases.POSE_CASES.items():
        pred, lbl_set = cases.pose_inputs(c)
        coords, sharp = ref_tu.process_pose(pred.clone(), lbl_set.numpy())
        torch.save(dict(coords=coords.clone(), sharp=torch.from_numpy(sharp)), os.path.join(OUT, name + ".pt"))
        print("golden", name, tuple(coords.shape))


def gen_lp(ref_model, ref_utils, ref_tu, only=None):
    """Runs the reference's own evaluator loop (test.py:67-160) on a fake loader / fake encoder and
    captures (Ws, Is) from mem_efficient_batched_affinity and each `pred` handed to dump_predictions."""
    import test as ref_test
    ref_test.vis = None          # test.py:201 reads an undefined module global; supply it

    for name, c in list(cases.LP_CASES.items()) + list(cases.LP_TC_CASES.items()) + [("lp_normmask", cases.LP_NORM_CASE)]:
        if only is not None and name not in only:
            continue
        feats, lbls = cases.lp_inputs(c)
        Nf = feats.shape[2]

        class Model:
            def encoder(self, x):                       # test.py:87 - called on 5-frame chunks
                b0 = int(x[0, 0, :, 0, 0].tolist()[0])
                return feats[:, :, b0:b0 + x.shape[2]]

        # imgs only carries the frame index so the fake encoder can slice the prepared features
        imgs = torch.arange(Nf).float()[None, :, None, None, None].repeat(1, 1, 3, 2, 2)
        imgs_orig = torch.zeros(1, Nf, 3, 4, 4)
        lbl_map = torch.zeros(1, c["L"], 3)
        loader = [(imgs, imgs_orig, lbls[None].clone(), None, lbl_map, {})]
        captured = {}
        orig_aff = ref_tu.mem_efficient_batched_affinity
        orig_dump = ref_tu.dump_predictions

        def aff(*a, **k):
            Ws, Is = orig_aff(*a, **k)
            captured["Ws"] = torch.stack(Ws)
            captured["Is"] = torch.stack(Is)
            return Ws, Is

        preds = []

        def dump(pred, *a, **k):
            preds.append(torch.from_numpy(pred).clone())
            return None, None, None

        ref_tu.mem_efficient_batched_affinity = aff
        ref_tu.dump_predictions = dump
        args = argparse.Namespace(videoLen=c["n_ctx"], long_mem=c["long_mem"], radius=c["radius"],
                                  temperature=c["tau"], topk=c["k"], device="cpu", no_l2=True, pca_vis=False,
                                  norm_mask=(name == "lp_normmask"), filelist="davis", save_path=tempfile.mkdtemp(), visdom=False)
        try:
            with torch.no_grad():
                ref_test.test(loader, Model(), args)
        finally:
            ref_tu.mem_efficient_batched_affinity = orig_aff
            ref_tu.dump_predictions = orig_dump
        fx = dict(Ws=captured["Ws"], Is=captured["Is"], preds=torch.stack(preds))
        torch.save(fx, os.path.join(OUT, name + ".pt"))
        print(name, tuple(fx["Ws"].shape), tuple(fx["preds"].shape))


def gen_post(ref_model, ref_utils, ref_tu):
    """The reference's own dump_predictions (utils/test_utils.py:85-123) on seeded soft label maps; only its file writes and
    the matplotlib colour map (both outside the path) are replaced.  --norm_mask is applied as test.py:162-164 does."""
    import types
    import numpy as np
    orig_io, orig_cm = ref_tu.imageio, ref_tu.cm
    ref_tu.imageio = types.SimpleNamespace(imwrite=lambda *a, **k: None)
    ref_tu.cm = types.SimpleNamespace(jet=lambda x: np.zeros(x.shape + (4,), dtype=np.float32))
    try:
        for name, c in cases.POST_CASES.items():
            pred, lbl_set, img = cases.post_inputs(c)
            pred = pred.clone()
            if c["norm_mask"]:
                pred[:, :, :] -= pred.min(-1)[0][:, :, None]
                pred[:, :, :] /= pred.max(-1)[0][:, :, None]
            blend, lbl, _ = ref_tu.dump_predictions(pred.cpu().numpy(), lbl_set, img.numpy(), os.path.join(tempfile.gettempdir(), "x.jpg"))
            fx = {"pred_lbl": torch.from_numpy(np.asarray(lbl)).to(torch.uint8)}        # colours are 0..255: bytes keep the fixture small
            if name == "post_shrink":
                fx["blend"] = torch.from_numpy(np.asarray(blend))
            torch.save(fx, os.path.join(OUT, name + ".pt"))
            print("golden", name, tuple(fx["pred_lbl"].shape))
    finally:
        ref_tu.imageio, ref_tu.cm = orig_io, orig_cm


def gen_pg(ref_model, ref_utils):
    """The reference's own patch_grid (utils/augs.py:59-82) on a seeded frame.  skimage is absent here: its view_as_windows
    (pure index math) is supplied by numpy's sliding_window_view with the same window shape and step.  The fixture stores the
    8-bit patches (the float output is an exact function of them, checked below)."""
    import numpy as np
    from torchvision import transforms
    augs = ref_utils.augs

    def view_as_windows(x, shape, step):
        v = np.lib.stride_tricks.sliding_window_view(x, tuple(int(s) for s in shape))
        return v[::step[0], ::step[1], ::step[2]]

    augs.skimage.util.view_as_windows = view_as_windows
    c = cases.PG_CASE
    frame = cases.pg_frame(c).numpy()
    np.random.seed(c["np_seed"])
    torch.manual_seed(c["torch_seed"])
    aug = augs.patch_grid(transforms.Compose(augs.NORM), shape=np.array([64, 64, 3]))
    out = aug(frame)                                                               # (P*3, 64, 64) fp32
    mean = torch.tensor(augs.IMG_MEAN)[:, None, None]
    std = torch.tensor(augs.IMG_STD)[:, None, None]
    P = out.shape[0] // 3
    u8 = torch.round((out.view(P, 3, 64, 64) * std + mean) * 255).to(torch.uint8)
    back = ((u8.float().div(255) - mean) / std).view(-1, 64, 64)
    assert torch.equal(back, out), "the float patches must be an exact function of the 8-bit ones"
    torch.save(dict(patches_u8=u8), os.path.join(OUT, "pg_160x128.pt"))
    print("golden pg", tuple(u8.shape))


def gen_misc(ref_model, ref_utils, ref_tu):
    """Small known-answer vectors: ZeroSoftmax, radius mask, context_index_bank, affinity, stoch_mat."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 7, generator=g) * 3
    x[0, 0] = -1e20 / 0.07
    x[1, 1] = 0.0
    zs = ref_utils.ZeroSoftmax()(x, dim=-1)
    D = ref_utils.MaskedAttention(3, flat=False).mask(5, 6)[None].clone()
    D = D.flatten(-4, -3).flatten(-2)
    D[D == 0] = -1e10
    D[D == 1] = 0
    banks = {}
    for nc, lm, N in [(4, [0], 6), (3, [0, 2], 8), (2, [], 5), (20, [0], 50)]:
        banks[(nc, tuple(lm), N)] = torch.cat(ref_tu.context_index_bank(nc, lm, N), dim=-1)
    fx = dict(zs_in=x, zs_out=zs, mask_5x6_r3=D, banks=banks)
    torch.save(fx, os.path.join(OUT, "misc.pt"))
    print("misc ok")


def main():
    ref_model, ref_utils, ref_tu = ref_import.load()
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "post":           # only the post-processing fixtures (added later)
        gen_post(ref_model, ref_utils, ref_tu)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "spd":            # only the dilated-superpixel fixtures (added later)
        gen_spd(ref_model, ref_utils)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "pose":           # only the key-point fixtures (added later)
        gen_pose(ref_model, ref_utils, ref_tu)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "lp_tc":          # only the tensor-core-shaped label-propagation fixtures (round 2)
        gen_lp(ref_model, ref_utils, ref_tu, only=set(cases.LP_TC_CASES) | {"lp_normmask"})
        return
    if len(sys.argv) > 1 and sys.argv[1] == "pg":             # only the patch-grid fixture (round 2)
        gen_pg(ref_model, ref_utils)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "ts":             # only the teacher-student fixtures (added later)
        gen_ts(ref_model, ref_utils)
        return
    gen_misc(ref_model, ref_utils, ref_tu)
    gen_post(ref_model, ref_utils, ref_tu)
    gen_walk(ref_model, ref_utils)
    gen_sp(ref_model, ref_utils)
    gen_spd(ref_model, ref_utils)
    gen_ts(ref_model, ref_utils)
    gen_pose(ref_model, ref_utils, ref_tu)
    gen_lp(ref_model, ref_utils, ref_tu)
    gen_pg(ref_model, ref_utils)
    gen_cfg1(ref_model, ref_utils)


if __name__ == "__main__":
    main()
