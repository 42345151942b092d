"""CPU-side checks: the C ABI library builds, loads and exports every symbol of include/crw_b200.h; the host logic
(index banks, radius masks, argument checks) matches the reference's golden vectors; the product never touches the
oracle and has no CPU path."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sapienza_video_contrastive_b200")


def header_symbols():
    src = open(os.path.join(ROOT, "include", "crw_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crw_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_header_symbols():
    from sapienza_video_contrastive_b200 import _lib
    path = _lib.build()
    lib = _lib.CrwLib(path)                                   # no compute calls: loading + symbol lookup only
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib._dll, s), s
    assert sorted(_lib.EXPORTS) == syms                       # the ctypes table and the header agree
    assert lib.crw_version() >= 100
    assert lib.crw_walk_workspace_bytes(20, 49, 4, 128, 0) > 0
    assert lib.crw_walk_workspace_bytes(20, 49, 4, 128, _lib.WALK_FORCE_GENERAL) > 0
    assert lib.crw_segmean_workspace_bytes(2, 3, 32, 32, 256, 256, 100) > 0
    assert lib.crw_segmean_dilated_workspace_bytes(2, 3, 32, 32, 256, 256, 100) > lib.crw_segmean_workspace_bytes(2, 3, 32, 32, 256, 256, 100)
    assert lib.crw_segmean_dilated_workspace_bytes(2, 3, 32, 32, 256, 256, 256) == 0          # SP <= 255 in the dilated path


def test_flag_constants_match_the_header():
    """The Python copies of the CRW_WALK_* / CRW_LP_* flag bits are the header's; workspace queries honour them."""
    from sapienza_video_contrastive_b200 import _lib
    src = open(os.path.join(ROOT, "include", "crw_b200.h")).read()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+CRW_(WALK_[A-Z0-9_]+)\s+(\d+)u", src)}
    assert set(flags) >= {"WALK_SOFTMAX", "WALK_FLIP", "WALK_FORCE_GENERAL", "WALK_FORCE_SIMT", "WALK_FORCE_TC", "WALK_NO_CLUSTER", "WALK_NO_TF32"}
    for name, val in flags.items():
        assert getattr(_lib, name) == val, name
    assert len(set(flags.values())) == len(flags)                       # distinct bits
    lib = _lib.CrwLib(_lib.build())
    # the general path reserves tensor-core operand planes from N = 64 on, unless the SIMT GEMM is forced
    big = lib.crw_walk_workspace_bytes(2, 128, 4, 128, 0)
    simt = lib.crw_walk_workspace_bytes(2, 128, 4, 128, _lib.WALK_FORCE_SIMT)
    assert big > simt > 0
    assert lib.crw_walk_workspace_bytes(2, 60, 16, 128, 0) == lib.crw_walk_workspace_bytes(2, 60, 16, 128, _lib.WALK_FORCE_SIMT)
    assert lib.crw_head_wgrad_workspace_bytes(3920, 128, 512) >= 4 * 35 * 128 * 512
    assert lib.crw_bmm_tc_workspace_bytes(2, 128, 128, 128) > 0


def test_product_never_imports_the_oracle_or_a_cpu_path():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("crw_oracle", "oracle") or f == "__none__", f
                assert "cusim" not in txt or f.endswith((".cu", ".cuh")), f


def test_ops_refuse_cpu_tensors():
    from sapienza_video_contrastive_b200 import ops
    with pytest.raises(RuntimeError):
        ops.pool_patch(torch.zeros(4, 2, 8, 8))
    with pytest.raises(RuntimeError):
        ops.walk(torch.zeros(1, 4, 3, 8), 0.07, 0.0)
    with pytest.raises(RuntimeError):
        ops.lp_prepare(torch.zeros(4, 2, 6), True)


def test_host_index_math_matches_reference_goldens():
    from sapienza_video_contrastive_b200.test_utils import MaskedAttention, context_index_bank
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "misc.pt"), weights_only=False)
    for (nc, lm, N), bank in fx["banks"].items():
        assert torch.equal(torch.cat(context_index_bank(nc, list(lm), N), dim=-1), bank)
    D = MaskedAttention(3, flat=False).mask(5, 6)[None].clone()
    D = D.flatten(-4, -3).flatten(-2)
    D[D == 0] = -1e10
    D[D == 1] = 0
    assert torch.equal(D, fx["mask_5x6_r3"])
    with pytest.raises(AssertionError):
        context_index_bank(3, [7], 5)


def test_crw_module_surface_on_cpu():
    """Construction is plain PyTorch (stock encoder + head): names, shapes and defaults match the reference's."""
    import argparse
    from sapienza_video_contrastive_b200 import CRW
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "cfg1_resnet18.pt"), weights_only=False)
    torch.manual_seed(0)
    args = argparse.Namespace(device="cpu", dropout=0.1, featdrop=0.0, temp=0.07, head_depth=0, model_type="scratch",
                              remove_layers=[], dilate_superpixels=False, flip=False, sk_targets=False)
    crw = CRW(args)
    assert list(crw.state_dict().keys()) == fx["state_dict_keys"]
    assert abs(float(sum(p.double().sum() for p in crw.parameters())) - fx["param_checksum"]) < 1e-6   # same init stream
    assert crw.enc_hid_dim == 512 and crw.map_scale == 8 and crw.temperature == 0.07 and crw.edgedrop_rate == 0.1
    assert torch.equal(crw.xent_targets(torch.zeros(2, 5, 5)), torch.arange(5).repeat(2))
    with pytest.raises(RuntimeError):
        crw(torch.zeros(1, 4, 6, 64, 64))                       # CPU input: refused, no fallback
    assert crw.dilation is None
    dil = CRW(argparse.Namespace(**{**vars(args), "dilate_superpixels": True}))
    assert dil.dilation == (51, "L1")                             # utils/arguments.py:209-210 defaults
    with pytest.raises(AssertionError):                          # utils/__init__.py:591
        CRW(argparse.Namespace(**{**vars(args), "dilate_superpixels": True, "dilation_kernel_size": 50}))
    with pytest.raises(ValueError):
        CRW(argparse.Namespace(**{**vars(args), "dilate_superpixels": True, "dilation_kernel_shape": "square"}))


def test_teacher_student_module_surface_on_cpu(tmp_path):
    """teacherstudent.py:11-53, 294-341: biased head, softmax walk, frozen teacher loaded from args.path_to_pretrained,
    the reference's state-dict names, alpha checked; no CPU forward."""
    import argparse
    from sapienza_video_contrastive_b200 import CRWBase, CRWTeacherStudent, SoftCrossEntropyLoss
    torch.manual_seed(0)
    args = argparse.Namespace(device="cpu", dropout=0.1, featdrop=0.0, temp=0.07, head_depth=0, model_type="scratch",
                              remove_layers=[], dilate_superpixels=False, flip=False, sk_targets=False,
                              alpha_teacher_student=0.5, path_to_pretrained=str(tmp_path / "pretrained.pth"))
    base = CRWBase(args)
    assert base.use_softmax and base.selfsim_fc[0].bias is not None
    assert "selfsim_fc.0.bias" in base.state_dict()
    torch.save({"model": base.state_dict()}, args.path_to_pretrained)
    ts = CRWTeacherStudent(args)
    assert all(not p.requires_grad for p in ts.teacher.parameters()) and any(p.requires_grad for p in ts.parameters())
    for k, v in base.state_dict().items():
        assert torch.equal(ts.state_dict()["teacher." + k], v)
    assert {k.split(".")[0] for k in ts.state_dict()} == {"encoder", "selfsim_fc", "teacher"}
    with pytest.raises(AssertionError):
        CRWTeacherStudent(argparse.Namespace(**{**vars(args), "alpha_teacher_student": 1.5}), teacher=base)
    with pytest.raises(RuntimeError):
        ts(torch.zeros(1, 4, 6, 64, 64))
    # SoftCrossEntropyLoss (:270-292): hard targets reduce to the usual cross-entropy
    x = torch.randn(5, 7)
    y = torch.randint(0, 7, (5,))
    onehot = torch.nn.functional.one_hot(y, 7).float()
    torch.testing.assert_close(SoftCrossEntropyLoss()(x, onehot), torch.nn.functional.cross_entropy(x, y))
    assert SoftCrossEntropyLoss(reduction="none")(x, onehot).shape == (5,)
    with pytest.raises(ValueError):
        SoftCrossEntropyLoss(reduction="median")(x, onehot)


def test_davis_index_maps_matches_convert_davis():
    """eval/convert_davis.py:52-66 restated with numpy + cv2 on a synthetic label image: colour -> palette id, unknown
    colours -> 0, INTER_NEAREST resize to the ground-truth size (up and down, non-integer ratios)."""
    cv2 = pytest.importorskip("cv2")
    import numpy as np
    from sapienza_video_contrastive_b200.test_utils import davis_index_maps
    g = torch.Generator().manual_seed(0)
    palette = torch.randint(0, 256, (40, 3), generator=g)
    palette[0] = 0
    lbl_set = torch.stack([palette[0], palette[7], palette[21], torch.tensor([1, 2, 3]), palette[39]])   # one colour off-palette
    cls = torch.randint(0, 5, (2, 37, 53), generator=g).to(torch.uint8)
    pal_np = palette.numpy().astype(np.uint8)
    for size in (None, (37, 53), (60, 107), (19, 30), (480, 854)):
        got = davis_index_maps(cls, lbl_set, palette, size)
        for f in range(2):
            lblimg = lbl_set.numpy().astype(np.uint8)[cls[f].numpy()]                # what dump_predictions wrote / imread returns
            idx = np.zeros(lblimg.shape[:2])
            for c in np.unique(lblimg.reshape(-1, 3), axis=0):
                cid = np.arange(0, pal_np.shape[0])[np.all(pal_np == c, axis=-1)]
                if len(cid) > 0:
                    idx[np.all(lblimg == c, axis=-1)] = cid
            idx = idx.astype(np.uint8)
            if size is not None:
                idx = cv2.resize(idx, (size[1], size[0]), interpolation=cv2.INTER_NEAREST)
            assert np.array_equal(got[f].numpy(), idx), size


def test_small_test_utils_mirrors():
    """hard_prop (utils/test_utils.py:51-56) against the reference's three in-place statements; infer_downscale (:212-216)."""
    from sapienza_video_contrastive_b200.test_utils import hard_prop, infer_downscale
    g = torch.Generator().manual_seed(2)
    pred = torch.rand(4, 5, 6, generator=g)
    pred[1, 0, 0] = pred[2, 0, 0] = 2.0                           # a tie shares the mass
    ref = pred.clone()
    mx = ref.max(axis=0)[0]
    ref[ref < mx] = 0
    ref[ref >= mx] = 1
    ref /= ref.sum(0)[None]
    out = hard_prop(pred)
    assert out is pred and torch.equal(pred, ref) and float(pred[1, 0, 0]) == 0.5
    neg = -torch.rand(3, 2, 2, generator=g) - 0.5                # negative maxima: the zeroed entries pass the second test too
    ref = neg.clone()
    mx = ref.max(axis=0)[0]
    ref[ref < mx] = 0
    ref[ref >= mx] = 1
    ref /= ref.sum(0)[None]
    assert torch.equal(hard_prop(neg), ref)
    assert list(infer_downscale()) == [8, 8]


def test_library_staleness_is_by_content_not_by_mtime(tmp_path):
    """The built library travels to other machines by copy (modification times arrive in any order) and is imported by all ranks of
    a torchrun job at once: `_lib` decides by a digest of the sources written next to the library, and builds under a file lock."""
    import time
    from sapienza_video_contrastive_b200 import _lib
    _lib.build()
    assert not _lib._stale()
    src = _lib.sources()[0]
    st = os.stat(src)
    try:
        os.utime(src, (time.time() + 3600, time.time() + 3600))          # "newer" than the library: still current
        assert not _lib._stale()
    finally:
        os.utime(src, (st.st_atime, st.st_mtime))
    with open(_lib.HASH_PATH) as f:
        good = f.read()
    try:
        with open(_lib.HASH_PATH, "w") as f:
            f.write("0" * 40 + "\n")
        assert _lib._stale()                                            # a library built from other sources is stale
    finally:
        with open(_lib.HASH_PATH, "w") as f:
            f.write(good)
    assert not _lib._stale()
