"""Superpixel label-map producer (csrc/slic.cu, SURVEY 8f rank 4 second half).

scikit-image is not in this image, so the oracle (oracle/slic_oracle.py) restates its published SLIC and parity with
scikit-image itself is UNPINNED.  What these tests pin: OpenCV's min-max normalisation (against cv2 here), the regular grid
(hand-computed values), structural invariants of the labelling, and - the parity test proper - the kernels against the oracle
bit for bit, on the host simulator here and on the GPU under -m gpu.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import slic_oracle as SO


def smooth_video(F, H, W, seed, sigma=4.0):
    """(F, 3, H, W) float32: blurred noise, so that segments are blobs rather than salt and pepper."""
    import cv2
    rng = np.random.default_rng(seed)
    fr = [cv2.GaussianBlur(rng.standard_normal((H, W, 3)).astype(np.float32), (0, 0), sigma) * rng.uniform(0.5, 3) + rng.uniform(-1, 1)
          for _ in range(F)]
    return np.ascontiguousarray(np.stack([np.moveaxis(f, -1, 0) for f in fr]).astype(np.float32))


# ---- oracle pins ------------------------------------------------------------------------------------------------------------

def test_normalize_matches_opencv():
    import cv2
    rng = np.random.default_rng(3)
    for shape, k in (((64, 48, 3), 1.0), ((120, 160, 3), 3.0), ((512, 512, 3), 0.2)):
        img = (rng.standard_normal(shape) * k + rng.uniform(-2, 2)).astype(np.float32)
        ref = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8U)
        assert np.array_equal(SO.normalize_minmax_u8(img), ref)
    flat = np.full((8, 8, 3), 0.25, dtype=np.float32)
    assert np.array_equal(SO.normalize_minmax_u8(flat), cv2.normalize(flat, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8U))


def test_regular_grid_hand_values():
    # 256 x 256, 100 points: step sqrt(65536 / 100) = 25.6 -> start 12, step 26 (10 x 10 centres)
    assert SO.regular_grid_2d(256, 256, 100) == (12, 26, 12, 26)
    # 30 points: step 46.7 -> start 23, step 47 (5 x 5 = 25 centres, fewer than asked, as scikit-image)
    assert SO.regular_grid_2d(256, 256, 30) == (23, 47, 23, 47)
    # more points than pixels: every pixel a centre
    assert SO.regular_grid_2d(4, 4, 100) == (0, 1, 0, 1)
    # non-square frames share one step
    assert SO.regular_grid_2d(64, 48, 20) == (6, 12, 6, 12)


def test_lab_known_colours():
    u8 = np.array([[[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255]]], dtype=np.uint8)
    lab = SO.rgb8_to_lab(u8)[0]
    # CIE Lab (D65) of black, white and the sRGB primaries (scikit-image's rgb2lab documentation values, 2 decimals)
    want = np.array([[0, 0, 0], [100, 0, 0], [53.24, 80.09, 67.20], [87.73, -86.18, 83.18], [32.30, 79.19, -107.86]])
    assert np.abs(lab - want).max() < 0.02
    x = np.linspace(0.008857, 1.1, 20001)
    assert np.abs(SO.cbrt_newton(x) - np.cbrt(x)).max() < 3e-16


def four_connected(labels):
    """True if every label's pixels form one 4-connected component."""
    from scipy import ndimage
    for l in np.unique(labels):
        _, n = ndimage.label(labels == l)
        if n != 1:
            return False
    return True


@pytest.mark.parametrize("H,W,n,comp", [(64, 48, 20, 10.0), (96, 128, 30, 200.0), (80, 80, 60, 1.0)])
def test_oracle_structure(H, W, n, comp):
    img = np.moveaxis(smooth_video(1, H, W, seed=H + n)[0], 0, -1)
    near, K = SO.slic_nearest(img, n, comp)
    assert near.min() >= 0 and near.max() < K
    lab = SO.slic_labels(img, n, comp)
    ids = np.unique(lab)
    assert ids[-1] <= K + 4 and ids[-1] >= 1
    keep = ids[ids > 0]
    assert np.array_equal(keep, np.arange(1, len(keep) + 1))                 # contiguous from start_label
    min_size = int(0.5 * H * W / K)
    sizes = np.bincount(lab.ravel())
    assert (sizes[1:] >= min_size).all()                                      # small components were merged away
    if comp >= 10.0:                                                          # compact segments: merging cannot split them again
        assert four_connected(np.where(lab > 0, lab, -1))
    # without the connectivity pass the labels are the nearest centres + 1
    assert np.array_equal(SO.slic_labels(img, n, comp, connectivity=False), near + 1)


def test_connectivity_hand_case():
    # two stripes of label 0 separated by label 1, plus an isolated pixel of label 2 inside the first stripe
    seg = np.zeros((6, 6), dtype=np.int64)
    seg[:, 3] = 1
    seg[2, 1] = 2
    out = SO.enforce_connectivity(seg, min_size=2, max_size=100)
    # scan order: left stripe -> 1, column 3 -> 2, right stripe -> 3; the single pixel (size 1 < 2) joins its neighbour
    assert out[0, 0] == 1 and out[0, 3] == 2 and out[0, 5] == 3 and out[2, 1] == 1
    assert sorted(np.unique(out)) == [1, 2, 3]
    # max_size cuts a big component: 6 x 3 = 18 pixels with max_size 10 -> the search stops at 10, the rest is a new segment
    out = SO.enforce_connectivity(np.zeros((6, 3), dtype=np.int64), min_size=1, max_size=10)
    assert np.bincount(out.ravel())[1] == 10 and out.max() >= 2


# ---- the kernels on the host simulator ------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def sim():
    from sapienza_video_contrastive_b200 import _lib
    from tests.cusim import build_sim
    return _lib.CrwLib(build_sim.build())


def run_slic(lib, vid, counts, comp, conn, n_iter=10, device=None, reps=2):
    F, _, H, W = vid.shape
    ns = (ctypes.c_int * F)(*counts)
    wb = lib.crw_slic_workspace_bytes(F, H, W, ns, n_iter)
    assert wb > 0
    t = torch.from_numpy(vid)
    if device is not None:
        t = t.to(device)
    ws = torch.zeros(wb, dtype=torch.uint8, device=t.device)
    out = torch.zeros(F, H, W, dtype=torch.int32, device=t.device)
    for _ in range(reps):                              # twice: the workspace must be reusable without re-zeroing
        lib.check(lib.crw_slic(t.data_ptr(), F, H, W, ns, comp, n_iter, conn, out.data_ptr(), ws.data_ptr(), wb, None), "slic")
    if device is not None:
        torch.cuda.synchronize()
    return out.cpu().numpy()


def oracle_labels(vid, counts, comp, conn, n_iter=10):
    return np.stack([SO.slic_labels(np.moveaxis(vid[f], 0, -1), counts[f], comp, n_iter, bool(conn)) for f in range(len(counts))])


@pytest.mark.parametrize("F,H,W,counts,comp,conn", [
    (2, 40, 32, [12, 7], 10.0, 0),
    (2, 40, 32, [12, 7], 10.0, 1),
    (1, 48, 64, [30], 200.0, 1),             # the reference's defaults (--num-sp 30 --compactness 200)
    (1, 40, 40, [60], 1.0, 1),               # colour-dominated: ragged segments, many small components to merge
    (1, 16, 16, [400], 5.0, 1),              # more centres asked than pixels
])
def test_sim_matches_oracle(sim, F, H, W, counts, comp, conn):
    vid = smooth_video(F, H, W, seed=H * W + counts[0], sigma=2.0 if comp < 5 else 4.0)
    got = run_slic(sim, vid, counts, comp, conn, reps=2 if H == 16 else 1)       # (the simulator is slow: one case checks the re-use)
    assert np.array_equal(got, oracle_labels(vid, counts, comp, conn))


def test_sim_constant_frame_and_bad_arguments(sim):
    vid = np.full((1, 3, 32, 32), 0.5, dtype=np.float32)
    got = run_slic(sim, vid, [16], 10.0, 1)
    assert np.array_equal(got, oracle_labels(vid, [16], 10.0, 1))
    ns = (ctypes.c_int * 1)(0)
    assert sim.crw_slic_workspace_bytes(1, 32, 32, ns, 10) == 0
    ns = (ctypes.c_int * 1)(100000)                    # 256 x 256 with 100000 points: more than 2048 centres
    assert sim.crw_slic_workspace_bytes(1, 256, 256, ns, 10) == 0
    t = torch.zeros(1, 3, 32, 32)
    out = torch.zeros(1, 32, 32, dtype=torch.int32)
    ns = (ctypes.c_int * 1)(16)
    assert sim.crw_slic(t.data_ptr(), 1, 32, 32, ns, 10.0, 10, 1, out.data_ptr(), out.data_ptr(), 16, None) != 0    # workspace too small
    assert sim.crw_slic(t.data_ptr(), 1, 32, 32, ns, 0.0, 10, 1, out.data_ptr(), out.data_ptr(), 1 << 30, None) != 0  # compactness


def random_label_maps(F, H, W, n_labels, seed, blobs):
    """Label maps with components of every size: blurred noise quantised into n_labels levels (blobs) or raw noise."""
    import cv2
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(F):
        v = rng.standard_normal((H, W)).astype(np.float32)
        if blobs:
            v = cv2.GaussianBlur(v, (0, 0), blobs)
        r = np.argsort(np.argsort(v.ravel())).reshape(H, W)
        out.append((r * n_labels // (H * W)).astype(np.int32))
    return np.stack(out)


CONNECT_CASES = [   # F, H, W, labels, min_size, max_size, blur
    (2, 24, 31, 3, 4, 40, 2.0),        # big blobs against a small max_size: the search is cut again and again
    (2, 24, 31, 5, 6, 10_000, 1.0),    # no cut, many merges
    (1, 40, 40, 2, 0, 7, 3.0),         # min_size 0: nothing merges, segments of at most 7 pixels
    (1, 16, 16, 4, 3, 3, 0),           # salt and pepper, max_size == min_size
    (1, 33, 65, 1, 1, 100, 0),         # one label everywhere: a frontier much wider than a warp
    (1, 8, 8, 2, 100, 1, 0),           # max_size 1: no search at all, every pixel merges into its predecessor
]


def run_connect(lib, seg, mn, mx, device=None, reps=2):
    F, H, W = seg.shape
    t = torch.from_numpy(seg)
    if device is not None:
        t = t.to(device)
    wb = lib.crw_label_connectivity_workspace_bytes(F, H, W)
    ws = torch.zeros(wb, dtype=torch.uint8, device=t.device)
    out = torch.zeros(F, H, W, dtype=torch.int32, device=t.device)
    for _ in range(reps):
        lib.check(lib.crw_label_connectivity(t.data_ptr(), F, H, W, mn, mx, out.data_ptr(), ws.data_ptr(), wb, None), "connectivity")
    if device is not None:
        torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("F,H,W,nl,mn,mx,blobs", CONNECT_CASES)
def test_sim_connectivity_matches_oracle(sim, F, H, W, nl, mn, mx, blobs):
    seg = random_label_maps(F, H, W, nl, seed=H * W + nl, blobs=blobs)
    got = run_connect(sim, seg, mn, mx)
    for f in range(F):
        assert np.array_equal(got[f], SO.enforce_connectivity(seg[f].astype(np.int64), mn, mx))


# ---- the kernels on the GPU ---------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("F,H,W,counts,comp,conn", [
    (3, 64, 48, [20, 12, 33], 10.0, 0),
    (3, 64, 48, [20, 12, 33], 10.0, 1),
    (2, 128, 160, [30, 100], 200.0, 1),
    (1, 256, 256, [30], 200.0, 1),           # the training shape (img_size 256) with the reference's defaults
    (1, 256, 256, [150], 2.0, 1),
    (40, 32, 32, [9 + (i % 7) for i in range(40)], 20.0, 1),     # more frames than one launch carries
])
def test_gpu_matches_oracle(F, H, W, counts, comp, conn):
    from sapienza_video_contrastive_b200 import _lib
    vid = smooth_video(F, H, W, seed=H + W + F, sigma=2.0 if comp < 5 else 5.0)
    got = run_slic(_lib.lib(), vid, counts, comp, conn, device="cuda")
    assert np.array_equal(got, oracle_labels(vid, counts, comp, conn))


@pytest.mark.gpu
def test_gpu_compute_mask_mirror():
    """compute_mask (superpixels.py:24-63): shapes, dtype, the random per-frame counts drawn like the reference, and the
    labels of every frame equal to the oracle's; the result feeds the superpixel walk."""
    from sapienza_video_contrastive_b200 import superpixels as SP
    vid = torch.from_numpy(smooth_video(4, 64, 64, seed=11))
    torch.manual_seed(5)
    m = SP.compute_mask(vid, "slic", 16, 1.0, True, 6, 30.0)
    torch.manual_seed(5)
    counts = [torch.randint(low=13, high=19, size=(1,)).item() for _ in range(4)]
    assert m.shape == (4, 3, 64, 64) and m.dtype == torch.int64 and m.is_cuda
    want = oracle_labels(vid.numpy(), counts, 30.0, 1)
    for c in range(3):
        assert np.array_equal(m[:, c].cpu().numpy(), want)
    b = SP.compute_mask(vid.reshape(2, 2, 3, 64, 64), "slic", 16, 1.0, False, 0, 30.0)
    assert b.shape == (2, 2, 3, 64, 64)
    assert np.array_equal(b[:, :, 0].reshape(4, 64, 64).cpu().numpy(), oracle_labels(vid.numpy(), [16] * 4, 30.0, 1))
    one = SP.compute_sp_slic(vid[0].permute(1, 2, 0).numpy(), 16, 30.0)
    assert np.array_equal(one.cpu().numpy(), want[0] if counts[0] == 16 else oracle_labels(vid.numpy()[:1], [16], 30.0, 1)[0])
    with pytest.raises(NotImplementedError):
        SP.compute_mask(vid, "fh", 16, 1.0, False, 0, 30.0)


@pytest.mark.gpu
@pytest.mark.parametrize("F,H,W,nl,mn,mx,blobs", CONNECT_CASES + [(3, 128, 96, 6, 50, 700, 3.0), (1, 256, 256, 4, 300, 4000, 6.0)])
def test_gpu_connectivity_matches_oracle(F, H, W, nl, mn, mx, blobs):
    from sapienza_video_contrastive_b200 import _lib
    seg = random_label_maps(F, H, W, nl, seed=H * W + nl, blobs=blobs)
    got = run_connect(_lib.lib(), seg, mn, mx, device="cuda")
    for f in range(F):
        assert np.array_equal(got[f], SO.enforce_connectivity(seg[f].astype(np.int64), mn, mx))


# ---- randomised: the connectivity pass on the simulator, host-mirror argument handling -------------------------------------

def test_sim_connectivity_random_maps(sim):
    """hypothesis over small label maps and size limits: the warp-parallel queue must reproduce the sequential scan-order search
    (cuts at max_size, merges below min_size, the "last labelled neighbour met" rule) exactly."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=30, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.integers(3, 14), st.integers(3, 40), st.integers(1, 5), st.integers(0, 12), st.integers(1, 70), st.integers(0, 2 ** 31 - 1),
           st.sampled_from([0.0, 1.0, 2.5]))
    def run(H, W, nl, mn, mx, seed, blobs):
        seg = random_label_maps(1, H, W, nl, seed=seed, blobs=blobs)
        got = run_connect(sim, seg, mn, mx, reps=1)
        assert np.array_equal(got[0], SO.enforce_connectivity(seg[0].astype(np.int64), mn, mx))

    run()


def test_compute_mask_rejects_other_methods_before_touching_the_gpu():
    from sapienza_video_contrastive_b200 import superpixels as SP
    vid = torch.zeros(2, 3, 16, 16)
    with pytest.raises(NotImplementedError):
        SP.compute_mask(vid, "fh", 8, 1.0, False, 0, 10.0)
    with pytest.raises(RuntimeError):
        SP.slic_frames(vid, 8, 10.0)                       # CPU tensor: there is no CPU path
