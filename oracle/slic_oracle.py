"""CPU oracle of the superpixel label-map producer (second half).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED against the reference's dependency: code/data/superpixels.py:9-63 calls `skimage.segmentation.slic` (and
`cv2.normalize`), scikit-image is absent from this image, and the reference ships no label-map fixtures.  This file
restates the PUBLISHED algorithm - scikit-image 0.19+ `segmentation/slic_superpixels.py` + `_slic.pyx` (Achanta et al., SLIC
superpixels, TPAMI 2012) as called by compute_sp_slic: RGB input, convert2lab=True, sigma=0, max_num_iter=10, slic_zero=False,
start_label=1 - and is the definition the CUDA kernels (csrc/slic.cu) are tested against, bit for bit.  What is pinned: the
`cv2.normalize(img, None, 0, 255, NORM_MINMAX, CV_8U)` step against OpenCV itself (tests), the regular grid of initial centres
against hand-computed values, and structural properties (every label in 1..K, compact segments).

Two stated deviations from scikit-image, chosen so that a massively parallel implementation is bit-reproducible:
  * the Lab features (already scaled by 1 / compactness) are quantised to 2^-20 before clustering, so that the centre
    update is an exact integer sum (order-independent) instead of a floating-point sum in pixel order;
  * `enforce_connectivity` (a sequential flood fill whose merge rule depends on the visiting order) is NOT applied: the labels are
    the nearest-centre indices after 10 iterations (`enforce_connectivity=False` semantics).  Segment-mean pooling does not need
    connected segments; the number of labels stays <= the number of grid centres.
"""
from __future__ import annotations

import numpy as np

_M = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
_WHITE = np.array([0.95047, 1.0, 1.08883])          # D65, 2 degree observer (skimage.color.rgb2lab defaults)
Q = float(1 << 20)


def srgb_to_linear_table() -> np.ndarray:
    """skimage.color.rgb2xyz on the 256 possible 8-bit values (float64): the table both the oracle and the kernel use."""
    v = np.arange(256, dtype=np.float64) / 255.0
    return np.where(v > 0.04045, np.power((v + 0.055) / 1.055, 2.4), v / 12.92)


def normalize_minmax_u8(img: np.ndarray) -> np.ndarray:
    """cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8U) on a float32 (H, W, 3) frame: scale and shift from the global
    min / max (double), applied in float32, rounded half to even, saturated."""
    img = np.asarray(img, dtype=np.float32)
    mn, mx = float(img.min()), float(img.max())
    scale = 255.0 / (mx - mn) if mx > mn else 0.0
    shift = 0.0 - mn * scale
    v = img * np.float32(scale) + np.float32(shift)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def rgb8_to_lab(u8: np.ndarray) -> np.ndarray:
    """skimage.color.rgb2lab(img_as_float(u8)) in float64."""
    lin = srgb_to_linear_table()[u8]                                   # (H, W, 3)
    xyz = np.stack([lin[..., 0] * _M[r, 0] + lin[..., 1] * _M[r, 1] + lin[..., 2] * _M[r, 2] for r in range(3)], -1)
    xyz = xyz / _WHITE
    f = np.where(xyz > 0.008856, np.cbrt(xyz), 7.787 * xyz + 16.0 / 116.0)
    L = 116.0 * f[..., 1] - 16.0
    a = 500.0 * (f[..., 0] - f[..., 1])
    b = 200.0 * (f[..., 1] - f[..., 2])
    return np.stack([L, a, b], -1)


def regular_grid_2d(H: int, W: int, n_points: int):
    """skimage.util.regular_grid((1, H, W), n_points) -> (start_y, step_y, start_x, step_x) of the slices along y and x."""
    dims = np.array([1, H, W], dtype=float)
    order = np.argsort(dims, kind="stable")
    sd = dims[order]
    space = float(np.prod(sd))
    if space <= n_points:
        return 0, 1, 0, 1
    steps = np.full(3, (space / n_points) ** (1.0 / 3))
    if (sd < steps).any():
        for d in range(3):
            steps[d] = sd[d]
            space = space / sd[d]
            steps[d + 1:] = (space / n_points) ** (1.0 / (3 - d - 1)) if d < 2 else steps[d + 1:]
            if (sd >= steps).all():
                break
    starts = (steps // 2).astype(int)
    steps = np.round(steps).astype(int)
    un = np.argsort(order, kind="stable")
    starts, steps = starts[un], steps[un]
    return int(starts[1]), int(steps[1]), int(starts[2]), int(steps[2])


def slic_labels(img: np.ndarray, n_segments: int, compactness: float, n_iter: int = 10) -> np.ndarray:
    """compute_sp_slic (data/superpixels.py:9-16) for one float (H, W, 3) frame -> (H, W) int64 labels in 1..K."""
    u8 = normalize_minmax_u8(img)
    H, W = u8.shape[:2]
    lab = rgb8_to_lab(u8) * (1.0 / compactness)
    feat_q = np.rint(lab * Q).astype(np.int64)                         # quantised features (deviation 1)
    feat = feat_q.astype(np.float64) / Q
    y0, sy, x0, sx = regular_grid_2d(H, W, n_segments)
    gy, gx = np.arange(y0, H, sy), np.arange(x0, W, sx)
    K = len(gy) * len(gx)
    cy = np.repeat(gy, len(gx)).astype(np.float64)
    cx = np.tile(gx, len(gy)).astype(np.float64)
    cc = np.zeros((K, 3))
    step = float(max(1, sy, sx))
    sw = 1.0 / (step * step)
    ys, xs = np.arange(H)[:, None], np.arange(W)[None, :]
    nearest = np.zeros((H, W), dtype=np.int64)
    for _ in range(n_iter):
        dist = np.full((H, W), np.finfo(np.float64).max)
        for k in range(K):                                             # increasing k, strict <: the first centre wins ties
            ya, yb = int(max(cy[k] - 2 * step, 0)), int(min(cy[k] + 2 * step + 1, H))
            xa, xb = int(max(cx[k] - 2 * step, 0)), int(min(cx[k] + 2 * step + 1, W))
            if ya >= yb or xa >= xb:
                continue
            dy = (cy[k] - ys[ya:yb]) ** 2
            dx = (cx[k] - xs[:, xa:xb]) ** 2
            d = (dy + dx) * sw
            dc = feat[ya:yb, xa:xb] - cc[k]
            d = d + (dc[..., 0] * dc[..., 0] + dc[..., 1] * dc[..., 1] + dc[..., 2] * dc[..., 2])
            win = dist[ya:yb, xa:xb]
            better = d < win
            win[better] = d[better]
            nearest[ya:yb, xa:xb][better] = k
        n = np.bincount(nearest.ravel(), minlength=K).astype(np.int64)
        sum_y = np.bincount(nearest.ravel(), weights=None, minlength=K) * 0
        sum_y = np.zeros(K, dtype=np.int64)
        sum_x = np.zeros(K, dtype=np.int64)
        sum_f = np.zeros((K, 3), dtype=np.int64)
        np.add.at(sum_y, nearest.ravel(), np.broadcast_to(ys, (H, W)).ravel())
        np.add.at(sum_x, nearest.ravel(), np.broadcast_to(xs, (H, W)).ravel())
        for c in range(3):
            np.add.at(sum_f[:, c], nearest.ravel(), feat_q[..., c].ravel())
        nn = np.maximum(n, 1).astype(np.float64)                       # an empty segment collapses to the origin, as in _slic.pyx
        cy, cx = sum_y / nn, sum_x / nn
        cc = (sum_f / nn[:, None]) / Q
    return nearest + 1
