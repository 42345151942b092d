"""CPU oracle of the superpixel label-map producer (second half).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED against the reference's dependency: code/data/superpixels.py:9-63 calls `skimage.segmentation.slic` (after
`cv2.normalize`); scikit-image is absent from this image and the reference ships no label-map fixtures.  This file
restates the PUBLISHED algorithm - scikit-image 0.19+ `segmentation/slic_superpixels.py`, `_slic.pyx` and `util/_regular_grid.py`
(Achanta et al., "SLIC superpixels", TPAMI 2012) as compute_sp_slic calls it: RGB input, convert2lab, sigma=0, max_num_iter=10,
slic_zero=False, enforce_connectivity=True (min / max size factors 0.5 / 3), start_label=1 - and it is the definition the CUDA
kernels (csrc/slic.cu) are tested against, bit for bit.  What IS pinned: the `cv2.normalize(img, None, 0, 255, NORM_MINMAX,
CV_8U)` step against OpenCV itself (tests/test_slic.py), the grid of initial centres against hand-computed values, and structural
properties of the result (labels contiguous from 1, 4-connected segments, no segment below the minimum size).

Stated deviations from scikit-image, chosen so that a parallel implementation is bit-reproducible:
  * the Lab features (already scaled by 1 / compactness) are quantised to 2^-24 before clustering, so that the centre update is an
    exact, order-independent integer sum instead of a floating-point sum in pixel order;
  * the cube root of rgb2lab is a fixed Newton iteration in IEEE double (+, *, / only), and the sRGB -> linear table of the 256
    possible inputs comes from libm `pow`: both are within an ulp or two of what numpy computes, and the same bits on the GPU;
  * a segment that lost all its pixels stays empty (scikit-image divides by zero there).
"""
from __future__ import annotations

import math

import numpy as np

_M = ((0.412453, 0.357580, 0.180423), (0.212671, 0.715160, 0.072169), (0.019334, 0.119193, 0.950227))   # skimage xyz_from_rgb
_WHITE = (0.95047, 1.0, 1.08883)                     # D65, 2 degree observer (skimage.color.rgb2lab defaults)
QBITS = 24
Q = float(1 << QBITS)
CBRT_ITERS = 12


def srgb_to_linear_table() -> np.ndarray:
    """skimage.color.rgb2xyz's gamma expansion on the 256 possible 8-bit values, with libm pow (the C side calls the same)."""
    out = np.empty(256)
    for i in range(256):
        v = i / 255.0
        out[i] = math.pow((v + 0.055) / 1.055, 2.4) if v > 0.04045 else v / 12.92
    return out


def normalize_minmax_u8(img: np.ndarray) -> np.ndarray:
    """cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8U) on a float32 frame: scale and shift from the global min / max
    in double, rounded to float32 and applied with ONE rounding (OpenCV's fused multiply-add), then round half to even and
    saturate."""
    img = np.asarray(img, dtype=np.float32)
    mn, mx = float(img.min()), float(img.max())
    scale = 255.0 * (1.0 / (mx - mn)) if (mx - mn) > np.finfo(np.float64).eps else 0.0
    shift = 0.0 - mn * scale
    v = (img.astype(np.float64) * np.float64(np.float32(scale)) + np.float64(np.float32(shift))).astype(np.float32)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def cbrt_newton(x: np.ndarray) -> np.ndarray:
    """Cube root on (0, ~1.1]: y <- (2 y + x / y^2) / 3 from y = (x + 2) / 3, a fixed number of IEEE-double steps."""
    y = (x + 2.0) / 3.0
    for _ in range(CBRT_ITERS):
        y = (2.0 * y + x / (y * y)) / 3.0
    return y


def rgb8_to_lab(u8: np.ndarray) -> np.ndarray:
    """skimage.color.rgb2lab(img_as_float(u8)) in float64, with the operation order the kernel uses."""
    lin = srgb_to_linear_table()[u8]                                   # (H, W, 3)
    xyz = [((lin[..., 0] * _M[r][0] + lin[..., 1] * _M[r][1]) + lin[..., 2] * _M[r][2]) / _WHITE[r] for r in range(3)]
    f = [np.where(v > 0.008856, cbrt_newton(np.maximum(v, 1e-3)), 7.787 * v + 16.0 / 116.0) for v in xyz]
    return np.stack([116.0 * f[1] - 16.0, 500.0 * (f[0] - f[1]), 200.0 * (f[1] - f[2])], -1)


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
        for d in range(2):
            steps[d] = sd[d]
            space = float(np.prod(sd[d + 1:]))
            steps[d + 1:] = (space / n_points) ** (1.0 / (3 - d - 1))
            if (sd >= steps).all():
                break
    starts = (steps // 2).astype(int)
    steps = np.round(steps).astype(int)
    un = np.argsort(order, kind="stable")
    starts, steps = starts[un], steps[un]
    return int(starts[1]), int(steps[1]), int(starts[2]), int(steps[2])


def slic_nearest(img: np.ndarray, n_segments: int, compactness: float, n_iter: int = 10):
    """The clustering of _slic_cython: -> (nearest (H, W) int64 in 0..K-1, K)."""
    u8 = normalize_minmax_u8(img)
    H, W = u8.shape[:2]
    lab = rgb8_to_lab(u8) * (1.0 / compactness)
    feat_q = np.rint(lab * Q).astype(np.int64)
    feat = feat_q.astype(np.float64) / Q
    y0, sy, x0, sx = regular_grid_2d(H, W, n_segments)
    gy, gx = np.arange(y0, H, sy), np.arange(x0, W, sx)
    K = len(gy) * len(gx)
    cy = np.repeat(gy, len(gx)).astype(np.float64)
    cx = np.tile(gx, len(gy)).astype(np.float64)
    cc = np.zeros((K, 3))                                              # the centres start with zero colour (slic_superpixels.py)
    alive = np.ones(K, dtype=bool)
    step = float(max(1, sy, sx))
    sw = 1.0 / (step * step)
    ys, xs = np.arange(H, dtype=np.float64)[:, None], np.arange(W, dtype=np.float64)[None, :]
    yi = np.broadcast_to(np.arange(H, dtype=np.int64)[:, None], (H, W)).ravel()
    xi = np.broadcast_to(np.arange(W, dtype=np.int64)[None, :], (H, W)).ravel()
    nearest = np.zeros((H, W), dtype=np.int64)
    for _ in range(n_iter):
        dist = np.full((H, W), np.finfo(np.float64).max)
        for k in range(K):                                             # increasing k and a strict <: the first centre wins ties
            if not alive[k]:
                continue
            ya, yb = int(max(cy[k] - 2 * sy, 0)), int(min(cy[k] + 2 * sy + 1, H))
            xa, xb = int(max(cx[k] - 2 * sx, 0)), int(min(cx[k] + 2 * sx + 1, W))
            if ya >= yb or xa >= xb:
                continue
            dy = (cy[k] - ys[ya:yb]) * (cy[k] - ys[ya:yb])
            dx = (cx[k] - xs[:, xa:xb]) * (cx[k] - xs[:, xa:xb])
            d = (dy + dx) * sw
            dc = feat[ya:yb, xa:xb] - cc[k]
            d = d + ((dc[..., 0] * dc[..., 0] + dc[..., 1] * dc[..., 1]) + dc[..., 2] * dc[..., 2])
            win = dist[ya:yb, xa:xb]
            better = win > d
            win[better] = d[better]
            nearest[ya:yb, xa:xb][better] = k
        flat = nearest.ravel()
        n = np.bincount(flat, minlength=K).astype(np.int64)
        sum_y, sum_x, sum_f = np.zeros(K, dtype=np.int64), np.zeros(K, dtype=np.int64), np.zeros((K, 3), dtype=np.int64)
        np.add.at(sum_y, flat, yi)
        np.add.at(sum_x, flat, xi)
        for c in range(3):
            np.add.at(sum_f[:, c], flat, feat_q[..., c].ravel())
        alive = n > 0
        nn = np.maximum(n, 1).astype(np.float64)
        cy, cx = sum_y.astype(np.float64) / nn, sum_x.astype(np.float64) / nn
        cc = (sum_f.astype(np.float64) / nn[:, None]) / Q
    return nearest, K


def enforce_connectivity(nearest: np.ndarray, min_size: int, max_size: int, start_label: int = 1) -> np.ndarray:
    """_enforce_label_connectivity_cython: scan-order breadth-first relabelling; a component smaller than min_size is merged into
    the last already-labelled neighbour the search met; the search stops at max_size pixels (the rest becomes a new segment)."""
    H, W = nearest.shape
    seg = nearest.tolist()
    con = [[-1] * W for _ in range(H)]
    new_label = start_label
    nb = ((0, 1), (0, -1), (1, 0), (-1, 0))                            # x+1, x-1, y+1, y-1
    for y in range(H):
        for x in range(W):
            if con[y][x] >= 0:
                continue
            adjacent = 0
            label = seg[y][x]
            con[y][x] = new_label
            coords = [(y, x)]
            visited = 0
            while visited < len(coords) < max_size:
                py, px = coords[visited]
                for dy, dx in nb:
                    yy, xx = py + dy, px + dx
                    if 0 <= xx < W and 0 <= yy < H:
                        if seg[yy][xx] == label and con[yy][xx] == -1:
                            con[yy][xx] = new_label
                            coords.append((yy, xx))
                            if len(coords) >= max_size:
                                break
                        elif con[yy][xx] >= 0 and con[yy][xx] != new_label:
                            adjacent = con[yy][xx]
                visited += 1
            if len(coords) < min_size:
                for py, px in coords:
                    con[py][px] = adjacent
            else:
                new_label += 1
    return np.asarray(con, dtype=np.int64)


def slic_labels(img: np.ndarray, n_segments: int, compactness: float, n_iter: int = 10, connectivity: bool = True) -> np.ndarray:
    """compute_sp_slic (data/superpixels.py:9-16) for one float (H, W, 3) frame -> (H, W) int64 labels starting at 1."""
    nearest, K = slic_nearest(img, n_segments, compactness, n_iter)
    if not connectivity:
        return nearest + 1
    H, W = nearest.shape
    seg_size = (H * W) / K
    return enforce_connectivity(nearest, int(0.5 * seg_size), int(3 * seg_size), 1)
