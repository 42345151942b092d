"""Superpixel label-map producer (second half): the reference computes one SLIC segmentation per frame on the
CPU, inside DataLoader workers (code/data/superpixels.py:9-63, called from data/kinetics.py:118-126: cv2.normalize ->
skimage.segmentation.slic per frame -> stack -> repeat over 3 channels).  Here one call segments every frame of a batch of
clips on the GPU (csrc/slic.cu); the per-frame segment counts of --randomise-superpixels stay on the host and are drawn with
the reference's own generator calls.

The segmentation follows scikit-image's published SLIC (DESIGN.md sections 2 and 4 list the deviations; scikit-image itself is not in
this image, so parity with it is unpinned; the min-max normalisation is bit-exact with OpenCV).  Felzenszwalb ("fh") is not
built: `compute_mask` raises for it rather than falling back to anything.
"""
from __future__ import annotations

import ctypes
from typing import Sequence, Union

import numpy as np
import torch

from . import _lib, ops


def slic_frames(video: torch.Tensor, n_segments: Union[int, Sequence[int]], compactness: float, n_iter: int = 10,
                enforce_connectivity: bool = True) -> torch.Tensor:
    """video (F, 3, H, W) fp32 CUDA -> (F, H, W) int32 CUDA label maps (labels from 1), one launch sequence for all frames."""
    if not video.is_cuda:
        raise RuntimeError("slic_frames needs a CUDA tensor: this package has no CPU path")
    if video.dim() != 4 or video.shape[1] != 3:
        raise ValueError("video must be (F, 3, H, W), got %s" % (tuple(video.shape),))
    ops.check_device(video.device)
    L = _lib.lib()
    v = video.detach().to(torch.float32).contiguous()
    F, _, H, W = v.shape
    counts = [int(n_segments)] * F if isinstance(n_segments, (int, np.integer)) else [int(n) for n in n_segments]
    if len(counts) != F:
        raise ValueError("need one segment count per frame (%d), got %d" % (F, len(counts)))
    ns = (ctypes.c_int * F)(*counts)
    wb = L.crw_slic_workspace_bytes(F, H, W, ns, int(n_iter))
    if wb == 0:
        raise ValueError("slic_frames: unsupported segment counts %s for %d x %d frames (1 <= centres <= 2048)" % (sorted(set(counts)), H, W))
    ws = torch.empty(wb, dtype=torch.uint8, device=v.device)
    out = torch.empty(F, H, W, dtype=torch.int32, device=v.device)
    with torch.cuda.device(v.device):
        L.check(L.crw_slic(v.data_ptr(), F, H, W, ns, float(compactness), int(n_iter), 1 if enforce_connectivity else 0, out.data_ptr(),
                           ws.data_ptr(), wb, torch.cuda.current_stream(v.device).cuda_stream), "slic")
    return out


def compute_sp_slic(img, num_components: int, compactness: float, device=None) -> torch.Tensor:
    """superpixels.py:9-16 for one frame: img (H, W, 3) float (tensor or ndarray) -> (H, W) int64 labels on the GPU."""
    t = torch.as_tensor(img)
    if not t.is_cuda:
        t = t.to(device if device is not None else "cuda")
    return slic_frames(t.permute(2, 0, 1).unsqueeze(0), num_components, compactness)[0].long()


def compute_mask(video, sp_method, num_components, p, randomise_superpixels, randomise_superpixels_range, compactness, device=None):
    """superpixels.py:24-63: video (T, 3, H, W) float -> mask (T, 3, H, W) int64, the label map of every frame repeated over
    the channel axis (an expanded view here, the consumer reads channel 0: model.py:298).  Unlike the reference the result stays
    a CUDA tensor (no .numpy()): it feeds CRW.forward(sp_mask=...) directly.  A (B, T, 3, H, W) batch is accepted too and is
    segmented in one go, drawing the random counts clip by clip as a DataLoader with one worker would."""
    if sp_method == "random":
        method = np.random.choice(["slic", "fh"], 1, p=[p, 1 - p])          # superpixels.py:30-32 (same generator call)
        method = str(method[0])
    else:
        method = sp_method
    if method != "slic":
        raise NotImplementedError("superpixel method %r: only SLIC is built (Felzenszwalb is out of scope, DESIGN.md section 7)" % (method,))
    t = torch.as_tensor(video)
    if not t.is_cuda:
        t = t.to(device if device is not None else "cuda")
    lead = t.shape[:-3]
    frames = t.reshape(-1, *t.shape[-3:])
    if randomise_superpixels:
        low = num_components - randomise_superpixels_range // 2
        high = num_components + randomise_superpixels_range // 2
        counts = [torch.randint(low=low, high=high, size=(1,)).item() for _ in range(frames.shape[0])]     # superpixels.py:41-45
    else:
        counts = [int(num_components)] * frames.shape[0]
    labels = slic_frames(frames, counts, compactness).long()
    return labels.unsqueeze(1).expand(-1, 3, -1, -1).reshape(*lead, 3, *labels.shape[-2:])      # splitting the leading axis: still a view
