"""Patch-grid producer: the reference builds the training input of the patch walk on the CPU, window by
window, inside DataLoader workers (code/utils/augs.py:59-82: view_as_windows -> RandomResizedCrop(64, scale (0.7, 0.9)) ->
ToTensor -> Normalize, 49 PIL resizes per frame).  Here the crop parameters stay on the host - they are index math, drawn with
the reference's own generator calls so that a seed reproduces the reference's crops - and one kernel does the pixel work of
every window of every frame of a batch (csrc/patchgrid.cu, bit-exact with Pillow's BILINEAR resize).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops

IMG_MEAN = (0.4914, 0.4822, 0.4465)         # code/utils/augs.py:10-12
IMG_STD = (0.2023, 0.1994, 0.2010)


def draw_patch_boxes(n_frames: int, n_windows: int, win: int = 64, scale: Tuple[float, float] = (0.7, 0.9),
                     ratio: Tuple[float, float] = (3.0 / 4.0, 4.0 / 3.0)) -> torch.Tensor:
    """The RandomResizedCrop parameters of augs.py:63-68 for every window of every frame, in the reference's order (frame by
    frame, window by window), consuming torch's global CPU generator exactly as torchvision's RandomResizedCrop.get_params
    does (augs.py builds `spatial_jitter` from that class).  -> (n_frames, n_windows, 4) int32 {top, left, height, width}."""
    from torchvision import transforms
    dummy = torch.empty(3, win, win)
    out = torch.empty(n_frames, n_windows, 4, dtype=torch.int32)
    for f in range(n_frames):
        for p in range(n_windows):
            i, j, h, w = transforms.RandomResizedCrop.get_params(dummy, list(scale), list(ratio))
            out[f, p] = torch.tensor([i, j, h, w], dtype=torch.int32)
    return out


def patch_grid_frames(frames: torch.Tensor, boxes: torch.Tensor, win: int = 64, stride: int = 32, out_size: Optional[int] = None,
                      mean: Sequence[float] = IMG_MEAN, std: Sequence[float] = IMG_STD) -> torch.Tensor:
    """frames (F, H, W, 3) uint8 CUDA, boxes (F, P, 4) int32 -> (F, P * 3, out_size, out_size) fp32 CUDA: every window cropped,
    resized (Pillow BILINEAR, bit-exact), scaled to [0, 1] and normalised.  One launch."""
    if not frames.is_cuda:
        raise RuntimeError("patch_grid_frames runs on CUDA tensors only; there is no CPU path")
    ops.check_device(frames.device)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError("frames must be (F, H, W, 3) uint8, got %s %s" % (tuple(frames.shape), frames.dtype))
    frames = frames.contiguous()
    F, H, W, _ = frames.shape
    out_size = int(out_size or win)
    nwx, nwy = (W - win) // stride + 1, (H - win) // stride + 1
    P = nwx * nwy
    boxes = boxes.to(device=frames.device, dtype=torch.int32).contiguous()
    if tuple(boxes.shape) != (F, P, 4):
        raise ValueError("boxes must be (%d, %d, 4), got %s" % (F, P, tuple(boxes.shape)))
    out = torch.empty(F, P * 3, out_size, out_size, dtype=torch.float32, device=frames.device)
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    L = _lib.lib()
    L.check(L.crw_patch_grid(frames.data_ptr(), boxes.data_ptr(), F, H, W, int(win), int(stride), out_size,
                             ctypes.addressof(m3), ctypes.addressof(s3), out.data_ptr(), torch.cuda.current_stream().cuda_stream), "patch_grid")
    return out


def patch_grid(transform=None, shape=(64, 64, 3), stride=(0.5, 0.5), device="cuda"):
    """Mirror of augs.py:59-82: returns aug(x) for ONE frame x ((H, W, 3) uint8 numpy / tensor, (3, H, W) tensor or PIL image)
    -> (P * 3, 64, 64) fp32 CPU tensor, as the reference's aug does.  `transform` must be the reference's NORM pair (ToTensor +
    Normalize): None = those constants, or a (mean, std) pair.  Consumes numpy's and torch's global generators like the reference
    (one np.random.random() for the stride at construction, RandomResizedCrop draws per window per call)."""
    st = np.random.random() * (stride[1] - stride[0]) + stride[0]                 # augs.py:60
    step = int(shape[0] * st)
    mean, std = (IMG_MEAN, IMG_STD) if transform is None else transform

    def aug(x):
        if torch.is_tensor(x):
            x = x.numpy().transpose(1, 2, 0) if x.shape[0] == 3 and x.dim() == 3 and x.shape[-1] != 3 else x.numpy()
        elif "PIL" in str(type(x)):
            x = np.array(x)
        x = np.ascontiguousarray(x, dtype=np.uint8)
        H, W = x.shape[:2]
        P = ((H - shape[0]) // step + 1) * ((W - shape[1]) // step + 1)
        boxes = draw_patch_boxes(1, P, shape[0])
        out = patch_grid_frames(torch.from_numpy(x)[None].to(device), boxes, shape[0], step, shape[0], mean, std)
        return out[0].cpu()

    return aug
