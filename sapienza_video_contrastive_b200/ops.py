"""Host-side operators: thin torch.autograd wrappers over the C ABI (include/crw_b200.h).

PyTorch is used for device memory, streams and autograd bookkeeping only; every arithmetic step of the hot path
is a kernel in libcrw_b200.so.  All functions require CUDA fp32 tensors and raise otherwise - there is no CPU or
PyTorch fallback.
"""
from __future__ import annotations

import threading
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import WALK_FLIP, WALK_FORCE_GENERAL, WALK_FORCE_SIMT, WALK_FORCE_TC, WALK_NO_CLUSTER, WALK_NO_TF32, WALK_SOFTMAX  # noqa: F401


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("crw_b200 operators run on CUDA tensors only (got %s); there is no CPU path" % t.device)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError("crw_b200 computes in fp32 (got %s)" % t.dtype)
    return t if t.is_contiguous() else t.contiguous()


_checked_devices = set()


def check_device(device) -> None:
    """sm_100a code only: refuse anything but a compute-capability-10.x GPU."""
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx in _checked_devices:
        return
    cap = torch.cuda.get_device_capability(idx)
    if cap[0] != 10:
        raise RuntimeError("libcrw_b200.so is built for sm_100a only; device %d has compute capability %d.%d" % (idx, *cap))
    _checked_devices.add(idx)


# ------------------------------------------------------------------------------------------------------------------
# workspaces: cached per (device, stream, key); zero-filled once (the walk's ticket counter relies on it)
# ------------------------------------------------------------------------------------------------------------------
_ws_cache = {}
_ws_lock = threading.Lock()


def _workspace(key, nbytes: int, device) -> torch.Tensor:
    k = (torch.device(device).index, _stream(), key)
    with _ws_lock:
        ws = _ws_cache.get(k)
        if ws is None or ws.numel() < nbytes:
            ws = torch.zeros(max(nbytes, 256), dtype=torch.uint8, device=device)
            _ws_cache[k] = ws
    return ws


# ------------------------------------------------------------------------------------------------------------------
# a1: patch mean pooling (model.py:116)
# ------------------------------------------------------------------------------------------------------------------
class _PoolPatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, maps, sm_limit):
        _need_cuda(maps)
        check_device(maps.device)
        maps = _f32c(maps)
        hw = maps.shape[-1] * maps.shape[-2]
        rows = maps.numel() // hw
        out = torch.empty(maps.shape[:-2], dtype=torch.float32, device=maps.device)
        L = _lib.lib()
        L.check(L.crw_pool_patch_fwd_sm(maps.data_ptr(), out.data_ptr(), rows, hw, sm_limit, _stream()), "pool_patch_fwd")
        ctx.shape = maps.shape
        ctx.sm_limit = sm_limit
        return out

    @staticmethod
    def backward(ctx, g):
        g = _f32c(g)
        shape = ctx.shape
        hw = shape[-1] * shape[-2]
        gm = torch.empty(shape, dtype=torch.float32, device=g.device)
        L = _lib.lib()
        L.check(L.crw_pool_patch_bwd_sm(g.data_ptr(), gm.data_ptr(), g.numel(), hw, ctx.sm_limit, _stream()), "pool_patch_bwd")
        return gm, None


def pool_patch(maps: torch.Tensor, sm_limit: int = 0) -> torch.Tensor:
    """(..., H, W) -> (...) spatial mean.  model.py:116.  `sm_limit` > 0 confines the forward and the backward kernel to that
    many SMs (see pipeline.PatchWalkPipeline); 0 = the whole GPU."""
    return _PoolPatch.apply(maps, int(sm_limit))


# ------------------------------------------------------------------------------------------------------------------
# a1: projection head (model.py:117).  Forward and input gradient are stock cuBLAS GEMMs (torch.matmul); the weight
# gradient - 128 x 512 outputs reduced over B*N*T rows - uses the split-K kernel.
# ------------------------------------------------------------------------------------------------------------------
# Per-DEVICE state (under nn.DataParallel every replica runs on its own Python thread, one per GPU, and autograd
# runs CUDA backward nodes on one engine thread per device - so "per device" is the granularity that is both visible from the
# forward thread and from the backward, and private to a replica): the opt-in flags, the side stream and the events still to
# be joined.  One replica can never pop another replica's events.
_dev_state = {}
_dev_lock = threading.Lock()
_warned = set()


def _warn_once(key, msg: str) -> None:
    """One line per process and reason whenever a call leaves the tensor-core path for a library / SIMT one."""
    if key not in _warned:
        _warned.add(key)
        import warnings
        warnings.warn("crw_b200: " + msg, RuntimeWarning, stacklevel=3)


class _DevState:
    __slots__ = ("pending", "side", "async_wgrad", "tc_head", "lock")

    def __init__(self):
        self.pending, self.side, self.async_wgrad, self.tc_head, self.lock = [], None, False, True, threading.Lock()


def _state(device=None) -> _DevState:
    idx = torch.device(device).index if device is not None else None
    idx = torch.cuda.current_device() if idx is None else idx
    st = _dev_state.get(idx)
    if st is None:
        with _dev_lock:
            st = _dev_state.setdefault(idx, _DevState())
    return st


def set_async_wgrad(flag: bool) -> None:
    """Opt-in (for the current device): run the head's weight-gradient kernel on a side stream so it overlaps the (HBM-bound)
    pooling backward.  The caller must then call `join_side_streams()` before anything consumes `weight.grad` (optimizer
    step, all-reduce, a copy to the host, the end of a CUDA-graph capture).  Off by default: the drop-in module needs no
    extra calls."""
    _state().async_wgrad = bool(flag)


def join_side_streams() -> None:
    """Make the current stream wait for every side-stream launch issued on the current device since the last join."""
    cur = torch.cuda.current_stream()
    st = _state()
    with st.lock:
        pend, st.pending = st.pending, []
    for ev in pend:
        cur.wait_event(ev)


def _side_stream(device) -> torch.cuda.Stream:
    st = _state(device)
    if st.side is None:
        with st.lock:
            if st.side is None:
                st.side = torch.cuda.Stream(device=device)
    return st.side


_ERR_UNSUPPORTED = -2


def set_tensor_core_head(enabled: bool) -> None:
    """Head forward / input gradient on the fused tcgen05 tf32 GEMM (default) or on the library fp32 GEMM (for the
    current device).  The tensor-core path feeds full-precision operands (tf32 big + small) but accumulates with the tensor core's
    truncating fp32 adder: ~5e-6 of the largest output at K = 512 against ~5e-7 for an fp32 sgemm."""
    _state().tc_head = bool(enabled)


_err_words = {}


def tc_error_word(device) -> torch.Tensor:
    """The device word (ONE per device) the tensor-core head GEMMs raise on a pipeline timeout.  The kernels also trap, so a
    timeout surfaces as a CUDA error at the next API call; the word tells a post-mortem which pipeline it was."""
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    with _ws_lock:
        w = _err_words.get(idx)
        if w is None:
            w = _err_words[idx] = torch.zeros(64, dtype=torch.int32, device="cuda:%d" % idx)
    return w[:1]


def _head_gemm(fn_name: str, a: torch.Tensor, weight: torch.Tensor, out_cols: int) -> Optional[torch.Tensor]:
    """a (R, K) and weight (D, C) through crw_head_fwd / crw_head_dgrad; None when the shape is not TMA-addressable."""
    L = _lib.lib()
    R = a.shape[0]
    D, C = weight.shape
    out = torch.empty(R, out_cols, dtype=torch.float32, device=a.device)
    rc = getattr(L, fn_name)(a.data_ptr(), weight.data_ptr(), out.data_ptr(), R, D, C, tc_error_word(a.device).data_ptr(), _stream())
    if rc == _ERR_UNSUPPORTED:
        return None
    L.check(rc, fn_name)
    return out


class _HeadLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        _need_cuda(x, weight)
        ctx.save_for_backward(x, weight)
        D, C = weight.shape
        st = _state(x.device)
        ctx.async_wgrad, ctx.tc_head = st.async_wgrad, st.tc_head        # the backward uses the forward's settings
        if st.tc_head and x.dtype == torch.float32 and weight.dtype == torch.float32:
            out = _head_gemm("crw_head_fwd", _f32c(x.reshape(-1, C)), _f32c(weight), D)
            if out is not None:
                return out.view(*x.shape[:-1], D)
            _warn_once(("head_fwd", D, C), "head forward (%d x %d): shape not TMA-addressable, using the library GEMM" % (D, C))
        return x.matmul(weight.t())               # shapes TMA cannot address (tiny or unaligned): library GEMM

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        gx = gw = None
        cur = torch.cuda.current_stream()
        fork = None
        if ctx.needs_input_grad[1] and ctx.async_wgrad:
            fork = torch.cuda.Event()
            fork.record(cur)                       # the side stream only needs g and x, not the input gradient below
        if ctx.needs_input_grad[0]:
            D, C = weight.shape
            gx = None
            if ctx.tc_head and g.dtype == torch.float32 and weight.dtype == torch.float32:
                gx = _head_gemm("crw_head_dgrad", _f32c(g.reshape(-1, D)), _f32c(weight), C)
                if gx is None:
                    _warn_once(("head_dgrad", D, C), "head input gradient (%d x %d): shape not TMA-addressable, using the library GEMM" % (D, C))
            gx = g.matmul(weight) if gx is None else gx.view(*g.shape[:-1], C)
        if ctx.needs_input_grad[1]:
            D, C = weight.shape
            g2, x2 = _f32c(g.reshape(-1, D)), _f32c(x.reshape(-1, C))
            R = g2.shape[0]
            L = _lib.lib()
            nbytes = L.crw_head_wgrad_workspace_bytes(R, D, C)
            gw = torch.empty(D, C, dtype=torch.float32, device=g.device)
            if fork is None:
                ws = _workspace(("wgrad", R, D, C), nbytes, g.device)
                L.check(L.crw_head_wgrad(g2.data_ptr(), x2.data_ptr(), gw.data_ptr(), R, D, C, ws.data_ptr(), ws.numel(),
                                         _stream()), "head_wgrad")
            else:
                side = _side_stream(g.device)
                side.wait_event(fork)
                with torch.cuda.stream(side):
                    ws = _workspace(("wgrad", R, D, C), nbytes, g.device)
                    L.check(L.crw_head_wgrad(g2.data_ptr(), x2.data_ptr(), gw.data_ptr(), R, D, C, ws.data_ptr(), ws.numel(),
                                             side.cuda_stream), "head_wgrad")
                    done = torch.cuda.Event()
                    done.record(side)
                for t in (g2, x2, gw):
                    t.record_stream(side)
                st = _state(g.device)
                with st.lock:
                    st.pending.append(done)
        return gx, gw


def head_linear(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """x (..., C) @ weight (D, C)^T -> (..., D); nn.Linear(bias=False) on the fused tensor-core GEMM (gemm_tf32.cu) with a
    split-K weight gradient; see set_tensor_core_head / set_async_wgrad."""
    return _HeadLinear.apply(x, weight)


# ------------------------------------------------------------------------------------------------------------------
# a2/a3: superpixel segment-mean pooling (model.py:296-325)
# ------------------------------------------------------------------------------------------------------------------
class _SegMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, maps, labels, SP, dilation):
        _need_cuda(maps, labels)
        check_device(maps.device)
        maps = _f32c(maps)
        if labels.dtype != torch.int64:
            labels = labels.long()
        B, C, T, Hm, Wm = maps.shape
        if labels.dim() != 4 or labels.shape[0] != B or labels.shape[1] != T:
            raise ValueError("labels must be (B,T,h,w), got %s" % (tuple(labels.shape),))
        h, w = labels.shape[-2:]
        L = _lib.lib()
        size_fn = L.crw_segmean_workspace_bytes if dilation is None else L.crw_segmean_dilated_workspace_bytes
        nbytes = size_fn(B, T, Hm, Wm, h, w, SP)
        if nbytes == 0:
            raise _lib.CrwError("segmean: %s" % L.crw_last_error().decode())
        ws = torch.empty(nbytes, dtype=torch.uint8, device=maps.device)     # kept for the backward
        out = torch.empty(B, SP, T, C, dtype=torch.float32, device=maps.device)
        sb, st, sy, sx = labels.stride()
        if dilation is None:
            L.check(L.crw_segmean_fwd(maps.data_ptr(), labels.data_ptr(), sb, st, sy, sx, B, C, T, Hm, Wm, h, w, SP,
                                      out.data_ptr(), ws.data_ptr(), nbytes, _stream()), "segmean_fwd")
        else:
            ksize, shape = dilation
            L.check(L.crw_segmean_dilated_fwd(maps.data_ptr(), labels.data_ptr(), sb, st, sy, sx, B, C, T, Hm, Wm, h, w, SP,
                                              int(ksize), _lib.DILATE_SHAPES[shape], out.data_ptr(), ws.data_ptr(), nbytes,
                                              _stream()), "segmean_dilated_fwd")
        ctx.ws = ws
        ctx.dims = (B, C, T, Hm, Wm, h, w, SP)
        ctx.dilated = dilation is not None
        return out

    @staticmethod
    def backward(ctx, g):
        g = _f32c(g)
        B, C, T, Hm, Wm, h, w, SP = ctx.dims
        gm = torch.empty(B, C, T, Hm, Wm, dtype=torch.float32, device=g.device)
        L = _lib.lib()
        fn = L.crw_segmean_dilated_bwd if ctx.dilated else L.crw_segmean_bwd
        L.check(fn(g.data_ptr(), ctx.ws.data_ptr(), ctx.ws.numel(), B, C, T, Hm, Wm, h, w, SP, gm.data_ptr(), _stream()),
                "segmean_bwd")
        return gm, None, None, None


def segment_mean(maps: torch.Tensor, labels: torch.Tensor, SP: int) -> torch.Tensor:
    """maps (B,C,T,Hm,Wm), labels (B,T,h,w) integer (any strides) -> (B,SP,T,C) per-superpixel feature means."""
    return _SegMean.apply(maps, labels, int(SP), None)


def segment_mean_dilated(maps: torch.Tensor, labels: torch.Tensor, SP: int, ksize: int, shape: str = "L1") -> torch.Tensor:
    """segment_mean over masks dilated by a ksize x ksize 'L1' | 'circle' | 'cross' structuring element
    (model.py:303-309, utils/__init__.py:590-608): overlapping masks, sizes and window counts of the dilated masks."""
    if shape not in _lib.DILATE_SHAPES:
        raise ValueError("dilation kernel shape must be one of %s, got %r" % (sorted(_lib.DILATE_SHAPES), shape))
    return _SegMean.apply(maps, labels, int(SP), (int(ksize), shape))


# ------------------------------------------------------------------------------------------------------------------
# L2 normalisation (model.py:118)
# ------------------------------------------------------------------------------------------------------------------
class _L2Norm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f):
        _need_cuda(f)
        check_device(f.device)
        f = _f32c(f)
        D = f.shape[-1]
        rows = f.numel() // D
        q = torch.empty_like(f)
        inv = torch.empty(rows, dtype=torch.float32, device=f.device)
        nrm = torch.empty_like(inv)
        L = _lib.lib()
        L.check(L.crw_l2norm_fwd(f.data_ptr(), q.data_ptr(), inv.data_ptr(), nrm.data_ptr(), rows, D, _stream()), "l2norm_fwd")
        ctx.save_for_backward(q, inv, nrm)
        return q

    @staticmethod
    def backward(ctx, g):
        q, inv, nrm = ctx.saved_tensors
        g = g.contiguous().clone()
        L = _lib.lib()
        L.check(L.crw_l2norm_bwd(q.data_ptr(), g.data_ptr(), inv.data_ptr(), nrm.data_ptr(), inv.numel(), q.shape[-1], _stream()),
                "l2norm_bwd")
        return g


def l2_normalize_last(f: torch.Tensor) -> torch.Tensor:
    """F.normalize(f, p=2, dim=-1, eps=1e-12) on contiguous rows."""
    return _L2Norm.apply(f)


# ------------------------------------------------------------------------------------------------------------------
# torch-compatible Philox replay
# ------------------------------------------------------------------------------------------------------------------
def torch_rand_threads(numel: int, device) -> int:
    """grid*block of torch's CUDA rand kernel for `numel` elements (ATen DistributionTemplates.h: block 256)."""
    p = torch.cuda.get_device_properties(device)
    grid = min((numel + 255) // 256, p.multi_processor_count * (p.max_threads_per_multi_processor // 256))
    return max(grid, 1) * 256


def torch_rand_offset_increment(numel: int, threads: int) -> int:
    return 4 * ((numel - 1) // (threads * 4) + 1)


def philox_uniform(n: int, seed: int, offset: int, threads: int, device) -> torch.Tensor:
    out = torch.empty(n, dtype=torch.float32, device=device)
    L = _lib.lib()
    L.check(L.crw_philox_uniform(out.data_ptr(), n, seed, offset, threads, _stream()), "philox_uniform")
    return out


_philox_ok = {}


def philox_replay_ok(device) -> bool:
    """One-time self check per device: does the in-kernel Philox replay reproduce torch.rand bit-for-bit?
    (It must, for the in-kernel edge dropout to be a drop-in for the reference's rand_like draws.)"""
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _philox_ok:
        gen = torch.cuda.default_generators[idx]
        state = gen.get_state()
        try:
            ok = True
            for numel in (1000, 48020, 700001):
                seed, off = gen.initial_seed(), gen.get_offset()
                ref = torch.rand(numel, device="cuda:%d" % idx)
                thr = torch_rand_threads(numel, idx)
                mine = philox_uniform(numel, seed, off, thr, "cuda:%d" % idx)
                ok = ok and bool(torch.equal(ref, mine)) and gen.get_offset() == off + torch_rand_offset_increment(numel, thr)
        finally:
            gen.set_state(state)
        _philox_ok[idx] = ok
    return _philox_ok[idx]


# ------------------------------------------------------------------------------------------------------------------
# a4-a6: the walk (model.py:366-413)
# ------------------------------------------------------------------------------------------------------------------
class _Walk(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, tau, rate, flags, rng, u12, u21p, rng_state, ts=None):
        _need_cuda(feats, u12, u21p)
        check_device(feats.device)
        feats = _f32c(feats)
        B, N, T, D = feats.shape
        dev = feats.device
        L = _lib.lib()
        need_grad = ctx.needs_input_grad[0]
        seed = off = thr = 0
        if rate > 0 and T > 1:
            if rng == "torch":
                if u12 is None:
                    # the reference's 2(T-1) rand_like draws, in its order; the backward ones are physically transposed
                    u12 = torch.stack([torch.rand(B, N, N, device=dev) for _ in range(T - 1)])
                    u21p = torch.stack([torch.rand(B, N, N, device=dev) for _ in range(T - 1)])
            elif rng == "philox":
                idx = dev.index if dev.index is not None else torch.cuda.current_device()
                gen = torch.cuda.default_generators[idx]
                numel = B * N * N
                thr = torch_rand_threads(numel, idx)
                seed, off = gen.initial_seed(), gen.get_offset()
                gen.set_offset(off + 2 * (T - 1) * torch_rand_offset_increment(numel, thr))
                u12 = u21p = None
            elif rng == "device":
                # graph-safe: {seed, offset} live in device memory and the kernel advances them itself
                if rng_state is None or rng_state.dtype != torch.int64 or rng_state.numel() != 2 or not rng_state.is_cuda:
                    raise ValueError("rng='device' needs rng_state: a CUDA int64 tensor {seed, offset}")
                thr = torch_rand_threads(B * N * N, dev)
                u12 = u21p = None
            else:
                raise ValueError("rng must be 'torch', 'philox' or 'device'")
            if u12 is not None:
                u12, u21p = _f32c(u12), _f32c(u21p)
                if tuple(u12.shape) != (T - 1, B, N, N) or tuple(u21p.shape) != (T - 1, B, N, N):
                    raise ValueError("dropout uniforms must be (T-1,B,N,N)")
        else:
            u12 = u21p = None
            rate = 0.0
        nbytes = L.crw_walk_workspace_bytes(B, N, T, D, flags)
        ws = _workspace(("walk", B, N, T, D, flags), nbytes, dev)
        q = torch.empty_like(feats)
        nw = max(T - 2, 0)
        # xent[0..nw) per-walk cross-entropies, xent[nw] their mean = the loss (written by the kernel)
        xent = torch.empty(nw + 1, dtype=torch.float32, device=dev) if nw > 0 else torch.zeros(1, dtype=torch.float32, device=dev)
        acc = torch.empty(max(nw, 1), dtype=torch.float32, device=dev)
        grad = torch.empty_like(feats) if (need_grad and nw > 0) else None
        if ts is None:
            L.check(L.crw_walk_fwd_bwd(feats.data_ptr(), B, N, T, D, float(tau), float(rate),
                                       u12.data_ptr() if u12 is not None else None,
                                       u21p.data_ptr() if u21p is not None else None,
                                       seed, off, thr, rng_state.data_ptr() if (rng == "device" and rate > 0) else None,
                                       flags, q.data_ptr(), xent.data_ptr(), acc.data_ptr(),
                                       grad.data_ptr() if grad is not None else None, ws.data_ptr(), ws.numel(), _stream()),
                    "walk_fwd_bwd")
        else:
            # teacher-student call (teacherstudent.py:472-580): `ts` carries the teacher's chain products and / or asks for ours
            teacher = ts.get("teacher")
            if teacher is not None:
                teacher = _f32c(teacher)
                if tuple(teacher.shape) != (B, nw, N, N):
                    raise ValueError("teacher chains must be (B,T-2,N,N) = %s, got %s" % ((B, nw, N, N), tuple(teacher.shape)))
                ts["ts_xent"] = torch.zeros(nw + 1, dtype=torch.float32, device=dev)
            if ts.get("want_chains"):
                ts["chains"] = torch.empty(B, nw, N, N, dtype=torch.float32, device=dev)
            L.check(L.crw_walk_ts_fwd_bwd(feats.data_ptr(), B, N, T, D, float(tau), float(rate),
                                          u12.data_ptr() if u12 is not None else None,
                                          u21p.data_ptr() if u21p is not None else None,
                                          seed, off, thr, rng_state.data_ptr() if (rng == "device" and rate > 0) else None,
                                          flags, teacher.data_ptr() if teacher is not None else None, float(ts.get("alpha", 1.0)),
                                          ts["chains"].data_ptr() if ts.get("want_chains") else None,
                                          q.data_ptr(), xent.data_ptr(), ts["ts_xent"].data_ptr() if teacher is not None else None,
                                          acc.data_ptr(), grad.data_ptr() if grad is not None else None, ws.data_ptr(), ws.numel(),
                                          _stream()), "walk_ts_fwd_bwd")
        loss = xent[nw:nw + 1]                                 # model.py:413 (zeros when there is no walk)
        xent, acc = xent[:nw], acc[:nw]
        ctx.save_for_backward(grad if grad is not None else torch.empty(0, device=dev), q, feats)
        ctx.has_grad = grad is not None
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(xent, acc)
        return q, loss, xent, acc

    @staticmethod
    def backward(ctx, gq, gloss, gxent, gacc):
        grad, q, feats = ctx.saved_tensors
        out = None
        if ctx.has_grad and gloss is not None:
            out = grad * gloss.reshape(())                     # d loss / d feats was produced by the forward launch
        if gq is not None:
            # gradient arriving through the returned node embeddings (not used by the reference's training loop):
            # normalisation backward with the l2norm kernels
            L = _lib.lib()
            D = feats.shape[-1]
            rows = feats.numel() // D
            inv = torch.empty(rows, dtype=torch.float32, device=feats.device)
            nrm = torch.empty_like(inv)
            tmp = torch.empty_like(feats)
            L.check(L.crw_l2norm_fwd(feats.data_ptr(), tmp.data_ptr(), inv.data_ptr(), nrm.data_ptr(), rows, D, _stream()), "l2norm_fwd")
            extra = gq.contiguous().clone()
            L.check(L.crw_l2norm_bwd(q.data_ptr(), extra.data_ptr(), inv.data_ptr(), nrm.data_ptr(), rows, D, _stream()), "l2norm_bwd")
            out = extra if out is None else out + extra
        if out is None:
            out = torch.zeros_like(feats)
        return out, None, None, None, None, None, None, None, None


def walk(feats: torch.Tensor, temperature: float, rate: float, flip: bool = False, softmax: bool = False,
         rng: str = "philox", u12: Optional[torch.Tensor] = None, u21p: Optional[torch.Tensor] = None,
         force_general: bool = False, rng_state: Optional[torch.Tensor] = None, force_simt: bool = False, force_tc: bool = False, no_cluster: bool = False, no_tf32: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """feats (B,N,T,D) pre-normalisation node vectors -> (q (B,N,T,D) unit-norm, loss [1], xent (T-2), acc (T-2)).

    One launch computes the forward AND d loss / d feats; backward() only scales it by the incoming gradient.
    `rng`: 'philox' replays torch's CUDA generator inside the kernel (advancing it exactly as the reference's
    2(T-1) rand_like calls would); 'torch' draws the uniforms with torch.rand and hands them to the kernel;
    'device' reads {seed, offset} from the CUDA int64 tensor `rng_state` and advances it on the device (CUDA-graph
    safe: every replay draws fresh masks); explicit (u12, u21p) override all of these.
    """
    flags = (WALK_FLIP if flip else 0) | (WALK_SOFTMAX if softmax else 0) | (WALK_FORCE_GENERAL if force_general else 0)
    flags |= WALK_FORCE_SIMT if force_simt else 0
    flags |= WALK_FORCE_TC if force_tc else 0
    flags |= WALK_NO_CLUSTER if no_cluster else 0
    flags |= WALK_NO_TF32 if no_tf32 else 0
    if u12 is not None:
        rng = "torch"
    return _Walk.apply(feats, float(temperature), float(rate), flags, rng, u12, u21p, rng_state, None)


def walk_chains(feats: torch.Tensor, temperature: float, rate: float = 0.0, flip: bool = False, softmax: bool = True,
                rng: str = "philox", u12: Optional[torch.Tensor] = None, u21p: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The teacher side of teacherstudent.py:523-538: feats (B,N,T,D) -> the palindrome chain products (B,T-2,N,N) of walks
    1..T-2 (`aar`, or `aal` with flip); no gradient.  Edge dropout as in `walk` (the reference drops the teacher's edges too:
    stoch_mat's do_dropout defaults to True at teacherstudent.py:527-528)."""
    flags = (WALK_FLIP if flip else 0) | (WALK_SOFTMAX if softmax else 0) | WALK_FORCE_GENERAL
    if u12 is not None:
        rng = "torch"
    ts = {"want_chains": True}
    with torch.no_grad():
        _Walk.apply(feats.detach(), float(temperature), float(rate), flags, rng, u12, u21p, None, ts)
    return ts["chains"]


def walk_teacher_student(feats: torch.Tensor, teacher_feats: torch.Tensor, temperature: float, rate: float, alpha: float,
                         flip: bool = False, softmax: bool = True, rng: str = "philox", uniforms=None):
    """teacherstudent.py:494-577 from the node vectors on: feats / teacher_feats (B,N,T,D) before normalisation.  The walk
    loss of `walk` blended with the soft cross-entropy of the teacher's chain products against log_softmax of the student's
    (SoftCrossEntropyLoss, :270-292): loss = alpha * walk loss + (1 - alpha) * teacher-student loss.
    `uniforms` = (us12, us21p, ut12, ut21p) supplies the dropout draws; otherwise they are made in the reference's order
    (the student's 2(T-1) draws, then the teacher's) by torch.rand ('torch') or by replaying the generator in the kernels
    ('philox').  -> (q, loss [1], xent (T-2), acc (T-2), ts_xent (T-2)); the gradient of `loss` with respect to `feats` comes
    out of the same launch."""
    _need_cuda(feats, teacher_feats)
    B, N, T, D = feats.shape
    flags = (WALK_FLIP if flip else 0) | (WALK_SOFTMAX if softmax else 0) | WALK_FORCE_GENERAL
    us12 = us21p = ut12 = ut21p = None
    dropping = rate > 0 and T > 1
    if uniforms is not None:
        us12, us21p, ut12, ut21p = uniforms
        rng = "torch"
    elif dropping and rng == "torch":
        us12, us21p, ut12, ut21p = (torch.stack([torch.rand(B, N, N, device=feats.device) for _ in range(T - 1)]) for _ in range(4))
    if dropping and rng == "philox":
        # the teacher's chains are needed first but its draws come second: run it on the later part of the generator stream
        idx = feats.device.index if feats.device.index is not None else torch.cuda.current_device()
        gen = torch.cuda.default_generators[idx]
        off = gen.get_offset()
        span = 2 * (T - 1) * torch_rand_offset_increment(B * N * N, torch_rand_threads(B * N * N, idx))
        try:
            gen.set_offset(off + span)
            teacher = walk_chains(teacher_feats, temperature, rate, flip, softmax, "philox")
        finally:
            gen.set_offset(off)
    elif rng in ("philox", "torch"):
        teacher = walk_chains(teacher_feats, temperature, rate if dropping else 0.0, flip, softmax, rng, ut12, ut21p)
    else:
        raise ValueError("rng must be 'torch' or 'philox'")
    ts = {"teacher": teacher, "alpha": float(alpha)}
    q, loss, xent, acc = _Walk.apply(feats, float(temperature), float(rate), flags, rng, us12, us21p, None, ts)
    if dropping and rng == "philox":
        gen.set_offset(off + 2 * span)
    return q, loss, xent, acc, ts["ts_xent"][:-1]


# ------------------------------------------------------------------------------------------------------------------
# a4 / a5 as standalone operators (CRW.affinity, CRW.stoch_mat)
# ------------------------------------------------------------------------------------------------------------------
def _affinity_raw(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """x1 (R,N1,D), x2 (R,N2,D) contiguous -> (R,N1,N2)."""
    R, N1, D = x1.shape
    N2 = x2.shape[1]
    out = torch.empty(R, N1, N2, dtype=torch.float32, device=x1.device)
    L = _lib.lib()
    L.check(L.crw_affinity(x1.data_ptr(), x2.data_ptr(), R, N1, N2, D, out.data_ptr(), _stream()), "affinity")
    return out


class _Affinity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2):
        _need_cuda(x1, x2)
        check_device(x1.device)
        x1, x2 = _f32c(x1), _f32c(x2)
        ctx.save_for_backward(x1, x2)
        return _affinity_raw(x1, x2)

    @staticmethod
    def backward(ctx, g):
        x1, x2 = ctx.saved_tensors
        g = _f32c(g)
        # dX1 = G X2 = G (X2^T)^T ; dX2 = G^T X1
        g1 = _affinity_raw(g, x2.transpose(1, 2).contiguous())
        g2 = _affinity_raw(g.transpose(1, 2).contiguous(), x1.transpose(1, 2).contiguous())
        return g1, g2


def affinity_nodes(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """node-major affinity: x1 (R,N1,D), x2 (R,N2,D) -> (R,N1,N2) = x1 x2^T (differentiable)."""
    return _Affinity.apply(x1, x2)


def bmm_tc(A: torch.Tensor, B: torch.Tensor, trans_a: bool = False, trans_b: bool = False,
           out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """Batched fp32 GEMM on the tensor cores (the large-graph walk's contraction engine, gemm_tc.cu):
    A (Z,M,K) [or (Z,K,M) if trans_a], B (Z,K,N) [or (Z,N,K) if trans_b] -> (Z,M,N).  Needs M >= 128, N, K >= 64."""
    _need_cuda(A, B)
    check_device(A.device)
    A, B = _f32c(A), _f32c(B)
    Z = A.shape[0]
    M, K = (A.shape[2], A.shape[1]) if trans_a else (A.shape[1], A.shape[2])
    N = B.shape[1] if trans_b else B.shape[2]
    if (B.shape[2] if trans_b else B.shape[1]) != K or B.shape[0] != Z:
        raise ValueError("bmm_tc: shape mismatch %s x %s" % (tuple(A.shape), tuple(B.shape)))
    if out is None:
        if accumulate:
            raise ValueError("bmm_tc: accumulate needs `out`")
        out = torch.empty(Z, M, N, device=A.device, dtype=torch.float32)
    L = _lib.lib()
    nbytes = L.crw_bmm_tc_workspace_bytes(Z, M, N, K)
    ws = _workspace("bmm_tc", nbytes, A.device)
    L.check(L.crw_bmm_tc(A.data_ptr(), B.data_ptr(), out.data_ptr(), Z, M, N, K, int(trans_a), int(trans_b), int(accumulate),
                         ws.data_ptr(), ws.numel(), _stream()), "bmm_tc")
    if int(ws[:4].view(torch.int32)[0]) != 0:
        raise RuntimeError("bmm_tc: tensor-core pipeline timed out (error flag %d)" % int(ws[:4].view(torch.int32)[0]))
    return out


def bmm_tf32(A: torch.Tensor, B: torch.Tensor, trans_a: bool = False, trans_b: bool = False,
             out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """Batched fp32 GEMM on tcgen05 kind::tf32 with in-kernel operand splitting (gemm_tf32.cu); same conventions as bmm_tc.
    Needs M, N >= 64, K >= 32 and dimensions that are multiples of 4 (TMA reads the operands in place)."""
    _need_cuda(A, B)
    check_device(A.device)
    A, B = _f32c(A), _f32c(B)
    Z = A.shape[0]
    M, K = (A.shape[2], A.shape[1]) if trans_a else (A.shape[1], A.shape[2])
    N = B.shape[1] if trans_b else B.shape[2]
    if (B.shape[2] if trans_b else B.shape[1]) != K or B.shape[0] != Z:
        raise ValueError("bmm_tf32: shape mismatch %s x %s" % (tuple(A.shape), tuple(B.shape)))
    if out is None:
        if accumulate:
            raise ValueError("bmm_tf32: accumulate needs `out`")
        out = torch.empty(Z, M, N, device=A.device, dtype=torch.float32)
    err = torch.zeros(1, dtype=torch.int32, device=A.device)
    L = _lib.lib()
    L.check(L.crw_bmm_tf32(A.data_ptr(), B.data_ptr(), out.data_ptr(), Z, M, N, K, int(trans_a), int(trans_b), int(accumulate),
                           err.data_ptr(), _stream()), "bmm_tf32")
    if int(err[0]) != 0:
        raise RuntimeError("bmm_tf32: tensor-core pipeline timed out (error flag %d)" % int(err[0]))
    return out


def stoch_mat_(A: torch.Tensor, temperature: float, rate: float = 0.0, softmax: bool = False,
               uniform: Optional[torch.Tensor] = None) -> torch.Tensor:
    """model.py:74-90 on a contiguous (..., N, M) stack: in-place dropout to -1e20 where uniform < rate, then
    ZeroSoftmax (or softmax) of A / temperature along the last dim.  Forward only."""
    _need_cuda(A, uniform)
    check_device(A.device)
    if A.dtype != torch.float32 or not A.is_contiguous():
        raise ValueError("stoch_mat_ needs a contiguous fp32 tensor")
    N, M = A.shape[-2:]
    R = A.numel() // (N * M)
    out = torch.empty_like(A)
    if uniform is not None:
        uniform = _f32c(uniform)
    L = _lib.lib()
    L.check(L.crw_stoch_mat(A.data_ptr(), uniform.data_ptr() if uniform is not None else None, float(rate),
                            float(temperature), WALK_SOFTMAX if softmax else 0, R, N, M, out.data_ptr(), _stream()), "stoch_mat")
    return out


class _StochMat(torch.autograd.Function):
    """model.py:74-90 as a differentiable operator: forward = crw_stoch_mat (the in-place dropout lands in the caller's
    tensor, as in the reference), backward = crw_stoch_mat_bwd (dropped entries get no gradient)."""

    @staticmethod
    def forward(ctx, A, temperature, rate, softmax, uniform):
        _need_cuda(A, uniform)
        check_device(A.device)
        if A.dtype != torch.float32:
            raise TypeError("crw_b200 computes in fp32 (got %s)" % A.dtype)
        work = A.detach()
        aliased = work.is_contiguous()
        if not aliased:
            work = work.contiguous()                   # a copy: the side effect is written back below
        out = stoch_mat_(work, temperature, rate, softmax, uniform)
        if not aliased and rate > 0 and uniform is not None:
            A.detach().copy_(work)                     # propagate the in-place side effect through the (strided) view
        ctx.save_for_backward(work, out)
        ctx.cfg = (float(temperature), bool(softmax))
        return out

    @staticmethod
    def backward(ctx, g):
        work, out = ctx.saved_tensors
        tau, softmax = ctx.cfg
        g = _f32c(g)
        N, M = work.shape[-2:]
        gA = torch.empty_like(work)
        L = _lib.lib()
        L.check(L.crw_stoch_mat_bwd(work.data_ptr(), out.data_ptr(), g.data_ptr(), tau, WALK_SOFTMAX if softmax else 0,
                                    work.numel() // (N * M), N, M, gA.data_ptr(), _stream()), "stoch_mat_bwd")
        return gA, None, None, None, None


def stoch_mat(A: torch.Tensor, temperature: float, rate: float = 0.0, softmax: bool = False,
              uniform: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Differentiable model.py:74-90 on a (..., N, M) tensor (any strides): in-place dropout to -1e20 where uniform < rate
    (written into `A`, like the reference), then ZeroSoftmax / softmax of A / temperature along the last dim."""
    return _StochMat.apply(A, float(temperature), float(rate), bool(softmax), uniform)


# ------------------------------------------------------------------------------------------------------------------
# a9-a12: label propagation
# ------------------------------------------------------------------------------------------------------------------
def lp_prepare(feats_cf: torch.Tensor, normalize: bool) -> torch.Tensor:
    """(C, Nf, hw) channel-first -> (Nf, hw, C) channel-last, optionally L2-normalised over C (test.py:93)."""
    _need_cuda(feats_cf)
    check_device(feats_cf.device)
    feats_cf = _f32c(feats_cf)
    C, Nf, hw = feats_cf.shape
    out = torch.empty(Nf, hw, C, dtype=torch.float32, device=feats_cf.device)
    L = _lib.lib()
    L.check(L.crw_lp_prepare(feats_cf.data_ptr(), C, Nf, hw, 1 if normalize else 0, out.data_ptr(), _stream()), "lp_prepare")
    return out


def lp_topk(feats_cl: torch.Tensor, key_frames: torch.Tensor, query_frames: torch.Tensor, n_long: int, h: int, w: int,
            radius: float, temperature: float, k: int, dense_mask: Optional[torch.Tensor] = None, force_simt: bool = False,
            check: bool = True, exact_only: bool = False, stats: Optional[dict] = None):
    """feats_cl (Nf,hw,C); key_frames (Nt,S) int64; query_frames (Nt) int64 -> Ws (Nt,k,hw) fp32, Is (Nt,k,hw) int64.

    Tensor-core path (C % 64 == 0, C <= 256, k <= 12, radius <= 12, no dense mask): fp16 pre-ranking on tcgen05, exact fp32
    re-ranking of a 16-key shortlist per query, certification, and the fp32-faithful 3-MMA pass for the tiles that could not
    be certified (`exact_only` sends every tile there).  `stats`, if given, receives {"tensor_cores", "tiles", "listed_tiles",
    "uncertified_queries"} (needs `check`, which reads the workspace header back and therefore synchronises)."""
    _need_cuda(feats_cl, key_frames, query_frames, dense_mask)
    check_device(feats_cl.device)
    feats_cl = _f32c(feats_cl)
    key_frames = key_frames.to(torch.int64).contiguous()
    query_frames = query_frames.to(torch.int64).contiguous()
    Nf, hw, C = feats_cl.shape
    if hw != h * w:
        raise ValueError("feats have %d positions, expected %d x %d" % (hw, h, w))
    Nt, S = key_frames.shape
    if Nt and (int(key_frames.max()) >= Nf or int(query_frames.max()) >= Nf or int(key_frames.min()) < 0):
        raise ValueError("frame index out of range")
    if dense_mask is not None:
        dense_mask = _f32c(dense_mask).reshape(hw, hw)
    dev = feats_cl.device
    Ws = torch.empty(Nt, k, hw, dtype=torch.float32, device=dev)
    Is = torch.empty(Nt, k, hw, dtype=torch.int64, device=dev)
    L = _lib.lib()
    on_tc = bool(L.crw_lp_topk_uses_tensor_cores(C, k, float(radius), int(dense_mask is not None))) and not force_simt
    if not force_simt and not on_tc:
        _warn_once(("lp_simt", C, k, dense_mask is not None),
                   "label propagation with C=%d, k=%d%s runs on the exact-fp32 SIMT kernel (tensor-core kernel: C %% 64 == 0, "
                   "C <= 256, k <= 12, radius <= 12, no dense mask); expect ~10x lower throughput"
                   % (C, k, ", dense mask" if dense_mask is not None else ""))
    nbytes = L.crw_lp_topk_workspace_bytes(Nf, Nt, S, h, w, C, k)
    ws = _workspace(("lp", Nf, Nt, h, w, C, k), nbytes, dev)
    flags = (_lib.LP_FORCE_SIMT if force_simt else 0) | (_lib.LP_EXACT_ONLY if exact_only else 0)
    L.check(L.crw_lp_topk(feats_cl.data_ptr(), Nf, key_frames.data_ptr(), query_frames.data_ptr(), Nt, S, n_long, h, w, C,
                          float(radius), dense_mask.data_ptr() if dense_mask is not None else None, float(temperature), k,
                          flags, Ws.data_ptr(), Is.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "lp_topk")
    if check:
        hdr = ws[:32].view(torch.int32).tolist()          # one small read-back: error word + certification counters
        if hdr[0] == 2:
            raise _lib.CrwError("lp_topk: features exceed the fp16 range of the tensor-core path (|x| >= 6e4); L2-normalise "
                                "them (test.py:93) or pass force_simt=True")
        if hdr[0] != 0:
            raise _lib.CrwError("lp_topk: tensor-core kernel reported an internal barrier timeout")
        if stats is not None:
            tiles = ((w + 7) // 8) * ((h + 15) // 16) * Nt
            unc = hdr[4] if on_tc else 0
            stats.update(tensor_cores=on_tc, tiles=tiles, listed_tiles=hdr[1] if on_tc else 0, uncertified_queries=unc,
                         settled_by=("nothing to settle" if unc == 0 else "exact fp32 evaluation per query" if unc <= 1024
                                     else "fp32-faithful tensor-core pass over the listed tiles") if on_tc and not exact_only
                         else ("fp32-faithful tensor-core pass over every tile" if on_tc else "SIMT kernel"))
    return Ws, Is


def lp_gather_(lbls: torch.Tensor, key_frames_n: torch.Tensor, Ws_n: torch.Tensor, Is_n: torch.Tensor, out_frame: int) -> None:
    """One step of test.py:147-157 in place: lbls (Nf,hw,L); writes lbls[out_frame]."""
    _need_cuda(lbls, key_frames_n, Ws_n, Is_n)
    if lbls.dtype != torch.float32 or not lbls.is_contiguous():
        raise ValueError("lbls must be contiguous fp32 (Nf,hw,L)")
    Nf, hw, Lc = lbls.shape
    k = Ws_n.shape[0]
    L = _lib.lib()
    L.check(L.crw_lp_gather(lbls.data_ptr(), key_frames_n.contiguous().data_ptr(), _f32c(Ws_n).data_ptr(),
                            Is_n.contiguous().data_ptr(), hw, Lc, k, int(out_frame), _stream()), "lp_gather")


def lp_gather_all_(lbls: torch.Tensor, key_frames: torch.Tensor, Ws: torch.Tensor, Is: torch.Tensor, first_target: int, out_frame0: int) -> None:
    """test.py:145-157 for every target from `first_target` on, in place, one launch: lbls (Nf,hw,L); key_frames (Nt,S); Ws / Is
    (Nt,k,hw); target t writes lbls[out_frame0 + t]."""
    _need_cuda(lbls, key_frames, Ws, Is)
    if lbls.dtype != torch.float32 or not lbls.is_contiguous():
        raise ValueError("lbls must be contiguous fp32 (Nf,hw,L)")
    Nf, hw, Lc = lbls.shape
    Nt, k = Ws.shape[0], Ws.shape[1]
    L = _lib.lib()
    L.check(L.crw_lp_gather_all(lbls.data_ptr(), key_frames.to(torch.int64).contiguous().data_ptr(), _f32c(Ws).data_ptr(),
                                Is.contiguous().data_ptr(), Nt, key_frames.shape[1], hw, Lc, k, int(first_target), int(out_frame0), _stream()),
            "lp_gather_all")


def lp_minmax_normalize_(maps: torch.Tensor) -> torch.Tensor:
    """test.py:162-164 (--norm_mask) in place on a contiguous (..., L) fp32 tensor: rows -= min; rows /= max."""
    _need_cuda(maps)
    if maps.dtype != torch.float32 or not maps.is_contiguous():
        raise ValueError("lp_minmax_normalize_ needs a contiguous fp32 tensor")
    Lc = maps.shape[-1]
    L = _lib.lib()
    L.check(L.crw_lp_minmax_normalize(maps.data_ptr(), maps.numel() // Lc, Lc, _stream()), "lp_minmax_normalize")
    return maps


def lp_upsample_argmax(preds: torch.Tensor, size: Tuple[int, int], palette: Optional[torch.Tensor] = None,
                       norm_mask: bool = False, want_rgb: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """test_utils.py:85-123 on the device: preds (n,h,w,L) or (h,w,L) soft label maps -> (cls (n,H,W) uint8, rgb (n,H,W,3)
    uint8 = palette[cls]) with OpenCV's bilinear resize rule and numpy's first-maximum arg-max; palette (L,3) integer (any
    device: it is a few bytes)."""
    _need_cuda(preds)
    check_device(preds.device)
    if preds.dim() == 3:
        preds = preds[None]
    preds = _f32c(preds)
    n, h, w, Lb = preds.shape
    H, W = int(size[0]), int(size[1])
    cls = torch.empty(n, H, W, dtype=torch.uint8, device=preds.device)
    rgb = torch.empty(n, H, W, 3, dtype=torch.uint8, device=preds.device) if want_rgb else None
    pal = None
    if palette is not None:
        if palette.shape != (Lb, 3):
            raise ValueError("palette must be (L,3), got %s" % (tuple(palette.shape),))
        pal = palette.to(device=preds.device).to(torch.int64).to(torch.uint8).contiguous()      # np.uint8(lbl_set): wraps modulo 256
    L = _lib.lib()
    L.check(L.crw_lp_upsample_argmax(preds.data_ptr(), n, h, w, Lb, H, W, int(norm_mask), pal.data_ptr() if pal is not None else None,
                                     cls.data_ptr(), rgb.data_ptr() if rgb is not None else None, _stream()), "lp_upsample_argmax")
    return cls, rgb


def lp_pose_coords(preds: torch.Tensor, topk: int = 3) -> torch.Tensor:
    """utils/test_utils.py:60-84 (process_pose) on the device: preds (n,h,w,L) or (h,w,L) soft label maps, channel 0 =
    background -> key-point coordinates (n,2,L-1) float32 ((x, y) rows; -1 for channels that are zero everywhere)."""
    _need_cuda(preds)
    check_device(preds.device)
    if preds.dim() == 3:
        preds = preds[None]
    preds = _f32c(preds)
    n, h, w, Lb = preds.shape
    coords = torch.empty(n, 2, max(Lb - 1, 0), dtype=torch.float32, device=preds.device)
    L = _lib.lib()
    L.check(L.crw_lp_pose_coords(preds.data_ptr(), n, h, w, Lb, int(topk), coords.data_ptr(), _stream()), "lp_pose_coords")
    return coords


def sinkhorn_knopp(A: torch.Tensor, tol: float = 0.01, max_iter: int = 1000, exp_temperature: Optional[float] = None):
    """utils/__init__.py:615-641: A (R,N,M) or (N,M) positive -> the Sinkhorn-Knopp normalised matrix (new tensor, no
    gradient) and the number of sweeps.  `exp_temperature` = tau first maps A -> exp(A / tau) (the argument stoch_mat passes
    at model.py:84).  The stop rule is the reference's (std of all column sums > tol) and is read back once per sweep."""
    import ctypes
    _need_cuda(A)
    check_device(A.device)
    squeeze = A.dim() == 2
    work = _f32c(A.detach()).clone()
    if squeeze:
        work = work[None]
    R, N, M = work.shape
    L = _lib.lib()
    nbytes = L.crw_sinkhorn_workspace_bytes(R, N, M)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=work.device)
    n_it = ctypes.c_int(0)
    L.check(L.crw_sinkhorn_knopp(work.data_ptr(), R, N, M, int(exp_temperature is not None), float(exp_temperature or 1.0), float(tol),
                                 int(max_iter), ctypes.addressof(n_it), ws.data_ptr(), nbytes, _stream()), "sinkhorn_knopp")
    return (work[0] if squeeze else work), n_it.value
