"""ctypes binding of libcrw_b200.so (C ABI in include/crw_b200.h) and its in-tree nvcc build.

There is exactly one compute backend: the sm_100a CUDA library.  If it is missing or the device is not
a B200-class GPU the product raises; nothing here falls back to PyTorch or the CPU.
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess
import threading
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_PATH = os.path.join(PKG, "libcrw_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]

CRW_OK = 0
WALK_SOFTMAX = 1
WALK_FLIP = 2
WALK_FORCE_GENERAL = 4
WALK_FORCE_SIMT = 8
WALK_FORCE_TC = 16
WALK_NO_CLUSTER = 32
WALK_NO_TF32 = 64
LP_FORCE_SIMT = 1
LP_EXACT_ONLY = 2
DILATE_SHAPES = {"L1": 0, "circle": 1, "cross": 2}                # CRW_DILATE_* (utils/__init__.py:590-608 kernel names)

_SIGNATURES = {
    "crw_version": (c_int, []),
    "crw_last_error": (c_char_p, []),
    "crw_pool_patch_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "crw_pool_patch_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "crw_pool_patch_fwd_sm": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "crw_pool_patch_bwd_sm": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "crw_pool_patch_bwd_scaled": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_int, c_void_p]),
    "crw_segmean_workspace_bytes": (c_size_t, [c_int] * 7),
    "crw_segmean_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64] + [c_int] * 8
                        + [c_void_p, c_void_p, c_size_t, c_void_p]),
    "crw_segmean_bwd": (c_int, [c_void_p, c_void_p, c_size_t] + [c_int] * 8 + [c_void_p, c_void_p]),
    "crw_segmean_dilated_workspace_bytes": (c_size_t, [c_int] * 7),
    "crw_segmean_dilated_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64] + [c_int] * 10
                                + [c_void_p, c_void_p, c_size_t, c_void_p]),
    "crw_segmean_dilated_bwd": (c_int, [c_void_p, c_void_p, c_size_t] + [c_int] * 8 + [c_void_p, c_void_p]),
    "crw_affinity": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "crw_stoch_mat": (c_int, [c_void_p, c_void_p, c_float, c_float, c_uint32, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "crw_stoch_mat_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_uint32, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "crw_sinkhorn_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "crw_sinkhorn_knopp": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "crw_walk_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_uint32]),
    "crw_walk_fwd_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_uint64,
                                 c_uint64, c_uint32, c_void_p, c_uint32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "crw_walk_ts_fwd_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_uint64,
                                    c_uint64, c_uint32, c_void_p, c_uint32, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "crw_philox_uniform": (c_int, [c_void_p, c_int64, c_uint64, c_uint64, c_uint32, c_void_p]),
    "crw_lp_upsample_argmax": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "crw_lp_minmax_normalize": (c_int, [c_void_p, c_int64, c_int, c_void_p]),
    "crw_lp_pose_coords": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "crw_bmm_tc_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "crw_bmm_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "crw_bmm_tf32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "crw_lp_topk_workspace_bytes": (c_size_t, [c_int] * 7),
    "crw_lp_topk_uses_tensor_cores": (c_int, [c_int, c_int, c_float, c_int]),
    "crw_lp_topk": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p,
                            c_float, c_int, c_uint32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "crw_head_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "crw_head_dgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "crw_head_wgrad_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "crw_head_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "crw_head_wgrad_axpby": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "crw_patch_grid": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "crw_slic_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_void_p, c_int]),
    "crw_slic": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_double, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "crw_label_connectivity_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "crw_label_connectivity": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "crw_head_fwd_splitk_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "crw_head_fwd_splitk": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "crw_l2norm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "crw_l2norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "crw_lp_prepare": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "crw_lp_gather_all": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int64, c_void_p]),
    "crw_lp_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)


class CrwError(RuntimeError):
    pass


class CrwLib:
    """Typed handle on a built library.  `check(rc)` turns the C return code into an exception."""

    def __init__(self, path: str):
        self.path = path
        self._dll = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(self._dll, name)       # AttributeError if an export is missing: fail loudly
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, rc: int, what: str = "") -> None:
        if rc != CRW_OK:
            msg = self.crw_last_error()
            raise CrwError("%s failed (%d): %s" % (what or "crw call", rc, msg.decode() if msg else ""))


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


HASH_PATH = LIB_PATH + ".srchash"


def _source_hash() -> str:
    """Digest of everything the library is compiled from.  The library is current iff the digest written next to it at build time
    matches: modification times do not survive a copy of the tree to another machine in any useful order."""
    import hashlib
    deps = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(os.path.dirname(PKG), "include", "crw_b200.h")]
    h = hashlib.sha1()
    for d in deps:
        if os.path.exists(d):
            h.update(os.path.basename(d).encode())
            with open(d, "rb") as f:
                h.update(f.read())
    h.update(" ".join(ARCH_FLAGS).encode())
    return h.hexdigest()


def _stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every kernel for sm_100a into libcrw_b200.so next to this file (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    if not os.path.exists(NVCC):
        raise CrwError("nvcc not found at %s and %s is missing or stale" % (NVCC, LIB_PATH))
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    # one builder at a time across PROCESSES (the ranks of a torchrun job all import this module at once)
    import fcntl
    with open(os.path.join(objdir, ".lock"), "w") as lockf:
        fcntl.flock(lockf, fcntl.LOCK_EX)
        try:
            if not force and not _stale():          # another process built it while this one waited
                return LIB_PATH
            return _build_locked(objdir, verbose)
        finally:
            fcntl.flock(lockf, fcntl.LOCK_UN)


def _build_locked(objdir: str, verbose: bool) -> str:
    digest = _source_hash()
    procs, objs = [], []
    for s in sources():
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        cmd = ([NVCC, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + ARCH_FLAGS
               + os.environ.get("CRW_NVCC_EXTRA", "").split() + ["-c", s, "-o", o])
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        log.append("---- %s\n%s" % (os.path.basename(s), out))
        failed |= p.returncode != 0
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose or failed:
        print("\n".join(log))
    if failed:
        raise CrwError("nvcc failed, see log above")
    tmp = "%s.%d.tmp" % (LIB_PATH, os.getpid())
    subprocess.check_call([NVCC, "-shared", "-o", tmp] + ARCH_FLAGS + objs)
    os.replace(tmp, LIB_PATH)
    with open(HASH_PATH + ".tmp", "w") as f:
        f.write(digest + "\n")
    os.replace(HASH_PATH + ".tmp", HASH_PATH)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def lib() -> CrwLib:
    """The process-wide library handle.  Builds it if the sources are newer and nvcc is present."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if _stale() and os.path.exists(NVCC):
                    build()
                if not os.path.exists(LIB_PATH):
                    raise CrwError("libcrw_b200.so is not built (run `python __graft_entry__.py build`); "
                                   "there is no fallback path")
                _lib = CrwLib(LIB_PATH)
    return _lib
