"""crw_b200: the contrastive random walk (CRW) hot path and its label-propagation evaluator as hand-written sm_100a
CUDA kernels behind the reference's Python API (paolomandica/sapienza-video-contrastive, code/model.py and
code/utils/test_utils.py).  See DESIGN.md for the scope and INTEGRATION.md for the C ABI.

Importing this package never touches the GPU; the first operator call loads libcrw_b200.so (built in-tree by
`_lib.build()` / `__graft_entry__.build()`) and fails loudly if it is missing - there is no fallback path.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"

_LAZY = {
    "CRW": ("model", "CRW"),
    "ZeroSoftmax": ("model", "ZeroSoftmax"),
    "CRWBase": ("teacherstudent", "CRWBase"),
    "CRWTeacherStudent": ("teacherstudent", "CRWTeacherStudent"),
    "SoftCrossEntropyLoss": ("teacherstudent", "SoftCrossEntropyLoss"),
    "make_encoder": ("resnet", "make_encoder"),
    "From3D": ("resnet", "From3D"),
    "context_index_bank": ("test_utils", "context_index_bank"),
    "mem_efficient_batched_affinity": ("test_utils", "mem_efficient_batched_affinity"),
    "batched_affinity": ("test_utils", "batched_affinity"),
    "MaskedAttention": ("test_utils", "MaskedAttention"),
    "RadiusMask": ("test_utils", "RadiusMask"),
    "LabelPropagator": ("test_utils", "LabelPropagator"),
    "propagate_labels": ("test_utils", "propagate_labels"),
    "dump_predictions": ("test_utils", "dump_predictions"),
    "davis_index_maps": ("test_utils", "davis_index_maps"),
    "process_pose": ("test_utils", "process_pose"),
}


def __getattr__(name):
    if name in ("ops", "model", "test_utils", "resnet", "teacherstudent"):
        import importlib
        return importlib.import_module("." + name, __name__)
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module("." + mod, __name__), attr)
    raise AttributeError(name)
