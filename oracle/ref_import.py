"""Import the UNMODIFIED reference from /root/reference/code (authoring container only).  TEST INFRASTRUCTURE.

The reference imports a handful of plotting / image packages at module import time that are
absent here and unused on the hot path; they are replaced by empty stub
modules.  Nothing from the reference is copied into this repository: this module only makes
`import model`, `import utils.test_utils` resolve so oracle/gen_golden.py can execute it.
/root/reference does not exist on the GPU box; there the byte-identical files that oracle/build_ref.py placed under the
git-ignored oracle/_ref/code/ are imported instead (bench.py's reference arm only - the -m gpu tests never need them).
"""
from __future__ import annotations

import argparse
import os
import sys
import types

REF_ROOT = "/root/reference/code"
REF_BUILT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "code")      # oracle/build_ref.py


def root() -> str:
    """/root/reference/code in the authoring container, else the byte-identical files placed by oracle/build_ref.py."""
    return REF_ROOT if os.path.isfile(os.path.join(REF_ROOT, "model.py")) else REF_BUILT


def available() -> bool:
    return os.path.isfile(os.path.join(root(), "model.py"))


class _Any:
    def __getattr__(self, k):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """-> (model_module, utils_module, test_utils_module) of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present")
    for name in ("skimage", "imageio", "visdom", "matplotlib", "plotly"):
        if name in sys.modules:
            continue
    sk = _stub("skimage")
    sk.util = _stub("skimage.util", img_as_float=lambda x: x, view_as_windows=None)
    _stub("skimage.segmentation", mark_boundaries=None, felzenszwalb=None, slic=None)
    _stub("imageio")
    _stub("visdom")
    _stub("matplotlib", cm=_Any())
    _stub("matplotlib.pyplot")
    _stub("matplotlib.cm")
    _stub("plotly")
    _stub("plotly.subplots", make_subplots=None)
    _stub("plotly.graph_objects")
    _stub("plotly.express")
    # the reference's top-level package is called `utils`; make sure ours / others do not shadow it
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.") or k in ("model", "resnet")]:
        del sys.modules[k]
    sys.path.insert(0, root())
    hook = sys.excepthook
    import model as ref_model            # noqa: E402
    import utils as ref_utils            # noqa: E402
    import utils.test_utils as ref_tu    # noqa: E402
    sys.excepthook = hook                # utils/__init__.py:40 installs a pdb post-mortem hook
    _loaded = (ref_model, ref_utils, ref_tu)
    return _loaded


def namespace(**over):
    """The Namespace fields CRW.__init__ reads (model.py:19-38)."""
    d = dict(device="cpu", dropout=0.1, featdrop=0.0, temp=0.07, head_depth=0, model_type="scratch",
             remove_layers=[], dilate_superpixels=False, flip=False, sk_targets=False)
    d.update(over)
    return argparse.Namespace(**d)
