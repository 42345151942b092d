"""Recipe for oracle/_ref/: makes the UNMODIFIED reference runnable on the GPU box.  TEST / BENCH INFRASTRUCTURE.

The reference is pure Python (no compiled sources): "building" it means placing its own files, byte for byte, under the
git-ignored directory oracle/_ref/code/ so that they travel to the GPU box with the snapshot (oracle/_ref/ is listed in
.gitignore but not in .gpurunignore).  Nothing is copied into the tracked tree; nothing here is imported by the product.

    python -m oracle.build_ref            # in the authoring container, where /root/reference exists

Used by: bench.py --impl reference and bench.py's cpu_baseline legs (kind "reference"), through oracle/ref_import.py,
which prefers /root/reference/code and falls back to oracle/_ref/code.
"""
from __future__ import annotations

import filecmp
import os
import shutil

SRC = "/root/reference/code"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "code")


def build(verbose: bool = True) -> int:
    """Mirror every *.py of the reference's code/ tree into oracle/_ref/code/.  Returns the number of files in place
    (0 when the reference is absent and nothing was built before)."""
    if not os.path.isfile(os.path.join(SRC, "model.py")):
        n = sum(len([f for f in fs if f.endswith(".py")]) for _, _, fs in os.walk(DST)) if os.path.isdir(DST) else 0
        if verbose:
            print("oracle/_ref: reference tree not present, keeping %d prebuilt files" % n)
        return n
    n = 0
    for root, _, files in os.walk(SRC):
        for f in files:
            if not f.endswith(".py"):
                continue
            s = os.path.join(root, f)
            d = os.path.join(DST, os.path.relpath(s, SRC))
            os.makedirs(os.path.dirname(d), exist_ok=True)
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
            n += 1
    if verbose:
        print("oracle/_ref: %d reference files in place under %s" % (n, DST))
    return n


if __name__ == "__main__":
    build()
