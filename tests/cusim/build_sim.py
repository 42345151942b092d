"""Build libcrw_b200_sim.so: the product's .cu sources compiled for the HOST with cuda_sim.h force-included.
TEST INFRASTRUCTURE ONLY - see cuda_sim.h.  The product package never loads this library."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sapienza_video_contrastive_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libcrw_b200_sim.so")


def build(force: bool = False, opt: str = "-O1") -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu"))) + [os.path.join(HERE, "cuda_sim.cc")]
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "cuda_sim.h"),
                                                            os.path.join(ROOT, "include", "crw_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(os.path.dirname(OUT), os.path.basename(s) + ".o")
        objs.append(o)
        cmd = ["g++", "-std=c++17", opt, "-g", "-fPIC", "-DCRW_SIM", "-Wno-unknown-pragmas", "-Wno-attributes",
               "-include", os.path.join(HERE, "cuda_sim.h"), "-I", HERE, "-x", "c++", "-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    bad = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            bad = True
            sys.stderr.write("---- %s\n%s\n" % (s, out))
    if bad:
        raise RuntimeError("sim build failed")
    subprocess.check_call(["g++", "-shared", "-o", OUT] + objs)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
