#!/usr/bin/env python3
"""Builds liborbb200.so in-tree: every csrc/*.cu compiled by nvcc for sm_100a and linked into one shared library.

    python vi-orb-slam-icra2018_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX for other targets, no multi-arch fat binary
  -lineinfo                                 so ncu --import-source maps SASS back to these files
  -fmad=false                               float results on the path are bit-exact contracts (no FMA contraction);
                                            the exact spots additionally use __fmul_rn/__fadd_rn intrinsics
"""
import argparse
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "liborbb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off", "-Xptxas", "-v", "--threads", "4"]
FLAGS += os.environ.get("ORBB_NVCC_EXTRA", "").split()   # e.g. -DORBB_FW_STATS: debug counters of the FAST kernel


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_digest():
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h", ".inc")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "orbb200.h"), "rb").read())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp_path = os.path.join(OBJ, "deps.sha256")
    digest = _deps_digest()
    old = open(stamp_path).read() if os.path.exists(stamp_path) else ""
    rebuild_all = force or old != digest
    objs, changed = [], False
    for src in _sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if rebuild_all or not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(src):
            cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            else:
                with open(obj + ".ptxas.log", "w") as f:
                    f.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on " + src)
            changed = True
    if changed or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("link failed")
    with open(stamp_path, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
