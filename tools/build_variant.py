#!/usr/bin/env python3
"""A/B aid: builds gpurun_out/variants/liborbb200_<name>.so from the current objects with ONE source recompiled under extra
flags (usage: build_variant.py <name> <source.cu> <flag> [<flag> ...]).  tools/gpu_variants.sh swaps the variants in on
the GPU box."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vi-orb-slam-icra2018_b200")
sys.path.insert(0, PKG)
import build as B
B.build()
name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
out = os.path.join(PKG, "variants")
os.makedirs(out, exist_ok=True)
obj = os.path.join(out, "%s_%s.o" % (os.path.basename(src)[:-3], name))
subprocess.check_call([B.NVCC] + B.FLAGS + flags + ["-c", os.path.join(B.CSRC, src), "-o", obj], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
objs = [obj if os.path.basename(s) == src else os.path.join(B.OBJ, os.path.basename(s)[:-3] + ".o") for s in B._sources()]
lib = os.path.join(out, "liborbb200_%s.so" % name)
subprocess.check_call([B.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-cudart", "static"])
print(lib)
