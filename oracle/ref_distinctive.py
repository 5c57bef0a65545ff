"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libdistinctive_ref.so: the reference's own text of
MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:257-322) compiled where it lies (oracle/ref_shim/stereo)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libdistinctive_ref.so")


def build():
    """Needs /root/reference (absent on the GPU box, where the prebuilt file is used)."""
    subprocess.check_call(["make", "-s", "-f", os.path.join(HERE, "ref_shim", "Makefile"), LIB])
    return LIB


def available():
    return os.path.exists(LIB)


def distinctive(desc, bad=None):
    """index (among the rows of good keyframes) of the descriptor the reference keeps, -1 if it keeps none"""
    lib = C.CDLL(LIB)
    lib.ref_distinctive.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    b = None if bad is None else np.ascontiguousarray(bad, np.uint8)
    return lib.ref_distinctive(d.ctypes.data_as(C.c_void_p), None if b is None else b.ctypes.data_as(C.c_void_p), len(d))
