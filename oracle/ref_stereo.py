"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libstereo_ref.so: the reference's own text of
Frame::ComputeStereoMatches (Frame.cc:810-984), compiled where it lies (oracle/ref_shim/stereo), behind the same
flat-array call as oracle_py.Oracle.stereo."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle_py import KP_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libstereo_ref.so")


def build():
    """Needs /root/reference (absent on the GPU box, where the prebuilt file is used)."""
    subprocess.check_call(["make", "-s", "-f", os.path.join(HERE, "ref_shim", "Makefile"), LIB])
    return LIB


def available():
    return os.path.exists(LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def stereo(keys_l, desc_l, keys_r, desc_r, levels_l, levels_r, scale, inv_scale, mb, mbf):
    """-> (mvuRight, mvDepth, kept) as the reference's own loop computes them"""
    lib = C.CDLL(LIB)
    lib.ref_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p,
                               C.c_void_p]
    kl = np.ascontiguousarray(keys_l, KP_DTYPE); kr = np.ascontiguousarray(keys_r, KP_DTYPE)
    dl = np.ascontiguousarray(desc_l, np.uint8); dr = np.ascontiguousarray(desc_r, np.uint8)
    nl = len(levels_l)
    la = [np.ascontiguousarray(a, np.uint8) for a in levels_l]
    ra = [np.ascontiguousarray(a, np.uint8) for a in levels_r]
    pl = (C.c_void_p * nl)(*[a.ctypes.data for a in la])
    pr = (C.c_void_p * nl)(*[a.ctypes.data for a in ra])
    cols = np.array([a.shape[1] - 38 for a in la], np.int32)
    rows = np.array([a.shape[0] - 38 for a in la], np.int32)
    sf = np.ascontiguousarray(scale, np.float32); isf = np.ascontiguousarray(inv_scale, np.float32)
    ur = np.empty(len(kl), np.float32); depth = np.empty(len(kl), np.float32)
    kept = lib.ref_stereo(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), pl, pr, _p(cols), _p(rows), nl, _p(sf), _p(isf),
                          C.c_float(mb), C.c_float(mbf), _p(ur), _p(depth))
    return ur, depth, kept
