"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libvocio_ref.so: DBoW2's own vocabulary file readers and
writers (TemplatedVocabulary::loadFromTextFile / loadFromBinaryFile / saveToTextFile / saveToBinaryFile, FORB::fromString /
toString) as vendored in the reference, compiled where they lie (oracle/ref_shim/bow)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libvocio_ref.so")


def build():
    """Needs /root/reference (absent on the GPU box, where the prebuilt file is used)."""
    subprocess.check_call(["make", "-s", "-f", os.path.join(HERE, "ref_shim", "Makefile"), LIB])
    return LIB


def available():
    return os.path.exists(LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def load(path, binary):
    """flat arrays of the tree the reference's reader builds from `path`, or None if it refuses the file"""
    lib = C.CDLL(LIB)
    lib.ref_vocab_load.argtypes = [C.c_char_p, C.c_int] + [C.c_void_p] * 7
    hdr = np.zeros(5, np.int32)
    n = lib.ref_vocab_load(path.encode(), int(binary), _p(hdr), None, None, None, None, None, None)
    if n == 0:
        return None
    desc = np.zeros((n, 32), np.uint8)
    parent = np.zeros(n, np.int32); cs = np.zeros(n + 1, np.int32); ch = np.zeros(max(n, 1), np.int32)
    wid = np.zeros(n, np.int32); w = np.zeros(n, np.float64)
    lib.ref_vocab_load(path.encode(), int(binary), _p(hdr), _p(desc), _p(parent), _p(cs), _p(ch), _p(wid), _p(w))
    return dict(n_nodes=n, k=int(hdr[0]), L=int(hdr[1]), scoring=int(hdr[2]), weighting=int(hdr[3]), desc=desc, parent=parent,
                child_start=cs, children=ch[:cs[n]].copy(), word_id=wid, weight=w, n_words=int(hdr[4]))


def resave(path, binary, out_text, out_binary):
    lib = C.CDLL(LIB)
    lib.ref_vocab_resave.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_char_p]
    return lib.ref_vocab_resave(path.encode(), int(binary), out_text.encode(), out_binary.encode())
