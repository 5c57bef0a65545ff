"""TEST INFRASTRUCTURE ONLY -- ctypes handle of oracle/_ref/libbow_ref.so: DBoW2's own vocabulary-tree descent,
FORB::distance, BowVector.cpp and FeatureVector.cpp as vendored in the reference, compiled where they lie
(oracle/ref_shim/bow).  Call through oracle_py.Oracle.bow_transform(voc, desc, levelsup, lib=ref_bow.lib(),
fn="ref_bow_transform")."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libbow_ref.so")


def build():
    """Needs /root/reference (absent on the GPU box, where the prebuilt file is used)."""
    subprocess.check_call(["make", "-s", "-f", os.path.join(HERE, "ref_shim", "Makefile"), LIB])
    return LIB


def available():
    return os.path.exists(LIB)


def lib():
    return C.CDLL(LIB)
