"""The C-ABI library without a GPU: it loads, exports every symbol include/orbb200.h declares, and fails loudly
(status + message, never a CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

import orbb200
from orbb200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "orbb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(orb[xm]?_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("orbb200_build", os.path.join(ROOT, "vi-orb-slam-icra2018_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    return orbb200.lib()


def test_every_declared_symbol_is_exported(lib):
    names = header_functions()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_abi.SIGNATURES) == names          # the ctypes table mirrors the header one to one
    assert _abi.declare(lib) == []


def test_version_and_error_string(lib):
    assert lib.orb_version() == 100
    assert isinstance(lib.orb_last_error(), bytes)


def test_fails_loudly_without_a_device(lib):
    if lib.orb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    st = lib.orbm_create(0, C.byref(h))
    assert st == 3 and not h.value and b"orbm_create" in lib.orb_last_error() or st != 0
    st = lib.orbx_create(1000, C.c_float(1.2), 8, 20, 7, 752, 480, 1, 0, C.byref(h))
    assert st != 0 and not h.value
    with pytest.raises(orbb200.OrbError):
        orbb200.Extractor()
    with pytest.raises(orbb200.OrbError):
        orbb200.Matcher()


def test_argument_validation_needs_no_device(lib):
    assert lib.orbx_create(1000, C.c_float(1.2), 8, 20, 7, 752, 480, 1, 0, None) == 1          # null out
    h = C.c_void_p()
    assert lib.orbx_create(0, C.c_float(1.2), 8, 20, 7, 752, 480, 1, 0, C.byref(h)) == 1       # nfeatures < 1
    assert lib.orbx_create(1000, C.c_float(1.0), 8, 20, 7, 752, 480, 1, 0, C.byref(h)) == 1    # scaleFactor <= 1
    assert lib.orbx_create(1000, C.c_float(1.2), 17, 20, 7, 752, 480, 1, 0, C.byref(h)) == 1   # too many levels
    assert lib.orbx_destroy(None) == 0 and lib.orbm_destroy(None) == 0
    assert lib.orbx_synchronize(None) == 1 and b"null" in lib.orb_last_error()
