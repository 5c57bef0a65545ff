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


def test_frame_side_entry_points_reject_null_handles(lib):
    """the SURVEY section-8f entry points: a null handle is an error with a message, never a crash or a silent no-op"""
    buf = (C.c_float * 8)()
    ibuf = (C.c_int * 8)()
    out = C.c_void_p()
    calls = [
        lambda: lib.orbm_undistort_points(None, None, buf, 1, buf),
        lambda: lib.orbm_frame_create_device(None, buf, buf, ibuf, 4, None, C.c_float(0), C.c_float(0), C.c_float(1), C.c_float(1),
                                             None, C.byref(out)),
        lambda: lib.orbm_frame_size(None, ibuf),
        lambda: lib.orbm_frame_download(None, None, None),
        lambda: lib.orbx_compute_stereo_matches(None, 0, None, 0, buf, buf, 1, buf, buf, 1, C.c_float(0.1), C.c_float(40), buf, buf, ibuf),
        lambda: lib.orbx_compute_stereo_matches_device(None, 0, None, 0, buf, buf, ibuf, 1, buf, buf, ibuf, 1, C.c_float(0.1),
                                                       C.c_float(40), buf, buf, ibuf, ibuf, None),
        lambda: lib.orbm_distinctive_descriptors(None, buf, ibuf, 1, ibuf, ibuf),
        lambda: lib.orbm_vocabulary_create(None, 2, 1, buf, ibuf, ibuf, ibuf, buf, C.byref(out)),
        lambda: lib.orbm_bow_transform(None, None, buf, 1, 4, ibuf, buf, ibuf, None, None, None, None, None, None, None),
        lambda: lib.orbm_bow_transform_device(None, None, buf, 1, 4, ibuf, buf, ibuf, None),
    ]
    for call in calls:
        assert call() == 1                       # ORB_ERR_INVALID
        assert b"null" in lib.orb_last_error()
    # orbm_image_bounds needs the device only for a distorted camera: without one it is plain arithmetic (Frame.cc:803-808)
    b = (C.c_float * 4)()
    assert lib.orbm_image_bounds(None, None, 752, 480, b) == 0 and list(b) == [0.0, 0.0, 752.0, 480.0]
    assert lib.orbm_image_bounds(None, None, 0, 480, b) == 1
    assert lib.orbm_vocabulary_destroy(None) == 0


def test_frame_adapter_opencv_signatures_compile_and_link(lib, tmp_path):
    """adapter/Frame.h's overloads on the reference's member types, against the OpenCV stand-in; no device needed"""
    import subprocess
    pkg = os.path.join(ROOT, "vi-orb-slam-icra2018_b200")
    exe = str(tmp_path / "adapter_frame_opencv_sig")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "oracle", "ref_shim", "include"),
                           "-I", os.path.join(pkg, "adapter"), "-o", exe, os.path.join(ROOT, "tests", "adapter_frame_opencv_sig.cpp"),
                           os.path.join(ROOT, "oracle", "cvprims.cpp"), "-L", pkg, "-lorbb200", "-Wl,-rpath," + pkg])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr)
