"""Pins the matcher oracle against the REFERENCE'S OWN ORBmatcher.cc: oracle/_ref/libmatch_ref.so is that file compiled
where it lies under /root/reference (oracle/ref_shim/matcher stands in for Frame / KeyFrame / MapPoint / cv::Mat), behind
the same flat-array calls as the oracle.  Identical seeded inputs, identical outputs demanded (match indices, counts,
updated vbPrevMatched).  Not pinned by this: the 64x48 grid lookup (Frame.cc / KeyFrame.cc cannot be compiled here; both
sides share the oracle's restatement, which has its own known-answer tests in test_oracle.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_matcher  # noqa: E402
from matcher_cases import CASES, run_case, same  # noqa: E402
from orbb200.synth import shifted_pair  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_matcher.available() and not os.path.isdir("/root/reference"),
                                reason="oracle/_ref/libmatch_ref.so is built only where /root/reference is mounted")


@pytest.fixture(scope="module")
def ref():
    if not ref_matcher.available():
        ref_matcher.build()
    return ref_matcher.RefMatcher()


@pytest.fixture(scope="module")
def scenes(oracle):
    out = {}
    for name, (w, h, nf) in (("euroc", (752, 480, 2000)), ("small", (400, 300, 800))):
        a, b = shifted_pair(3, w, h)
        oe = oracle.extractor(nf)
        ka, da = oe.extract(a)
        kb, db = oe.extract(b)
        out[name] = (ka, da, kb, db, (0.0, 0.0, float(w), float(h)))
    return out


def test_descriptor_distance(oracle, ref):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    b[:50] = a[:50]
    b[50:60] = ~a[50:60]
    for i in range(200):
        assert ref.distance(a[i], b[i]) == oracle.lib.orbo_distance(a[i].ctypes.data, b[i].ctypes.data)


@pytest.mark.parametrize("scene", ["euroc", "small"])
@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_reference(oracle, ref, scenes, scene, case):
    ka, da, kb, db, bounds = scenes[scene]
    got = run_case(case, oracle.frame(ka, da, bounds), oracle.frame(kb, db, bounds), ka, da, kb, db)
    want = run_case(case, ref.frame(ka, da, bounds), ref.frame(kb, db, bounds), ka, da, kb, db)
    same(got, want)
    assert int(want[0]) > 10          # the case exercises the accept path
