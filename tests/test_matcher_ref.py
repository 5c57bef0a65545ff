"""Pins the matcher oracle against the REFERENCE'S OWN ORBmatcher.cc: oracle/_ref/libmatch_ref.so is that file compiled
where it lies under /root/reference (oracle/ref_shim/matcher stands in for Frame / KeyFrame / MapPoint / cv::Mat), behind
the same flat-array calls as the oracle.  Identical seeded inputs, identical outputs demanded (match indices, counts,
updated vbPrevMatched).  The 64x48 grid under the windowed searches is the reference's own as well: the text of
Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (Frame.cc:574-589, 671-736) and KeyFrame::GetFeaturesInArea
(KeyFrame.cc:1138-1177) is streamed into the same library (oracle/ref_shim/grid), and test_grid_* compare the oracle's
restatement with it directly."""
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


def _grid_keypoint_sets(scenes):
    rng = np.random.default_rng(77)
    ka, da, kb, db, bounds = scenes["euroc"]
    yield "extracted", kb, bounds
    # undistorted keypoints can leave the image (Frame.cc:730-733 drops them), sit on cell borders, on the bounds, or at
    # half-cell positions where round() decides the cell
    k = np.zeros(4000, ka.dtype)
    k["x"] = rng.uniform(-40, 800, len(k)).astype(np.float32)
    k["y"] = rng.uniform(-40, 520, len(k)).astype(np.float32)
    k["octave"] = rng.integers(0, 8, len(k))
    cw, chh = 752.0 / 64, 480.0 / 48
    k["x"][:600] = (np.arange(600) % 66 - 1) * np.float32(cw) * np.float32(0.5)       # multiples of half a cell
    k["y"][:600] = (np.arange(600) // 66) * np.float32(chh) * np.float32(0.5)
    k["x"][600:620] = [0, 752, 751.99994, -0.0, 1e-30, 5.875, 11.75, 17.625, 746.125, 740.25] * 2
    k["y"][600:620] = [0, 480, 479.99997, 0, 0, 5.0, 10.0, 15.0, 475.0, 470.0] * 2
    yield "adversarial", k, bounds
    # undistorted bounds that do not start at 0 (Frame::ComputeImageBounds with distortion)
    k2 = k.copy()
    yield "shifted_bounds", k2, (-13.6, -9.2, 771.3, 492.8)
    yield "empty", k[:0], bounds


def test_grid_csr_equals_reference(oracle, ref, scenes):
    for name, keys, bounds in _grid_keypoint_sets(scenes):
        rs, ri = ref.grid_csr(keys, bounds)
        desc = np.zeros((len(keys), 32), np.uint8)
        gs, gi = oracle.frame(keys, desc, bounds).grid()
        assert np.array_equal(gs, rs), name
        assert np.array_equal(gi, ri), name
        if name != "empty":
            assert rs[-1] > 0


def test_features_in_area_equals_reference(oracle, ref, scenes):
    rng = np.random.default_rng(78)
    for name, keys, bounds in _grid_keypoint_sets(scenes):
        desc = np.zeros((len(keys), 32), np.uint8)
        of = oracle.frame(keys, desc, bounds)
        n_hits = 0
        queries = [(float(x), float(y), float(r)) for x, y, r in zip(rng.uniform(-60, 820, 150), rng.uniform(-60, 540, 150),
                                                                   rng.choice([0.5, 3, 15, 40, 100, 1000], 150))]
        queries += [(0.0, 0.0, 10.0), (752.0, 480.0, 10.0), (bounds[0], bounds[1], 1e-3), (376.0, 240.0, 5000.0), (-500.0, -500.0, 10.0),
                    (376.0, 240.0, 0.0)]
        for (x, y, r) in queries:
            for (mn, mx) in ((-1, -1), (0, 3), (2, -1), (4, 4), (0, 0), (7, 2)):
                want = ref.features_in_area(keys, bounds, x, y, r, mn, mx)
                got = of.area(x, y, r, mn, mx)
                assert np.array_equal(got, want), (name, x, y, r, mn, mx)
                n_hits += len(want)
            # KeyFrame overload: no level filter, no empty-cell shortcut
            assert np.array_equal(of.area(x, y, r, -1, -1), ref.features_in_area(keys, bounds, x, y, r, keyframe=True)), (name, x, y, r)
        if name != "empty":
            assert n_hits > 1000
