"""Pins the oracle's ComputeStereoMatches against the REFERENCE'S OWN text of Frame::ComputeStereoMatches
(Frame.cc:810-984): oracle/_ref/libstereo_ref.so is those lines compiled where they lie (oracle/ref_shim/stereo holds the
stand-ins for Frame / ORBextractor / the slice of cv::Mat they use).  Identical inputs, identical mvuRight and mvDepth
demanded bit for bit; and the committed golden (outputs of that reference build) is checked against the oracle too."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_stereo  # noqa: E402
from stereo_cases import STEREO_CASES, oracle_inputs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "stereo_ref_vectors.npz")


@pytest.mark.skipif(not ref_stereo.available() and not os.path.isdir("/root/reference"),
                    reason="oracle/_ref/libstereo_ref.so is built only where /root/reference is mounted")
@pytest.mark.parametrize("name", sorted(STEREO_CASES))
def test_oracle_equals_reference(oracle, name):
    if not ref_stereo.available():
        ref_stereo.build()
    args = oracle_inputs(oracle, name)
    ur, depth, sad, kept = oracle.stereo(*args)
    rur, rdepth, rkept = ref_stereo.stereo(*args)
    assert kept == rkept and kept > 50
    assert np.array_equal(ur.view(np.uint32), rur.view(np.uint32))
    assert np.array_equal(depth.view(np.uint32), rdepth.view(np.uint32))
    assert np.array_equal(sad >= 0, rur >= 0)


@pytest.mark.skipif(not ref_stereo.available() and not os.path.isdir("/root/reference"),
                    reason="oracle/_ref/libstereo_ref.so is built only where /root/reference is mounted")
@pytest.mark.parametrize("name", ["euroc_s1", "small_wide"])
def test_oracle_equals_reference_on_perturbed_keypoints(oracle, name):
    """the same pin on inputs the extractor would not produce: sub-pixel jitter on both keypoint sets (rounding of the
    scaled coordinates, row bands and disparity windows move), octaves shifted by +-1, shuffled order (ties go to the lowest
    index), other baselines"""
    if not ref_stereo.available():
        ref_stereo.build()
    kl, dl, kr, dr, LL, RR, sf, isf, mb, mbf = oracle_inputs(oracle, name)
    for seed in range(8):
        rng = np.random.default_rng(seed)
        a, b = kl.copy(), kr.copy()
        for k in (a, b):
            k["x"] += rng.uniform(-1.5, 1.5, len(k)).astype(np.float32)
            k["y"] += rng.uniform(-1.5, 1.5, len(k)).astype(np.float32)
        b["octave"] = np.clip(b["octave"] + rng.integers(-1, 2, len(b)), 0, 7)
        pa, pb = rng.permutation(len(a)), rng.permutation(len(b))
        a, da, b, db = a[pa], dl[pa], b[pb], dr[pb]
        mbf2 = float(mbf * rng.uniform(0.3, 2.0))
        ur, depth, sad, kept = oracle.stereo(a, da, b, db, LL, RR, sf, isf, mb, mbf2)
        rur, rdepth, rkept = ref_stereo.stereo(a, da, b, db, LL, RR, sf, isf, mb, mbf2)
        assert kept == rkept and kept > 20
        assert np.array_equal(ur.view(np.uint32), rur.view(np.uint32))
        assert np.array_equal(depth.view(np.uint32), rdepth.view(np.uint32))


@pytest.mark.parametrize("name", sorted(STEREO_CASES))
def test_oracle_equals_golden(oracle, name):
    g = np.load(GOLDEN)
    ur, depth, sad, kept = oracle.stereo(*oracle_inputs(oracle, name))
    assert kept == int(g[name + "/kept"])
    assert np.array_equal(ur.view(np.uint32), g[name + "/u_right"].view(np.uint32))
    assert np.array_equal(depth.view(np.uint32), g[name + "/depth"].view(np.uint32))


def test_known_answers(oracle):
    """hand-made keypoints on a real pyramid: a right keypoint outside the row band, outside the octave window or right
    of the left keypoint is never matched; with no right keypoints nothing is matched"""
    args = list(oracle_inputs(oracle, "euroc_s1"))
    kl, dl, kr, dr = args[:4]
    ur, depth, sad, kept = oracle.stereo(kl, dl, kr[:0], dr[:0], *args[4:])
    assert kept == 0 and (ur == -1).all() and (depth == -1).all()
    i = int(np.flatnonzero((kl["octave"] == 0) & (kl["x"] > 100))[0])
    one_l, one_d = kl[i:i + 1], dl[i:i + 1]
    for dy, doct, dx, expect in ((0.0, 0, -6.0, True), (3.5, 0, -6.0, False), (0.0, 2, -6.0, False), (0.0, 0, 4.0, False)):
        r = one_l.copy()
        r["y"] += dy
        r["octave"] += doct
        r["x"] += dx
        # a second, unrelated right keypoint far away keeps the median step well defined
        ur, depth, sad, kept = oracle.stereo(one_l, one_d, r, one_d, *args[4:])
        assert (sad[0] >= 0 or ur[0] >= 0) == (expect and ur[0] >= 0)
        if not expect:
            assert ur[0] == -1 and depth[0] == -1


@pytest.mark.parametrize("name", sorted(STEREO_CASES))
def test_synthetic_stereo_images_are_reproducible(name):
    """the goldens hold outputs only; their inputs are regenerated from seeds, so the generator must not drift"""
    import hashlib
    from stereo_cases import images
    g = np.load(GOLDEN)
    left, right = images(name)
    assert hashlib.sha256(np.ascontiguousarray(left).tobytes()).hexdigest() == str(g[name + "/left_sha256"])
    assert hashlib.sha256(np.ascontiguousarray(right).tobytes()).hexdigest() == str(g[name + "/right_sha256"])
