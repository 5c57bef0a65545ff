"""The C++ adapter (reference class and method names over the C ABI) compiled with g++ and run on the GPU."""
import os
import subprocess

import numpy as np
import pytest

import orbb200
from orbb200.synth import synth_frame

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adapter_matches_c_abi(tmp_path, matcher):
    pkg = os.path.join(ROOT, "vi-orb-slam-icra2018_b200")
    exe = str(tmp_path / "adapter_smoke")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(pkg, "adapter"), "-o", exe,
                           os.path.join(ROOT, "tests", "adapter_smoke.cpp"), "-L", pkg, "-lorbb200", "-Wl,-rpath," + pkg])
    img = synth_frame(0)
    raw = tmp_path / "img.raw"
    img.tofile(raw)
    # a vocabulary file in the reference's text format (adapter/ORBVocabulary.h reads it)
    from bow_cases import make_vocab
    from test_vocab_io import write_text
    voc_file = str(tmp_path / "voc.txt")
    file_voc = make_vocab(seed=31, k=10, L=3)
    write_text(file_voc, voc_file, weight_fmt="%.17g")
    out = subprocess.run([exe, str(raw), "752", "480", voc_file], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    n_kp, checksum, n_match, levels, sf7 = out.stdout.split("\n")[0].split()
    ex = orbb200.Extractor(1000)
    kps, desc = ex(img)
    s = 0
    for b in desc.reshape(-1).tolist():
        s = (s * 131 + b) & 0xFFFFFFFFFFFFFFFF
    assert int(n_kp) == len(kps) and int(checksum) == s and int(levels) == 8
    f = matcher.frame(kps, desc, (0, 0, 752, 480))
    n, m12, _ = matcher.search_for_initialization(f, f, np.stack([kps["x"], kps["y"]], 1), 100, 0.9, True)
    assert int(n_match) == n
    assert int(out.stdout.split("\n")[1]) == int(matcher.distance(desc[0], desc[1])[0])
    # adapter/Frame.h: ComputeStereoMatches, ComputeImageBounds, UndistortKeyPoints, ComputeDistinctiveDescriptors
    kept, min_x, max_y, sum_x, best0, best1 = out.stdout.split("\n")[2].split()
    right = orbb200.Extractor(1000)
    kr, dr = right(img)
    ur, depth, n_st = ex.stereo_matches(right, kps, desc, kr, dr, np.float32(0.11), np.float32(47.9))
    assert int(kept) == n_st
    cam = orbb200.camera(458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)
    b = matcher.image_bounds(cam, 752, 480)
    assert np.float32(min_x) == b[0] and np.float32(max_y) == b[3]
    un = matcher.undistort_points(cam, np.stack([kps["x"], kps["y"]], 1))
    assert float(sum_x) == float(np.cumsum(un[:, 0].astype(np.float64))[-1])
    best, _ = matcher.distinctive_descriptors(desc[:12], [0, 5, 12])
    assert [int(best0), int(best1)] == best.tolist()
    # ComputeBoW through the adapter on the same two-level tree, rebuilt here
    n_words, n_fv, sum_v, first_idx = out.stdout.split("\n")[3].split()
    nn = 21
    ndesc = np.zeros((nn, 32), np.uint8)
    ndesc[1:] = desc[np.arange(1, nn) * 7]
    cstart = np.zeros(nn + 1, np.int32)
    cstart[1:6] = [4, 8, 12, 16, 20]
    cstart[6:] = 20
    voc = dict(n_nodes=nn, L=2, desc=ndesc, child_start=cstart, children=np.arange(1, nn, dtype=np.int32),
               word_id=np.maximum(np.arange(nn) - 5, 0).astype(np.int32),
               weight=np.where(np.arange(nn) >= 5, 1.0 + 0.25 * (np.arange(nn) - 5), 0.0))
    v = matcher.vocabulary(voc)
    bow = matcher.bow_transform(v, desc, 1)
    assert int(n_words) == len(bow["bow_word"]) and int(n_fv) == len(bow["fv_node"]) and int(first_idx) == bow["fv_idx"][0]
    assert float(sum_v) == float(np.cumsum(bow["bow_value"])[-1])
    matcher.vocabulary_destroy(v)
    # ORBVocabulary: file -> flat arrays -> device -> transform; expected from DBoW2's own reader + the Python harness
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_vocio
    nodes, words, n_bow, n_fv2, sum_v2, first2 = out.stdout.split("\n")[4].split()
    assert int(nodes) == file_voc["n_nodes"] + 1 and int(words) == file_voc["n_words"] + 1   # + the empty-last-line node
    if ref_vocio.available():
        loaded = ref_vocio.load(voc_file, False)
        loaded["L"] = file_voc["L"]
        v2 = matcher.vocabulary(loaded)
        bow2 = matcher.bow_transform(v2, desc, 2)
        assert int(n_bow) == len(bow2["bow_word"]) and int(n_fv2) == len(bow2["fv_node"]) and int(first2) == bow2["fv_idx"][0]
        assert float(sum_v2) == float(np.cumsum(bow2["bow_value"])[-1])
        matcher.vocabulary_destroy(v2)
    right.close()
    ex.close()


def test_reference_signatures_on_object_graphs():
    """The adapter's reference-signature overloads (adapter/ORBmatcher.h + ORBmatcher_orbslam.inl: what a maintainer
    compiles instead of ORBmatcher.cc) run on Frame / KeyFrame / MapPoint object graphs -- the stand-ins of
    oracle/ref_shim/matcher, prebuilt into oracle/_ref/libmatch_adapter.so where /root/reference is mounted -- and must
    reproduce what the reference's OWN ORBmatcher.cc produced on the same graphs (tests/golden/matcher_ref_vectors.npz,
    and the live libmatch_ref.so when it is there): every search, match for match."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_matcher
    from matcher_cases import run_case, same
    if not ref_matcher.adapter_available():
        pytest.skip("oracle/_ref/libmatch_adapter.so is built only where /root/reference is mounted (DBoW2's FeatureVector)")
    gold = os.path.join(ROOT, "tests", "golden")
    g = np.load(os.path.join(gold, "matcher_vectors.npz"))
    r = np.load(os.path.join(gold, "matcher_ref_vectors.npz"))
    ka, da, kb, db, bounds = g["ka"], g["da"], g["kb"], g["db"], g["bounds"]
    ad = ref_matcher.RefMatcher(adapter=True)
    a1, a2 = ad.frame(ka, da, bounds), ad.frame(kb, db, bounds)
    live = ref_matcher.RefMatcher() if ref_matcher.available() else None
    for c in [str(x) for x in r["cases"]]:
        got = run_case(c, a1, a2, ka, da, kb, db)
        same(got, [r["%s__%d" % (c, j)] for j in range(len(got))])
        if live:
            same(got, run_case(c, live.frame(ka, da, bounds), live.frame(kb, db, bounds), ka, da, kb, db))


def test_adapter_frame_cache_and_device_hand_over(oracle):
    """adapter: ORBextractor::operator() -> ORBmatcher::RegisterFrame (device frame built from what the extractor left on the
    device, cached under Frame::mnId) -> SearchForInitialization from two ORBmatcher objects, the second on a copy of the
    Frame.  No search uploads keypoints (4 cache hits, 0 misses), the handle pool hands the same matcher back, and the matches
    equal the oracle's on the host keypoints the extractor returned."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_matcher
    from orbb200.synth import shifted_pair
    if not ref_matcher.adapter_available():
        pytest.skip("oracle/_ref/libmatch_adapter.so is built only where /root/reference is mounted")
    a, b = shifted_pair(4)
    ka, da, kb, db, m12, n1, n2, hits, misses = ref_matcher.adapter_chain_init(a, b)
    assert (hits, misses) == (4, 0)
    ex = orbb200.Extractor(1000)
    k, d = ex(a)
    assert k.tobytes() == ka.tobytes() and np.array_equal(d, da)
    ex.close()
    bounds = (0.0, 0.0, 752.0, 480.0)
    o1, o2 = oracle.frame(ka, da, bounds), oracle.frame(kb, db, bounds)
    prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    rn, rm12, rp = o1.search_init(o2, prev, 100, 0.9, True)
    assert n1 == rn and np.array_equal(m12, rm12) and rn > 50
    rn2, _, _ = o1.search_init(o2, rp, 100, 0.9, True)           # the second call sees the updated vbPrevMatched
    assert n2 == rn2


def test_extractor_opencv_signature(tmp_path):
    """adapter/ORBextractor.h with -DORBB200_WITH_OPENCV, compiled against the OpenCV stand-in the reference's own
    ORBextractor.cc is built against (oracle/ref_shim/include): operator()(InputArray, InputArray, vector<KeyPoint>&,
    OutputArray) and the public mvImagePyramid member behave like the reference's (Frame.cc:591-597, 817)."""
    pkg = os.path.join(ROOT, "vi-orb-slam-icra2018_b200")
    exe = str(tmp_path / "adapter_opencv_sig")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "oracle", "ref_shim", "include"),
                           "-I", os.path.join(pkg, "adapter"), "-o", exe, os.path.join(ROOT, "tests", "adapter_opencv_sig.cpp"),
                           os.path.join(ROOT, "oracle", "cvprims.cpp"), "-L", pkg, "-lorbb200", "-Wl,-rpath," + pkg])
    img = synth_frame(0)
    raw = tmp_path / "img.raw"
    img.tofile(raw)
    out = subprocess.run([exe, str(raw), "752", "480"], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stderr)
    n_kp, checksum, levels = out.stdout.split()
    ex = orbb200.Extractor(1000)
    kps, desc = ex(img)
    s = 0
    for b in desc.reshape(-1).tolist():
        s = (s * 131 + b) & 0xFFFFFFFFFFFFFFFF
    assert int(n_kp) == len(kps) and int(checksum) == s and int(levels) == 8
    ex.close()
