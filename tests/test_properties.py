"""Property tests (hypothesis) of the oracle's matcher-side functions against independent numpy reasoning -- SURVEY.md
section 4: "property tests for the matcher".  CPU only; the GPU parity tests then compare the CUDA path with this oracle."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

SET = settings(max_examples=40, deadline=None)


def _desc(rng, n, cluster=True):
    d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    if cluster and n > 2:           # near-duplicates, so that thresholds, ties and the ratio test are exercised
        src = rng.integers(0, n, n // 2)
        dst = rng.integers(0, n, n // 2)
        for s, t in zip(src, dst):
            bits = np.unpackbits(d[s])
            bits[rng.permutation(256)[:rng.integers(0, 40)]] ^= 1
            d[t] = np.packbits(bits)
    return d


def _dist(a, b):
    return np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(2).astype(np.int64)


@SET
@given(st.integers(0, 2 ** 31), st.integers(1, 60), st.integers(1, 60), st.sampled_from([0.6, 0.75, 0.9]))
def test_bruteforce_best_second_and_acceptance(oracle, seed, nq, nt, ratio):
    """best / index = first minimum, second = second-smallest with multiplicity, acceptance rule ORBmatcher.cc:459-461"""
    rng = np.random.default_rng(seed)
    q, t = _desc(rng, nq), _desc(rng, nt)
    qa = (rng.random(nq) * 360).astype(np.float32)
    ta = (rng.random(nt) * 360).astype(np.float32)
    n, best, second, idx, m12 = oracle.bruteforce(q, qa, t, ta, ratio, False)
    D = _dist(q, t)
    assert np.array_equal(idx, D.argmin(1))                       # argmin returns the FIRST minimum: strict '<'
    assert np.array_equal(best, D.min(1))
    if nt > 1:
        assert np.array_equal(second, np.sort(D, 1)[:, 1])
    for i in range(nq):
        sec = float(np.sort(D[i])[1]) if nt > 1 else float(second[i])
        ok = best[i] <= 50 and np.float32(best[i]) < np.float32(sec) * np.float32(ratio)
        assert (m12[i] >= 0) == bool(ok)
        if ok:
            assert m12[i] == idx[i]
    assert n == int((m12 >= 0).sum())


@SET
@given(st.integers(0, 2 ** 31), st.integers(2, 80))
def test_rotation_histogram_only_removes(oracle, seed, n):
    """the orientation check can only withdraw matches, never add or move one (ORBmatcher.cc:489-512)"""
    rng = np.random.default_rng(seed)
    q = _desc(rng, n)
    t = q.copy()
    rng.shuffle(t)
    qa = (rng.random(n) * 360).astype(np.float32)
    ta = (rng.random(n) * 360).astype(np.float32)
    n0, _, _, _, m_off = oracle.bruteforce(q, qa, t, ta, 0.9, False)
    n1, _, _, _, m_on = oracle.bruteforce(q, qa, t, ta, 0.9, True)
    assert n1 <= n0
    kept = m_on >= 0
    assert np.array_equal(m_on[kept], m_off[kept])
    # all kept matches fall into at most three of the 30 bins
    bins = {oracle.lib.orbo_rotation_bin(float(qa[i]), float(ta[m_on[i]])) for i in np.flatnonzero(kept)}
    assert len(bins) <= 3


@SET
@given(st.integers(0, 2 ** 31), st.integers(1, 40))
def test_distinctive_descriptor_has_the_least_median(oracle, seed, n):
    """MapPoint.cc:305-318: the chosen row's median (rank floor((n-1)/2), own 0 included) is minimal, first such row"""
    rng = np.random.default_rng(seed)
    d = _desc(rng, n)
    best, med = oracle.distinctive(d, [0, n])
    D = _dist(d, d)
    medians = np.sort(D, 1)[:, (n - 1) // 2]
    assert med[0] == medians.min() and best[0] == int(np.argmin(medians))


@SET
@given(st.integers(0, 2 ** 31), st.integers(0, 300))
def test_grid_partitions_the_keypoints(oracle, seed, n):
    """AssignFeaturesToGrid (Frame.cc:574-589): every keypoint whose rounded cell is inside 64x48 sits in exactly that
    bucket, buckets hold ascending indices, nothing else is stored"""
    from oracle_py import KP_DTYPE
    rng = np.random.default_rng(seed)
    k = np.zeros(n, KP_DTYPE)
    k["x"] = rng.uniform(-5, 757, n).astype(np.float32)
    k["y"] = rng.uniform(-5, 485, n).astype(np.float32)
    f = oracle.frame(k, np.zeros((n, 32), np.uint8), (0.0, 0.0, 752.0, 480.0))
    start, idx = f.grid()
    inv_w, inv_h = np.float32(64) / np.float32(752), np.float32(48) / np.float32(480)
    px = np.round(k["x"] * inv_w)     # round half away from zero == numpy's half-to-even except at exact .5, excluded below
    py = np.round(k["y"] * inv_h)
    frac_x, frac_y = np.abs(k["x"] * inv_w % 1 - 0.5), np.abs(k["y"] * inv_h % 1 - 0.5)
    sure = (frac_x > 1e-3) & (frac_y > 1e-3)
    inside = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    assert start[0] == 0 and np.all(np.diff(start) >= 0)
    stored = np.zeros(n, bool)
    for c in range(64 * 48):
        members = idx[start[c]:start[c + 1]]
        assert np.all(np.diff(members) > 0)
        stored[members] = True
        for i in members:
            if sure[i]:
                assert int(px[i]) * 48 + int(py[i]) == c
    assert np.array_equal(stored[sure], inside[sure])
    assert len(idx) == int(stored.sum())


@SET
@given(st.integers(0, 2 ** 31), st.sampled_from([(-0.2834, 0.0739, 0.00019, 1.76e-05), (0.12, 0.05, -0.001, 0.0007), (-0.05, 0.0, 0.0, 0.0)]))
def test_undistort_inverts_the_distortion_model(oracle, seed, dist):
    """cv::undistortPoints solves x_d = distort(x_u) by FIVE fixed-point iterations: re-distorting its output with the
    Brown model gives back the input to within half a pixel even in the corners of the EuRoC camera (0.28 px there: the
    iteration has not converged, and must not be "improved"), and to about a tenth elsewhere"""
    rng = np.random.default_rng(seed)
    K4 = (458.654, 457.296, 367.215, 248.375)
    pts = (rng.random((200, 2)) * [752, 480]).astype(np.float32)
    un = oracle.undistort(pts, K4, dist).astype(np.float64)
    k1, k2, p1, p2 = dist
    x = (un[:, 0] - K4[2]) / K4[0]
    y = (un[:, 1] - K4[3]) / K4[1]
    r2 = x * x + y * y
    radial = 1 + k1 * r2 + k2 * r2 * r2
    xd = x * radial + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * radial + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    back = np.stack([xd * K4[0] + K4[2], yd * K4[1] + K4[3]], 1)
    assert np.abs(back - pts).max() < (0.5 if dist[0] < -0.2 else 0.15)


@SET
@given(st.integers(0, 2 ** 31), st.integers(2, 6), st.integers(1, 3))
def test_bow_descent_is_greedy_and_l1_normalised(oracle, seed, k, L):
    """every step of the descent goes to a child at minimal distance (the first of them); values sum to 1"""
    from bow_cases import features_for, make_vocab
    voc = make_vocab(seed % 1000, k=k, L=L, stop_fraction=0.0)
    f = features_for(voc, seed, 60)
    r = oracle.bow_transform(voc, f, 1)
    leaf_of_word = {int(voc["word_id"][i]): i for i in range(1, voc["n_nodes"])
                    if voc["child_start"][i + 1] == voc["child_start"][i]}
    for i in range(len(f)):
        node = 0
        while voc["child_start"][node + 1] > voc["child_start"][node]:
            kids = voc["children"][voc["child_start"][node]:voc["child_start"][node + 1]]
            d = _dist(f[i:i + 1], voc["desc"][kids])[0]
            node = int(kids[int(np.argmin(d))])
        assert leaf_of_word[int(r["word"][i])] == node
    assert abs(r["bow_value"].sum() - 1.0) < 1e-12
    assert sorted(r["fv_idx"].tolist()) == list(range(len(f)))
