"""GPU parity of the Frame grid and the windowed / vocabulary-gated searches against the oracle restatement of
Frame.cc:574-736 and ORBmatcher.cc:45-129, 405-520, 657-823, 1341-1498.  Everything here is integer or bit-exact float."""
import numpy as np
import pytest

import orbb200
from orbb200.synth import shifted_pair

pytestmark = pytest.mark.gpu

SF = np.array([1.2 ** i for i in range(8)], np.float32)


@pytest.fixture(scope="module")
def pair(oracle):
    """Two views of one synthetic scene (second shifted by (+7,+3) px), 2000 features each (config 1 / 2 inputs)."""
    out = {}
    for name, (w, h) in (("euroc", (752, 480)), ("kitti", (1241, 376))):
        a, b = shifted_pair(3, w, h)
        oe = oracle.extractor(2000)
        ka, da = oe.extract(a)
        kb, db = oe.extract(b)
        out[name] = (ka, da, kb, db, (0.0, 0.0, float(w), float(h)))
    return out


def _frames(matcher, oracle, k, d, bounds):
    return matcher.frame(k, d, bounds), oracle.frame(k, d, bounds)


@pytest.mark.parametrize("name", ["euroc", "kitti"])
def test_grid_and_area(matcher, oracle, pair, name):
    ka, da, kb, db, bounds = pair[name]
    # keys outside the bounds must be dropped (PosInGrid false), keys on a .5 boundary go to round-half-away
    k = ka.copy()
    k["x"][:5] = [-3.0, bounds[2] + 5, 0.0, bounds[2], bounds[2] * (10.5 / 64)]
    k["y"][5:9] = [-1.0, bounds[3] + 9, bounds[3], bounds[3] * (7.5 / 48)]
    g, o = _frames(matcher, oracle, k, da, bounds)
    gs, gi = g.grid()
    os_, oi = o.grid()
    assert np.array_equal(gs, os_) and np.array_equal(gi, oi)
    assert gs[-1] < len(k)
    rng = np.random.default_rng(1)
    xyr = np.stack([rng.uniform(-50, bounds[2] + 50, 300), rng.uniform(-50, bounds[3] + 50, 300),
                    rng.choice([2.5, 15.0, 30.0, 100.0, 400.0], 300)], 1).astype(np.float32)
    for (mn, mx) in ((-1, -1), (0, 0), (2, 4), (3, -1), (0, 7)):
        got = g.area(xyr, mn, mx)
        for i in range(len(xyr)):
            ref = o.area(float(xyr[i, 0]), float(xyr[i, 1]), float(xyr[i, 2]), mn, mx)
            assert np.array_equal(got[i], ref), (i, mn, mx)


@pytest.mark.parametrize("name,ratio,ori,window", [("euroc", 0.9, True, 100), ("euroc", 0.9, False, 30),
                                                    ("kitti", 0.6, True, 100)])
def test_search_for_initialization(matcher, oracle, pair, name, ratio, ori, window):
    ka, da, kb, db, bounds = pair[name]
    g1, o1 = _frames(matcher, oracle, ka, da, bounds)
    g2, o2 = _frames(matcher, oracle, kb, db, bounds)
    prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)   # Tracking.cc:1637-1639
    n, m12, p = matcher.search_for_initialization(g1, g2, prev, window, ratio, ori)
    rn, rm12, rp = o1.search_init(o2, prev, window, ratio, ori)
    assert n == rn and np.array_equal(m12, rm12) and np.array_equal(p, rp)
    assert n > 50
    # second call with the updated vbPrevMatched, roles of the frames kept (as the tracker does frame after frame)
    n2, m12b, p2 = matcher.search_for_initialization(g1, g2, p, window, ratio, ori)
    rn2, rm12b, rp2 = o1.search_init(o2, rp, window, ratio, ori)
    assert n2 == rn2 and np.array_equal(m12b, rm12b) and np.array_equal(p2, rp2)


def _proj_queries(rng, ka, n_levels=8):
    q = np.zeros(len(ka), orbb200.PROJ_QUERY_DTYPE)
    q["u"] = ka["x"] - 7 + rng.normal(0, 1.5, len(ka))
    q["v"] = ka["y"] - 3 + rng.normal(0, 1.5, len(ka))
    q["invz"] = rng.uniform(0.02, 0.5, len(ka))
    q["octave"] = ka["octave"]
    q["valid"] = rng.random(len(ka)) < 0.8
    q["obs_positive"] = rng.random(len(ka)) < 0.9
    q["angle"] = ka["angle"]
    q["u"][:3] = [-5, 1e4, 10]
    q["v"][:3] = [10, 10, -8]
    return q


@pytest.mark.parametrize("name,th,mode,stereo,ori", [("kitti", 15.0, 0, False, True), ("kitti", 30.0, 0, False, True),
                                                      ("euroc", 7.0, 1, True, True), ("euroc", 15.0, 2, True, False)])
def test_search_by_projection(matcher, oracle, pair, name, th, mode, stereo, ori):
    ka, da, kb, db, bounds = pair[name]
    rng = np.random.default_rng(int(th) + mode)
    g, o = _frames(matcher, oracle, kb, db, bounds)
    q = _proj_queries(rng, ka)
    occ = (rng.random(len(kb)) < 0.1).astype(np.uint8)
    ur = None
    mbf = 0.0
    if stereo:
        mbf = 40.0
        ur = np.where(rng.random(len(kb)) < 0.6, kb["x"] - mbf * rng.uniform(0.02, 0.5, len(kb)), -1).astype(np.float32)
    n, match = matcher.search_by_projection(g, SF, q, da, th, mode, occ, ur, mbf, ori)
    oq = q.view(np.dtype([(a, b) for a, b in zip(("u", "v", "invz", "octave", "valid", "obsPositive", "angle"),
                                                   ("<f4", "<f4", "<f4", "<i4", "<i4", "<i4", "<f4"))]))
    rn, rmatch = o.search_projection(SF, oq, da, th, mode, occ, ur, mbf, ori)
    assert n == rn and np.array_equal(match, rmatch)
    if not stereo:
        assert n > 100


def test_project_points_bits(matcher, oracle):
    """orbm_project_points == match_oracle.cpp::project_points (ORBmatcher.cc:1376-1388; the oracle's arithmetic is pinned
    against cv2.gemm in test_cvprims.py), bit for bit, incl. points at and behind the camera plane."""
    from matcher_cases import camera_pose
    rng = np.random.default_rng(5)
    R, t, K4 = camera_pose(rng)
    xyz = (rng.normal(size=(50000, 3)) * [4, 3, 6] + [0, 0, 5]).astype(np.float32)
    xyz[:4] = [[0, 0, 0], [1, 2, -3], [1e6, -1e6, 1e-3], [0.5, 0.5, 1e-30]]
    u, v, iz = matcher.project_points(R, t, K4, xyz)
    ru, rv, riz, _ = oracle.project(R, t, K4, [-np.inf, -np.inf, np.inf, np.inf], xyz)
    for a, b in ((u, ru), (v, rv), (iz, riz)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert (iz < 0).sum() > 1000


@pytest.mark.parametrize("case", ["projw_window", "projw_forward_stereo", "projw_backward_stereo"])
def test_search_by_projection_world(matcher, oracle, pair, case):
    """orbm_search_by_projection_world (projection on the device) == oracle projection + search, through a non-identity
    pose with points behind the camera and outside the image."""
    from matcher_cases import GpuFrame, run_case, same
    for name in ("kitti", "euroc"):
        ka, da, kb, db, bounds = pair[name]
        got = run_case(case, GpuFrame(matcher, ka, da, bounds), GpuFrame(matcher, kb, db, bounds), ka, da, kb, db)
        want = run_case(case, oracle.frame(ka, da, bounds), oracle.frame(kb, db, bounds), ka, da, kb, db)
        same(got, want)
        assert got[0] > 100


def test_search_by_projection_relocalisation_threshold(matcher, oracle, pair):
    """SearchByProjection(Frame, KeyFrame, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1500-1627): predicted levels come
    from the host, any assigned keypoint is occupied, acceptance threshold ORBdist = 64 (Tracking.cc:2691)."""
    ka, da, kb, db, bounds = pair["euroc"]
    rng = np.random.default_rng(64)
    g, o = _frames(matcher, oracle, kb, db, bounds)
    q = _proj_queries(rng, ka)
    q["octave"] = np.clip(ka["octave"] + rng.integers(-1, 2, len(ka)), 0, 7)    # PredictScale on the host
    q["obs_positive"] = 1
    occ = (rng.random(len(kb)) < 0.2).astype(np.uint8)
    oq = q.view(np.dtype([(a, b) for a, b in zip(("u", "v", "invz", "octave", "valid", "obsPositive", "angle"),
                                                   ("<f4", "<f4", "<f4", "<i4", "<i4", "<i4", "<f4"))]))
    res = {}
    for dist in (100, 64, 30):
        n, match = matcher.search_by_projection(g, SF, q, da, 10.0, 0, occ, None, 0.0, True, max_distance=dist)
        rn, rmatch = o.search_projection(SF, oq, da, 10.0, 0, occ, None, 0.0, True, max_distance=dist)
        assert n == rn and np.array_equal(match, rmatch)
        res[dist] = n
    assert res[100] >= res[64] >= res[30] and res[64] > 50


@pytest.mark.parametrize("name,th,ratio", [("euroc", 1.0, 0.8), ("kitti", 3.0, 0.8), ("euroc", 5.0, 0.6)])
def test_search_by_projection_points(matcher, oracle, pair, name, th, ratio):
    ka, da, kb, db, bounds = pair[name]
    rng = np.random.default_rng(int(th * 10))
    g, o = _frames(matcher, oracle, kb, db, bounds)
    q = np.zeros(len(ka), orbb200.POINT_QUERY_DTYPE)
    q["proj_x"] = ka["x"] - 7 + rng.normal(0, 1.0, len(ka))
    q["proj_y"] = ka["y"] - 3 + rng.normal(0, 1.0, len(ka))
    q["proj_xr"] = q["proj_x"] - 5
    q["view_cos"] = rng.choice([0.9, 0.998, 0.9985, 1.0], len(ka))
    q["level"] = np.clip(ka["octave"] + rng.integers(0, 2, len(ka)), 0, 7)
    q["in_view"] = rng.random(len(ka)) < 0.85
    q["obs_positive"] = rng.random(len(ka)) < 0.9
    occ = (rng.random(len(kb)) < 0.05).astype(np.uint8)
    ur = np.where(rng.random(len(kb)) < 0.3, kb["x"] - 5 + rng.normal(0, 3, len(kb)), -1).astype(np.float32)
    n, match = matcher.search_by_projection_points(g, SF, q, da, th, ratio, occ, ur)
    oq = q.view(np.dtype([(a, b) for a, b in zip(("projX", "projY", "projXR", "viewCos", "level", "inView", "obsPositive"),
                                                   ("<f4", "<f4", "<f4", "<f4", "<i4", "<i4", "<i4"))]))
    rn, rmatch = o.search_points(SF, oq, da, th, ratio, occ, ur)
    assert n == rn and np.array_equal(match, rmatch)
    assert n > 50


def _feature_vector(keys, dx, dy, cell=48):
    """Stand-in for DBoW2::FeatureVector (node id -> keypoint indices, ascending ids): a coarse spatial hash, shifted so
    that corresponding points of the two views mostly share a node."""
    node = ((keys["y"] + dy) // cell).astype(np.int64) * 100 + ((keys["x"] + dx) // cell).astype(np.int64)
    ids = np.unique(node)
    start, idx = [0], []
    for i in ids:
        members = np.nonzero(node == i)[0]
        idx.extend(members.tolist())
        start.append(len(idx))
    return ids.astype(np.int32), np.array(start, np.int32), np.array(idx, np.int32)


@pytest.mark.parametrize("name,only_stereo,ori", [("euroc", False, False), ("kitti", False, True), ("euroc", True, True)])
def test_search_for_triangulation(matcher, oracle, pair, name, only_stereo, ori):
    ka, da, kb, db, bounds = pair[name]
    rng = np.random.default_rng(5 + only_stereo)
    g1, o1 = _frames(matcher, oracle, ka, da, bounds)
    g2, o2 = _frames(matcher, oracle, kb, db, bounds)
    fv1 = _feature_vector(ka, 0, 0)
    fv2 = _feature_vector(kb, 7, 3)
    fv2 = (fv2[0][::1].copy(), fv2[1], fv2[2])
    # pure image translation t=(-7,-3,0): F12 = [t]x, epipolar lines are parallel to t
    F12 = np.array([[0, 0, -3.0], [0, 0, 7.0], [3.0, -7.0, 0]], np.float32) * 1e-2
    has1 = (rng.random(len(ka)) < 0.3).astype(np.uint8)
    has2 = (rng.random(len(kb)) < 0.3).astype(np.uint8)
    ur1 = np.where(rng.random(len(ka)) < 0.5, 10.0, -1.0).astype(np.float32) if only_stereo else None
    ur2 = np.where(rng.random(len(kb)) < 0.5, 10.0, -1.0).astype(np.float32) if only_stereo else None
    sigma2 = (SF * SF).astype(np.float32)
    n, m12 = matcher.search_for_triangulation(g1, g2, fv1, fv2, F12, 3000.0, 200.0, SF, sigma2, has1, has2, ur1, ur2,
                                              only_stereo, ori)
    rn, rm12 = o1.search_triangulation(o2, fv1, fv2, F12, 3000.0, 200.0, SF, sigma2, has1, has2, ur1, ur2, only_stereo, ori)
    assert n == rn and np.array_equal(m12, rm12)
    assert n > 20


def test_search_edge_cases(matcher, oracle, pair):
    ka, da, kb, db, bounds = pair["euroc"]
    empty = matcher.frame(ka[:0], da[:0], bounds)
    g2 = matcher.frame(kb, db, bounds)
    n, m12, p = matcher.search_for_initialization(empty, g2, np.zeros((0, 2), np.float32))
    assert n == 0 and len(m12) == 0
    g1 = matcher.frame(ka, da, bounds)
    n, m12, p = matcher.search_for_initialization(g1, empty, np.stack([ka["x"], ka["y"]], 1))
    assert n == 0 and (m12 == -1).all()
    q = np.zeros(4, orbb200.PROJ_QUERY_DTYPE)
    q["valid"] = 1
    q["octave"] = 9                                  # outside the scale table
    with pytest.raises(orbb200.OrbError):
        matcher.search_by_projection(g2, SF, q, da[:4], 15.0)
    with pytest.raises(orbb200.OrbError):
        matcher.frame(ka, da, (0.0, 0.0, 0.0, 480.0))   # empty bounds


@pytest.mark.parametrize("name,strict,ratio,ori,use_valid2", [("euroc", False, 0.7, True, False),   # KeyFrame -> Frame
                                                              ("kitti", True, 0.75, True, True),     # KeyFrame -> KeyFrame
                                                              ("euroc", True, 0.9, False, True)])
def test_search_by_bow(matcher, oracle, pair, name, strict, ratio, ori, use_valid2):
    """SearchByBoW (ORBmatcher.cc:159-288, 522-655): brute force inside shared vocabulary nodes, one-to-one."""
    ka, da, kb, db, bounds = pair[name]
    rng = np.random.default_rng(17 + strict)
    g1, o1 = _frames(matcher, oracle, ka, da, bounds)
    g2, o2 = _frames(matcher, oracle, kb, db, bounds)
    fv1 = _feature_vector(ka, 0, 0, cell=64)
    fv2 = _feature_vector(kb, 7, 3, cell=64)
    valid1 = (rng.random(len(ka)) < 0.7).astype(np.uint8)
    valid2 = (rng.random(len(kb)) < 0.8).astype(np.uint8) if use_valid2 else None
    n, m12, m21 = matcher.search_by_bow(g1, g2, fv1, fv2, valid1, valid2, ratio, ori, strict)
    rn, rm12, rm21 = o1.search_bow(o2, fv1, fv2, valid1, valid2, ratio, ori, strict)
    assert n == rn and np.array_equal(m12, rm12) and np.array_equal(m21, rm21)
    assert n > 50
    # one-to-one and mutually consistent
    i1 = np.nonzero(m12 >= 0)[0]
    assert len(i1) == n and len(set(m12[i1].tolist())) == n and np.array_equal(m21[m12[i1]], i1)


@pytest.mark.parametrize("name,chi2,stereo", [("euroc", False, False), ("kitti", True, False), ("euroc", True, True)])
def test_search_projected_best(matcher, oracle, pair, name, chi2, stereo):
    """The independent projected search of Fuse / SearchBySim3 (ORBmatcher.cc:892-944, 1051-1075, 1191-1215, 1271-1295)."""
    ka, da, kb, db, bounds = pair[name]
    rng = np.random.default_rng(31 + chi2 + 2 * stereo)
    g, o = _frames(matcher, oracle, kb, db, bounds)
    q = np.zeros(len(ka), orbb200.BEST_QUERY_DTYPE)
    q["u"] = ka["x"] - 7 + rng.normal(0, 1.5, len(ka))
    q["v"] = ka["y"] - 3 + rng.normal(0, 1.5, len(ka))
    q["level"] = np.clip(ka["octave"] + rng.integers(0, 2, len(ka)), 0, 7)
    q["radius"] = 3.0 * SF[q["level"]]
    q["ur"] = q["u"] - 4.0
    q["valid"] = rng.random(len(ka)) < 0.9
    ur = np.where(rng.random(len(kb)) < 0.5, kb["x"] - 4 + rng.normal(0, 2, len(kb)), -1).astype(np.float32) if stereo else None
    inv_s2 = (1.0 / (SF * SF)).astype(np.float32)
    bi, bd = matcher.search_projected_best(g, q, da, chi2, ur, inv_s2 if chi2 else None)
    ri, rd = o.search_best(q, da, chi2, ur, inv_s2 if chi2 else None)
    assert np.array_equal(bi, ri) and np.array_equal(bd, rd)
    assert (bi >= 0).sum() > 200 and (bd[bi >= 0] < 256).all() and (bd[bi < 0] == 256).all()


def test_search_by_projection_keyframe_sim3_window(matcher, oracle, pair):
    """SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (ORBmatcher.cc:290-403): octave window [level-1, level],
    occupied = vpMatched[idx] != NULL, acceptance TH_LOW, no orientation check -> mode 3 of the projection search."""
    ka, da, kb, db, bounds = pair["euroc"]
    rng = np.random.default_rng(290)
    g, o = _frames(matcher, oracle, kb, db, bounds)
    q = _proj_queries(rng, ka)
    q["obs_positive"] = 1
    occ = (rng.random(len(kb)) < 0.3).astype(np.uint8)
    oq = q.view(np.dtype([(a, b) for a, b in zip(("u", "v", "invz", "octave", "valid", "obsPositive", "angle"),
                                                   ("<f4", "<f4", "<f4", "<i4", "<i4", "<i4", "<f4"))]))
    n, match = matcher.search_by_projection(g, SF, q, da, 10.0, 3, occ, None, 0.0, False, max_distance=50)
    rn, rmatch = o.search_projection(SF, oq, da, 10.0, 3, occ, None, 0.0, False, max_distance=50)
    assert n == rn and np.array_equal(match, rmatch) and n > 50


def _oq(q):
    return q.view(np.dtype([(a, b) for a, b in zip(("u", "v", "invz", "octave", "valid", "obsPositive", "angle"),
                                                   ("<f4", "<f4", "<f4", "<i4", "<i4", "<i4", "<f4"))]))


def test_search_by_projection_batch(matcher, oracle, pair):
    """orbm_search_by_projection_batch: 72 jobs of different sizes (two scenes, query subsets, with and without occupancy
    / stereo, an empty query set) give exactly the results of 72 single searches -- checked against the oracle."""
    rng = np.random.default_rng(606)
    frames = {}
    for name in ("kitti", "euroc"):
        ka, da, kb, db, bounds = pair[name]
        frames[name] = (ka, da, kb, db) + _frames(matcher, oracle, kb, db, bounds)
    jobs, want = [], []
    for j in range(72):
        name = "kitti" if j % 3 else "euroc"
        ka, da, kb, db, g, o = frames[name]
        q = _proj_queries(rng, ka)
        take = len(ka) if j % 5 else int(rng.integers(0, len(ka) // 2))        # j = 0: possibly empty
        if j == 7:
            take = 0
        q, qd = q[:take].copy(), da[:take].copy()
        occ = (rng.random(len(kb)) < 0.1).astype(np.uint8) if j % 2 else None
        ur = None
        if j % 4 == 1:
            ur = np.where(rng.random(len(kb)) < 0.6, kb["x"] - 40.0 * rng.uniform(0.02, 0.5, len(kb)), -1).astype(np.float32)
        jobs.append((g, q, qd, occ, ur))
        want.append(o.search_projection(SF, _oq(q), qd, 15.0, 0, occ, ur, 40.0, True) if take else (0, np.full(len(kb), -1, np.int32)))
    got, cand = matcher.search_by_projection_batch(jobs, SF, 15.0, mode=0, mbf=40.0, check_ori=True)
    assert cand > 100000
    for j, ((n, match), (rn, rmatch)) in enumerate(zip(got, want)):
        assert n == rn, j
        assert np.array_equal(match, rmatch), j
    assert sum(n for n, _ in got) > 10000
    # a batch of one is the single call
    n1, m1 = matcher.search_by_projection(jobs[1][0], SF, jobs[1][1], jobs[1][2], 15.0, 0, jobs[1][3], jobs[1][4], 40.0, True)
    assert n1 == got[1][0] and np.array_equal(m1, got[1][1])


def test_search_for_initialization_batch(matcher, oracle, pair):
    """orbm_search_for_initialization_batch over 64 frame pairs (both scenes, both directions, perturbed vbPrevMatched)."""
    rng = np.random.default_rng(707)
    pairs, want = [], []
    fr = {}
    for name in ("kitti", "euroc"):
        ka, da, kb, db, bounds = pair[name]
        fr[name] = (ka, kb) + _frames(matcher, oracle, ka, da, bounds) + _frames(matcher, oracle, kb, db, bounds)
    for j in range(64):
        ka, kb, g1, o1, g2, o2 = fr["kitti" if j % 2 else "euroc"]
        if j % 4 >= 2:
            ka, kb, g1, o1, g2, o2 = kb, ka, g2, o2, g1, o1
        prev = (np.stack([ka["x"], ka["y"]], 1) + rng.normal(0, 4.0 * (j % 3), (len(ka), 2))).astype(np.float32)
        pairs.append((g1, g2, prev))
        want.append(o1.search_init(o2, prev, 100, 0.9, True))
    got, cand = matcher.search_for_initialization_batch(pairs, 100, 0.9, True)
    assert cand > 100000
    for j, ((n, m12, p), (rn, rm12, rp)) in enumerate(zip(got, want)):
        assert n == rn and np.array_equal(m12, rm12) and np.array_equal(p, rp), j
    assert sum(n for n, _, _ in got) > 5000
