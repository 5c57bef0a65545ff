"""Seeded inputs for the ORBmatcher searches, shared by
  * tests/test_matcher_ref.py   -- oracle restatement vs the reference's own ORBmatcher.cc (oracle/_ref/libmatch_ref.so),
  * tools/gen_golden.py         -- golden outputs of that reference build (tests/golden/matcher_ref_vectors.npz),
  * tests/test_golden_gpu.py    -- the CUDA path against those goldens.
A "backend" is anything with the OracleFrame call surface (oracle_py.OracleFrame, ref_matcher.RefFrame, or the thin
wrapper around orbb200.Matcher below).  Projection queries carry invz == 1: the reference build is driven with an identity
camera at unit depth, for which its own projection arithmetic reproduces the given coordinates bit for bit."""
import numpy as np

SF = np.array([1.2 ** i for i in range(8)], np.float32)
PROJ = np.dtype([("u", "<f4"), ("v", "<f4"), ("invz", "<f4"), ("octave", "<i4"), ("valid", "<i4"), ("obsPositive", "<i4"),
                 ("angle", "<f4")])
POINT = np.dtype([("projX", "<f4"), ("projY", "<f4"), ("projXR", "<f4"), ("viewCos", "<f4"), ("level", "<i4"),
                  ("inView", "<i4"), ("obsPositive", "<i4")])


def feature_vector(keys, dx, dy, cell=48):
    """Stand-in for DBoW2::FeatureVector (node id -> keypoint indices, ascending ids): a coarse spatial hash."""
    node = ((keys["y"] + dy) // cell).astype(np.int64) * 100 + ((keys["x"] + dx) // cell).astype(np.int64)
    ids = np.unique(node)
    start, idx = [0], []
    for i in ids:
        idx.extend(np.nonzero(node == i)[0].tolist())
        start.append(len(idx))
    return ids.astype(np.int32), np.array(start, np.int32), np.array(idx, np.int32)


CASES = ["init_w100", "init_w30_noori", "init_ratio06", "proj_window", "proj_forward_stereo", "proj_backward_stereo",
         "points_th1", "points_th3", "tri_mono", "tri_only_stereo", "bow_kf_frame", "bow_kf_kf", "bow_kf_kf_noori",
         "reloc_orbdist64", "reloc_orbdist100_noori", "loop_window", "fuse_mono", "fuse_stereo", "fuse_scw", "sim3",
         "projw_window", "projw_forward_stereo", "projw_backward_stereo"]
WORLD = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("octave", "<i4"), ("valid", "<i4"), ("obsPositive", "<i4"),
                  ("angle", "<f4")])


def camera_pose(rng):
    """A non-trivial current pose (Rcw, tcw) in float32 and EuRoC-like intrinsics (fx, fy, cx, cy)."""
    ax = rng.normal(0, 1, 3)
    ax /= np.linalg.norm(ax)
    ang = 0.35
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
    t = np.array([0.31, -0.12, 0.57])
    return R.astype(np.float32), t.astype(np.float32), np.array([458.654, 457.296, 367.215, 248.375], np.float32)
BEST = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("ur", "<f4"), ("level", "<i4"), ("valid", "<i4")])
TH_LOW, TH_HIGH = 50, 100


def run_case(name, f1, f2, ka, da, kb, db, shift=(7.0, 3.0)):
    """f1 / f2: backend frames built from (ka, da) / (kb, db). Returns a tuple of arrays / ints."""
    dx, dy = shift                      # a keypoint of the first view sits at (+dx, +dy) in the second
    rng = np.random.default_rng(sum(map(ord, name)))
    n1, n2 = len(ka), len(kb)
    if name.startswith("init"):
        window, ratio, ori = {"init_w100": (100, 0.9, True), "init_w30_noori": (30, 0.9, False), "init_ratio06": (100, 0.6, True)}[name]
        prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)       # Tracking.cc:1637-1639
        n, m12, p = f1.search_init(f2, prev, window, ratio, ori)
        n_b, m12_b, p_b = f1.search_init(f2, p, window, ratio, ori)      # next frame: vbPrevMatched as updated
        return n, m12, p, n_b, m12_b, p_b
    if name.startswith("projw"):
        # SearchByProjection(Current, Last, th, bMono) from WORLD points through a real pose and intrinsics: the reference's
        # own projection lines (:1376-1393) produce the window centres; points behind the camera and outside the image included
        th, mode, stereo, ori = {"projw_window": (15.0, 0, False, True), "projw_forward_stereo": (7.0, 1, True, True),
                                 "projw_backward_stereo": (15.0, 2, True, False)}[name]
        R, t, K4 = camera_pose(rng)
        u = ka["x"].astype(np.float64) + dx + rng.normal(0, 1.5, n1)
        v = ka["y"].astype(np.float64) + dy + rng.normal(0, 1.5, n1)
        z = rng.uniform(9.5, 10.5, n1)
        z[3:40:4] *= -1                                    # behind the camera: invzc < 0
        u[:3] = [-500, 1e4, 10]
        v[:3] = [10, 10, -800]                             # outside the image bounds
        xc = np.stack([(u - K4[2]) / K4[0] * z, (v - K4[3]) / K4[1] * z, z], 1)
        xw = (xc - t.astype(np.float64)) @ R.astype(np.float64)          # R^T (xc - t)
        q = np.zeros(n1, WORLD)
        q["x"], q["y"], q["z"] = xw[:, 0], xw[:, 1], xw[:, 2]
        q["octave"] = ka["octave"]
        q["valid"] = rng.random(n1) < 0.8
        q["obsPositive"] = rng.random(n1) < 0.9
        q["angle"] = ka["angle"]
        occ = (rng.random(n2) < 0.1).astype(np.uint8)
        ur, mbf = None, 0.0
        if stereo:
            mbf = 40.0
            ur = np.where(rng.random(n2) < 0.6, kb["x"] - mbf / 10.0 + rng.normal(0, th / 2, n2), -1).astype(np.float32)
        return f2.search_projection_world(SF, R, t, K4, q, da, th, mode, occ, ur, mbf, ori)
    if name.startswith("proj"):
        th, mode, stereo, ori = {"proj_window": (15.0, 0, False, True), "proj_forward_stereo": (7.0, 1, True, True),
                                 "proj_backward_stereo": (15.0, 2, True, False)}[name]
        q = np.zeros(n1, PROJ)
        q["u"] = ka["x"] + dx + rng.normal(0, 1.5, n1)
        q["v"] = ka["y"] + dy + rng.normal(0, 1.5, n1)
        q["invz"] = 1.0
        q["octave"] = ka["octave"]
        q["valid"] = rng.random(n1) < 0.8
        q["obsPositive"] = rng.random(n1) < 0.9
        q["angle"] = ka["angle"]
        q["u"][:3] = [-5, 1e4, 10]
        q["v"][:3] = [10, 10, -8]
        occ = (rng.random(n2) < 0.1).astype(np.uint8)
        ur, mbf = None, 0.0
        if stereo:
            mbf = 40.0
            ur = np.where(rng.random(n2) < 0.6, kb["x"] - mbf + rng.normal(0, th / 2, n2), -1).astype(np.float32)
        return f2.search_projection(SF, q, da, th, mode, occ, ur, mbf, ori)
    if name.startswith("points"):
        th, ratio = {"points_th1": (1.0, 0.8), "points_th3": (3.0, 0.6)}[name]
        q = np.zeros(n1, POINT)
        q["projX"] = ka["x"] + dx + rng.normal(0, 1.0, n1)
        q["projY"] = ka["y"] + dy + rng.normal(0, 1.0, n1)
        q["projXR"] = q["projX"] - 5
        q["viewCos"] = rng.choice([0.9, 0.998, 0.9985, 1.0], n1)
        q["level"] = np.clip(ka["octave"] + rng.integers(0, 2, n1), 0, 7)
        q["inView"] = rng.random(n1) < 0.85
        q["obsPositive"] = rng.random(n1) < 0.9
        occ = (rng.random(n2) < 0.05).astype(np.uint8)
        ur = np.where(rng.random(n2) < 0.3, kb["x"] - 5 + rng.normal(0, 3, n2), -1).astype(np.float32)
        return f2.search_points(SF, q, da, th, ratio, occ, ur)
    if name.startswith("tri"):
        only_stereo, ori = {"tri_mono": (False, False), "tri_only_stereo": (True, True)}[name]
        fv1, fv2 = feature_vector(ka, 0, 0), feature_vector(kb, -dx, -dy)
        F12 = np.array([[0, 0, -dy], [0, 0, dx], [dy, -dx, 0]], np.float32) * 1e-2    # pure image translation: F12 = [t]x
        has1 = (rng.random(n1) < 0.3).astype(np.uint8)
        has2 = (rng.random(n2) < 0.3).astype(np.uint8)
        ur1 = np.where(rng.random(n1) < 0.5, 10.0, -1.0).astype(np.float32) if only_stereo else None
        ur2 = np.where(rng.random(n2) < 0.5, 10.0, -1.0).astype(np.float32) if only_stereo else None
        return f1.search_triangulation(f2, fv1, fv2, F12, 3000.0, 200.0, SF, (SF * SF).astype(np.float32), has1, has2, ur1,
                                       ur2, only_stereo, ori)
    if name.startswith("bow"):
        strict, ratio, ori, use_valid2 = {"bow_kf_frame": (False, 0.7, True, False), "bow_kf_kf": (True, 0.75, True, True),
                                          "bow_kf_kf_noori": (True, 0.9, False, True)}[name]
        fv1, fv2 = feature_vector(ka, 0, 0, cell=64), feature_vector(kb, -dx, -dy, cell=64)
        valid1 = (rng.random(n1) < 0.7).astype(np.uint8)
        valid2 = (rng.random(n2) < 0.8).astype(np.uint8) if use_valid2 else None
        return f1.search_bow(f2, fv1, fv2, valid1, valid2, ratio, ori, strict)
    if name.startswith("reloc") or name == "loop_window":
        # SearchByProjection(Frame, KeyFrame, sAlreadyFound, th, ORBdist) (:1500) and SearchByProjection(KF, Scw, ...) (:290):
        # the level is the one MapPoint::PredictScale returns; an assigned keypoint is simply taken afterwards
        q = np.zeros(n1, PROJ)
        q["u"] = ka["x"] + dx + rng.normal(0, 1.5, n1)
        q["v"] = ka["y"] + dy + rng.normal(0, 1.5, n1)
        q["invz"] = 1.0
        q["octave"] = np.clip(ka["octave"] + rng.integers(-1, 2, n1), 0, 7)
        q["valid"] = rng.random(n1) < 0.8
        q["obsPositive"] = 1
        q["angle"] = ka["angle"]
        occ = (rng.random(n2) < 0.1).astype(np.uint8)
        if name == "loop_window":
            if hasattr(f2, "search_projection_sim3"):
                return f2.search_projection_sim3(SF, q, da, 10, occ)
            return f2.search_projection(SF, q, da, 10.0, 3, occ, None, 0.0, False, TH_LOW)
        th, orb_dist, ori = {"reloc_orbdist64": (10.0, 64, True), "reloc_orbdist100_noori": (3.0, 100, False)}[name]
        if hasattr(f2, "search_projection_kf"):
            return f2.search_projection_kf(SF, q, da, th, orb_dist, occ, ori)
        return f2.search_projection(SF, q, da, th, 0, occ, None, 0.0, ori, orb_dist)
    if name.startswith("fuse"):
        # Fuse (:825 with the chi-square gate, :977 without): the keyframe has no map points, so "fused into keypoint k"
        # is all the reference does with a hit; the GPU / oracle side returns the closest keypoint and its distance
        chi2, stereo, th = {"fuse_mono": (True, False, 3.0), "fuse_stereo": (True, True, 3.0), "fuse_scw": (False, False, 4.0)}[name]
        bf = 40.0 if stereo else 0.0
        q = np.zeros(n1, BEST)
        q["u"] = (ka["x"] + dx + rng.normal(0, 1.0, n1)).astype(np.float32)
        q["v"] = (ka["y"] + dy + rng.normal(0, 1.0, n1)).astype(np.float32)
        q["level"] = np.clip(ka["octave"] + rng.integers(0, 2, n1), 0, 7)
        q["radius"] = np.float32(th) * SF[q["level"]]
        q["ur"] = q["u"] - np.float32(bf)
        q["valid"] = rng.random(n1) < 0.85
        inv_s2 = (1.0 / (SF * SF)).astype(np.float32)
        ur = np.where(rng.random(n2) < 0.5, kb["x"] - bf + rng.normal(0, 1.0, n2), -1).astype(np.float32) if stereo else None
        if hasattr(f2, "fuse"):
            return f2.fuse(SF, inv_s2, q, da, th, ur, bf, not chi2)
        bi, bd = f2.search_best(q, da, chi2, ur, inv_s2)
        fused = np.where(bd <= TH_LOW, bi, -1).astype(np.int32)
        return int((fused >= 0).sum()), fused
    if name == "sim3":
        # SearchBySim3 (:1102) with the identity Sim3: the points of each keyframe project into the other one at (u, v)
        th = 7.5
        uv1 = np.stack([ka["x"] + dx + rng.normal(0, 1.0, n1), ka["y"] + dy + rng.normal(0, 1.0, n1)], 1).astype(np.float32)
        uv2 = np.stack([kb["x"] - dx + rng.normal(0, 1.0, n2), kb["y"] - dy + rng.normal(0, 1.0, n2)], 1).astype(np.float32)
        l1 = np.clip(ka["octave"] + rng.integers(0, 2, n1), 0, 7).astype(np.int32)
        l2 = np.clip(kb["octave"] + rng.integers(0, 2, n2), 0, 7).astype(np.int32)
        has1 = (rng.random(n1) < 0.8).astype(np.uint8)
        has2 = (rng.random(n2) < 0.8).astype(np.uint8)
        if hasattr(f1, "search_sim3"):
            return f1.search_sim3(f2, SF, SF, uv1, l1, has1, uv2, l2, has2, th)

        def direction(frame, uv, lev, has, qdesc):
            q = np.zeros(len(uv), BEST)
            q["u"], q["v"], q["level"], q["valid"] = uv[:, 0], uv[:, 1], lev, has
            q["radius"] = np.float32(th) * SF[lev]
            bi, bd = frame.search_best(q, qdesc, False, None, None)
            return np.where((bd <= TH_HIGH) & (has != 0), bi, -1)
        v1 = direction(f2, uv1, l1, has1, da)        # points of keyframe 1 searched in keyframe 2 (:1141-1219)
        v2 = direction(f1, uv2, l2, has2, db)        # and the other way round (:1222-1299)
        m12 = np.full(n1, -1, np.int32)
        for i1 in range(n1):                           # mutual agreement (:1302-1318)
            if v1[i1] >= 0 and v2[v1[i1]] == i1:
                m12[i1] = v1[i1]
        return int((m12 >= 0).sum()), m12
    raise KeyError(name)


def same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        if isinstance(x, np.ndarray):
            assert np.array_equal(x, y)
        else:
            assert int(x) == int(y)


class GpuFrame:
    """orbb200.Matcher behind the OracleFrame call surface."""

    def __init__(self, matcher, keys, desc, bounds):
        self.m = matcher
        self.f = matcher.frame(keys, desc, bounds)

    def search_init(self, other, prev_xy, window=100, ratio=0.9, check_ori=True):
        return self.m.search_for_initialization(self.f, other.f, prev_xy, window, ratio, check_ori)

    def search_projection(self, sf, q, qdesc, th, mode=0, occupied=None, u_right=None, mbf=0.0, check_ori=True,
                          max_distance=100):
        import orbb200
        return self.m.search_by_projection(self.f, sf, q.view(orbb200.PROJ_QUERY_DTYPE), qdesc, th, mode, occupied, u_right,
                                           mbf, check_ori, max_distance)

    def search_projection_world(self, sf, Rcw, tcw, K4, q, qdesc, th, mode=0, occupied=None, u_right=None, mbf=0.0,
                                check_ori=True, max_distance=100):
        import orbb200
        return self.m.search_by_projection_world(self.f, sf, Rcw, tcw, K4, q.view(orbb200.WORLD_QUERY_DTYPE), qdesc, th, mode,
                                                 occupied, u_right, mbf, check_ori, max_distance)

    def search_best(self, q, qdesc, chi2=False, u_right=None, inv_sigma2=None):
        import orbb200
        return self.m.search_projected_best(self.f, q.view(orbb200.BEST_QUERY_DTYPE), qdesc, chi2, u_right, inv_sigma2)

    def search_points(self, sf, q, qdesc, th, ratio, occupied=None, u_right=None):
        import orbb200
        return self.m.search_by_projection_points(self.f, sf, q.view(orbb200.POINT_QUERY_DTYPE), qdesc, th, ratio, occupied,
                                                  u_right)

    def search_triangulation(self, other, fv1, fv2, F12, ex, ey, sf2, sigma2_2, has1=None, has2=None, ur1=None, ur2=None,
                             only_stereo=False, check_ori=False):
        return self.m.search_for_triangulation(self.f, other.f, fv1, fv2, F12, ex, ey, sf2, sigma2_2, has1, has2, ur1, ur2,
                                               only_stereo, check_ori)

    def search_bow(self, other, fv1, fv2, valid1=None, valid2=None, ratio=0.7, check_ori=True, strict_low=False):
        return self.m.search_by_bow(self.f, other.f, fv1, fv2, valid1, valid2, ratio, check_ori, strict_low)
