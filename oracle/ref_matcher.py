"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libmatch_ref.so: the reference's own ORBmatcher.cc
(compiled in place against oracle/ref_shim/matcher) behind the same flat-array calls as oracle_py.OracleFrame, so a test
hands identical inputs to the oracle restatement and to the reference's code."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle_py import KP_DTYPE, MP_QUERY_DTYPE, PROJ_QUERY_DTYPE, WORLD_QUERY_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libmatch_ref.so")
ADAPTER_LIB = os.path.join(HERE, "_ref", "libmatch_adapter.so")   # same entry points (adpm_*), ORBmatcher = the CUDA adapter


def build():
    """Needs /root/reference (absent on the GPU box, where the prebuilt file is used)."""
    subprocess.check_call(["make", "-s", "-f", os.path.join(HERE, "ref_shim", "Makefile"), LIB])
    return LIB


def available():
    return os.path.exists(LIB)


def adapter_available():
    return os.path.exists(ADAPTER_LIB)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


vp, i32, f32 = C.c_void_p, C.c_int, C.c_float


class _Prefixed:
    """lib.refm_x -> getattr(lib, prefix + x): the two libraries export the same functions under different prefixes."""

    def __init__(self, lib, prefix):
        object.__setattr__(self, "_lib", lib)
        object.__setattr__(self, "_prefix", prefix)

    def __getattr__(self, name):
        return getattr(self._lib, self._prefix + name[len("refm_"):] if name.startswith("refm_") else name)


class RefMatcher:
    def __init__(self, adapter=False):
        self.lib = L = _Prefixed(C.CDLL(ADAPTER_LIB if adapter else LIB), "adpm_" if adapter else "refm_")
        L.refm_distance.argtypes = [vp, vp]
        L.refm_frame_create.restype = vp
        L.refm_frame_create.argtypes = [vp, vp, i32, f32, f32, f32, f32]
        L.refm_frame_destroy.argtypes = [vp]
        L.refm_search_init.argtypes = [vp, vp, vp, vp, i32, f32, i32]
        L.refm_search_projection.argtypes = [vp, vp, i32, vp, f32, vp, vp, i32, f32, i32, vp, vp, i32]
        L.refm_search_points.argtypes = [vp, vp, i32, vp, vp, vp, i32, f32, f32, vp, vp]
        L.refm_search_triangulation.argtypes = ([vp, vp] + [i32, vp, vp, vp] * 2 + [vp] * 5 + [f32, f32, vp, vp, i32, i32, i32, vp])
        L.refm_search_bow.argtypes = ([vp, vp] + [i32, vp, vp, vp] * 2 + [vp, vp, f32, i32, i32, vp, vp])
        L.refm_search_projection_world.argtypes = [vp, vp, i32, vp, f32, vp, vp, vp, vp, i32, f32, i32, vp, vp, i32]
        L.refm_search_projection_kf.argtypes = [vp, vp, i32, vp, vp, i32, f32, i32, vp, vp, i32]
        L.refm_search_projection_sim3.argtypes = [vp, vp, i32, vp, vp, i32, i32, vp, vp]
        L.refm_fuse.argtypes = [vp, vp, vp, i32, vp, f32, vp, vp, i32, f32, i32, vp]
        L.refm_search_sim3.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, f32, vp]
        if not adapter:   # the reference's own grid (Frame.cc / KeyFrame.cc text, oracle/ref_shim/grid)
            L._lib.refm_grid_csr.argtypes = [vp, i32, f32, f32, f32, f32, vp, vp]
            L._lib.refm_features_in_area.argtypes = [vp, i32, f32, f32, f32, f32, i32, f32, f32, f32, i32, i32, vp, i32]

    def grid_csr(self, keys_un, bounds):
        """Frame::AssignFeaturesToGrid (Frame.cc:574-589) as CSR over cells ix * 48 + iy, members in push order."""
        k = np.ascontiguousarray(keys_un, KP_DTYPE)
        start = np.empty(64 * 48 + 1, np.int32)
        idx = np.empty(max(len(k), 1), np.int32)
        n = self.lib._lib.refm_grid_csr(_p(k), len(k), *[float(b) for b in bounds], _p(start), _p(idx))
        return start, idx[:n]

    def features_in_area(self, keys_un, bounds, x, y, r, min_level=-1, max_level=-1, keyframe=False):
        """Frame::GetFeaturesInArea (Frame.cc:671-724) or, keyframe=True, KeyFrame::GetFeaturesInArea (KeyFrame.cc:1138-1177)."""
        k = np.ascontiguousarray(keys_un, KP_DTYPE)
        out = np.empty(max(len(k), 1), np.int32)
        n = self.lib._lib.refm_features_in_area(_p(k), len(k), *[float(b) for b in bounds], int(keyframe), float(x), float(y),
                                                float(r), int(min_level), int(max_level), _p(out), len(out))
        return out[:n]

    def distance(self, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return self.lib.refm_distance(_p(a), _p(b))

    def frame(self, keys_un, desc, bounds):
        return RefFrame(self, keys_un, desc, bounds)


def adapter_chain_init(img_a, img_b, nfeatures=1000, window=100, ratio=0.9, cap=4096):
    """adpm_chain_init (adapter build only): extraction -> RegisterFrame -> SearchForInitialization twice on cached device
    frames.  Returns (keysA, descA, keysB, descB, matches12 of call 1, n1, n2, cache hits, cache misses)."""
    L = C.CDLL(ADAPTER_LIB)
    h, w = img_a.shape
    ka = np.zeros(cap, KP_DTYPE); kb = np.zeros(cap, KP_DTYPE)
    da = np.zeros((cap, 32), np.uint8); db = np.zeros((cap, 32), np.uint8)
    m12 = np.full(cap, -1, np.int32)
    counts = np.zeros(4, np.int32)
    stats = np.zeros(2, np.int64)
    a = np.ascontiguousarray(img_a, np.uint8); b = np.ascontiguousarray(img_b, np.uint8)
    L.adpm_chain_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float] + [C.c_void_p] * 4 + \
                                 [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.adpm_chain_init(_p(a), _p(b), w, h, nfeatures, window, ratio, _p(ka), _p(da), _p(kb), _p(db), cap, _p(m12), _p(counts),
                           _p(stats))
    assert rc == 0
    na, nb = int(counts[0]), int(counts[1])
    return ka[:na], da[:na], kb[:nb], db[:nb], m12[:na], int(counts[2]), int(counts[3]), int(stats[0]), int(stats[1])


class RefFrame:
    def __init__(self, ref, keys_un, desc, bounds):
        self.lib = ref.lib
        self.keys = np.ascontiguousarray(keys_un, KP_DTYPE)
        self.desc = np.ascontiguousarray(desc, np.uint8)
        self.n = len(self.keys)
        self.h = C.c_void_p(self.lib.refm_frame_create(_p(self.keys), _p(self.desc), self.n, *[float(b) for b in bounds]))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.refm_frame_destroy(self.h)
            self.h = None

    def search_init(self, other, prev_xy, window=100, ratio=0.9, check_ori=True):
        prev = np.ascontiguousarray(prev_xy, np.float32).copy()
        m12 = np.empty(self.n, np.int32)
        n = self.lib.refm_search_init(self.h, other.h, _p(prev), _p(m12), window, ratio, int(check_ori))
        return n, m12, prev

    def search_projection(self, scale_factors, queries, qdesc, th, mode=0, occupied=None, u_right=None, mbf=0.0, check_ori=True):
        """queries must carry invz == 1 (identity camera at unit depth, see match_ref.cpp)."""
        sf = np.ascontiguousarray(scale_factors, np.float32)
        q = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        occ = np.zeros(self.n, np.uint8) if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        match = np.empty(self.n, np.int32)
        n = self.lib.refm_search_projection(self.h, _p(sf), len(sf), _p(ur), mbf, _p(q), _p(qd), len(q), th, mode, _p(occ),
                                            _p(match), int(check_ori))
        assert n != -2, "refm_search_projection needs invz == 1"
        return n, match

    def search_projection_world(self, scale_factors, Rcw, tcw, K4, queries, qdesc, th, mode=0, occupied=None, u_right=None,
                                mbf=0.0, check_ori=True):
        """SearchByProjection(Current, Last, th, bMono) with world points, a real pose and intrinsics: the reference's own
        projection lines (ORBmatcher.cc:1376-1393) run."""
        sf = np.ascontiguousarray(scale_factors, np.float32)
        q = np.ascontiguousarray(queries, WORLD_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        pose = np.concatenate([np.asarray(Rcw, np.float32).ravel(), np.asarray(tcw, np.float32).ravel()])
        k4 = np.asarray(K4, np.float32)
        occ = np.zeros(self.n, np.uint8) if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        match = np.empty(self.n, np.int32)
        n = self.lib.refm_search_projection_world(self.h, _p(sf), len(sf), _p(ur), mbf, _p(pose), _p(k4), _p(q), _p(qd), len(q),
                                                  th, mode, _p(occ), _p(match), int(check_ori))
        return n, match

    def search_points(self, scale_factors, queries, qdesc, th, ratio, occupied=None, u_right=None):
        sf = np.ascontiguousarray(scale_factors, np.float32)
        q = np.ascontiguousarray(queries, MP_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        occ = np.zeros(self.n, np.uint8) if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        match = np.empty(self.n, np.int32)
        n = self.lib.refm_search_points(self.h, _p(sf), len(sf), _p(ur), _p(q), _p(qd), len(q), th, ratio, _p(occ), _p(match))
        return n, match

    def search_triangulation(self, other, fv1, fv2, F12, ex, ey, sf2, sigma2_2, has1=None, has2=None, ur1=None, ur2=None,
                             only_stereo=False, check_ori=False):
        n1, s1, i1 = [np.ascontiguousarray(a, np.int32) for a in fv1]
        n2, s2, i2 = [np.ascontiguousarray(a, np.int32) for a in fv2]
        has1 = np.zeros(self.n, np.uint8) if has1 is None else np.ascontiguousarray(has1, np.uint8)
        has2 = np.zeros(other.n, np.uint8) if has2 is None else np.ascontiguousarray(has2, np.uint8)
        ur1 = None if ur1 is None else np.ascontiguousarray(ur1, np.float32)
        ur2 = None if ur2 is None else np.ascontiguousarray(ur2, np.float32)
        F12 = np.ascontiguousarray(F12, np.float32)
        sf2 = np.ascontiguousarray(sf2, np.float32)
        sg2 = np.ascontiguousarray(sigma2_2, np.float32)
        m12 = np.empty(self.n, np.int32)
        n = self.lib.refm_search_triangulation(self.h, other.h, len(n1), _p(n1), _p(s1), _p(i1), len(n2), _p(n2), _p(s2),
                                               _p(i2), _p(has1), _p(has2), _p(ur1), _p(ur2), _p(F12), ex, ey, _p(sf2),
                                               _p(sg2), len(sf2), int(only_stereo), int(check_ori), _p(m12))
        return n, m12

    def search_bow(self, other, fv1, fv2, valid1=None, valid2=None, ratio=0.7, check_ori=True, strict_low=False):
        n1, s1, i1 = [np.ascontiguousarray(a, np.int32) for a in fv1]
        n2, s2, i2 = [np.ascontiguousarray(a, np.int32) for a in fv2]
        v1 = None if valid1 is None else np.ascontiguousarray(valid1, np.uint8)
        v2 = None if valid2 is None else np.ascontiguousarray(valid2, np.uint8)
        m12 = np.empty(self.n, np.int32)
        m21 = np.empty(other.n, np.int32)
        n = self.lib.refm_search_bow(self.h, other.h, len(n1), _p(n1), _p(s1), _p(i1), len(n2), _p(n2), _p(s2), _p(i2),
                                     _p(v1), _p(v2), ratio, int(check_ori), int(strict_low), _p(m12), _p(m21))
        return n, m12, m21

    # ---- the overloads whose GPU entry points are orbm_search_by_projection_ex / orbm_search_projected_best
    def search_projection_kf(self, scale_factors, queries, qdesc, th, orb_dist, occupied=None, check_ori=True):
        """SearchByProjection(Frame, KeyFrame, sAlreadyFound, th, ORBdist), ORBmatcher.cc:1500"""
        sf = np.ascontiguousarray(scale_factors, np.float32)
        q = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        occ = np.zeros(self.n, np.uint8) if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        match = np.empty(self.n, np.int32)
        n = self.lib.refm_search_projection_kf(self.h, _p(sf), len(sf), _p(q), _p(qd), len(q), th, orb_dist, _p(occ), _p(match),
                                               int(check_ori))
        assert n != -2
        return n, match

    def search_projection_sim3(self, scale_factors, queries, qdesc, th, occupied=None):
        """SearchByProjection(KeyFrame, Scw, vpPoints, vpMatched, th), ORBmatcher.cc:290"""
        sf = np.ascontiguousarray(scale_factors, np.float32)
        q = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        occ = np.zeros(self.n, np.uint8) if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        match = np.empty(self.n, np.int32)
        n = self.lib.refm_search_projection_sim3(self.h, _p(sf), len(sf), _p(q), _p(qd), len(q), int(th), _p(occ), _p(match))
        assert n != -2
        return n, match

    def fuse(self, scale_factors, inv_sigma2, queries, qdesc, th, u_right=None, bf=0.0, scw=False):
        """Fuse (ORBmatcher.cc:825 / :977 with scw=True) -> (nFused, keypoint index each point was fused into or -1)"""
        from oracle_py import BEST_QUERY_DTYPE
        sf = np.ascontiguousarray(scale_factors, np.float32)
        s2 = np.ascontiguousarray(inv_sigma2, np.float32)
        q = np.ascontiguousarray(queries, BEST_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        fused = np.empty(len(q), np.int32)
        n = self.lib.refm_fuse(self.h, _p(sf), _p(s2), len(sf), _p(ur), bf, _p(q), _p(qd), len(q), th, int(scw), _p(fused))
        return n, fused

    def search_sim3(self, other, sf1, sf2, uv1, level1, has1, uv2, level2, has2, th):
        """SearchBySim3 (ORBmatcher.cc:1102) with identity Sim3 and poses -> (nFound, matches12)"""
        sf1 = np.ascontiguousarray(sf1, np.float32); sf2 = np.ascontiguousarray(sf2, np.float32)
        uv1 = np.ascontiguousarray(uv1, np.float32); uv2 = np.ascontiguousarray(uv2, np.float32)
        level1 = np.ascontiguousarray(level1, np.int32); level2 = np.ascontiguousarray(level2, np.int32)
        has1 = np.ascontiguousarray(has1, np.uint8); has2 = np.ascontiguousarray(has2, np.uint8)
        m12 = np.empty(self.n, np.int32)
        n = self.lib.refm_search_sim3(self.h, other.h, _p(sf1), _p(sf2), len(sf1), _p(uv1), _p(level1), _p(has1), _p(uv2),
                                      _p(level2), _p(has2), th, _p(m12))
        return n, m12
