"""TEST INFRASTRUCTURE ONLY -- ctypes view of the CPU oracle (oracle/liborb_oracle.so).

Importable from tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() only; the product
package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

PROJ_QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("invz", "<f4"), ("octave", "<i4"), ("valid", "<i4"),
                             ("obsPositive", "<i4"), ("angle", "<f4")])
WORLD_QUERY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("octave", "<i4"), ("valid", "<i4"),
                              ("obsPositive", "<i4"), ("angle", "<f4")])
MP_QUERY_DTYPE = np.dtype([("projX", "<f4"), ("projY", "<f4"), ("projXR", "<f4"), ("viewCos", "<f4"),
                           ("level", "<i4"), ("inView", "<i4"), ("obsPositive", "<i4")])


def build(native=False):
    """Compile the oracle (idempotent). Building the checker is not using it."""
    target = os.path.join(HERE, "liborb_oracle_native.so" if native else "liborb_oracle.so")
    subprocess.check_call(["make", "-s", "-f", os.path.join(HERE, "Makefile"), target])
    return target


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


class Oracle:
    def __init__(self, native=False):
        path = os.path.join(HERE, "liborb_oracle_native.so" if native else "liborb_oracle.so")
        if not os.path.exists(path):
            build(native)
        self.lib = L = C.CDLL(path)
        L.orbo_create.restype = C.c_void_p
        L.orbo_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbo_destroy.argtypes = [C.c_void_p]
        L.orbo_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orbo_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orbo_level_size.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orbo_level_padded.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orbo_level_blurred.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orbo_level_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orbo_level_selected.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orbo_stage_ms.argtypes = [C.c_void_p, C.c_void_p]
        L.orbo_undistort.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orbo_distinctive.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orbo_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                  C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbo_atan2.restype = C.c_float
        L.orbo_atan2.argtypes = [C.c_float, C.c_float]
        L.orbo_cosf.restype = C.c_float
        L.orbo_cosf.argtypes = [C.c_float]
        L.orbo_sinf.restype = C.c_float
        L.orbo_sinf.argtypes = [C.c_float]
        L.orbo_round.argtypes = [C.c_float]
        L.orbo_pattern.restype = C.c_void_p
        L.orbo_ic_angle.restype = C.c_float
        L.orbo_ic_angle.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orbo_descriptor.argtypes = [C.c_float, C.c_void_p, C.c_int, C.c_void_p]
        L.orbo_frame_create.restype = C.c_void_p
        L.orbo_frame_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
        L.orbo_frame_destroy.argtypes = [C.c_void_p]
        L.orbo_frame_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbo_frame_area.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orbo_search_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int]
        L.orbo_search_projection.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                             C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orbo_search_projection_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                                C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orbo_search_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orbo_search_triangulation.argtypes = ([C.c_void_p, C.c_void_p] + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] * 2
                                                + [C.c_void_p] * 5 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                                                     C.c_int, C.c_int, C.c_void_p])
        L.orbo_bruteforce.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbo_kf_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                   C.c_int, C.c_void_p]
        L.orbo_distance.argtypes = [C.c_void_p, C.c_void_p]
        L.orbo_rotation_bin.argtypes = [C.c_float, C.c_float]
        L.orbo_distribute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]

    # ---- primitives ----
    def resize(self, src, dw, dh):
        src = np.ascontiguousarray(src, np.uint8)
        dst = np.empty((dh, dw), np.uint8)
        self.lib.orbo_resize(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
        return dst

    def resize_table(self, ssize, dsize):
        ofs = np.empty(dsize, np.int32)
        coef = np.empty((dsize, 2), np.int16)
        self.lib.orbo_resize_table(ssize, dsize, _p(ofs), _p(coef))
        return ofs, coef

    def border(self, src, b=19):
        src = np.ascontiguousarray(src, np.uint8)
        h, w = src.shape
        dst = np.empty((h + 2 * b, w + 2 * b), np.uint8)
        self.lib.orbo_border(_p(src), w, h, src.strides[0], _p(dst), w + 2 * b, b)
        return dst

    def fast(self, img, th):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = max(1, (w * h) // 2)
        out = np.zeros(cap, KP_DTYPE)
        n = self.lib.orbo_fast(_p(img), w, h, img.strides[0], th, _p(out), cap)
        return out[:n].copy()

    def blur(self, src):
        src = np.ascontiguousarray(src, np.uint8)
        h, w = src.shape
        dst = np.empty((h, w), np.uint8)
        self.lib.orbo_blur(_p(src), w, h, src.strides[0], _p(dst), w)
        return dst

    def undistort(self, xy, K4, dist):
        """cv::undistortPoints(pts, pts, K, D, noArray(), K); K4 = (fx, fy, cx, cy), dist = k1 k2 p1 p2 [k3]"""
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        K4 = np.ascontiguousarray(K4, np.float32)
        dist = np.ascontiguousarray(dist, np.float32).ravel()
        out = np.empty_like(xy)
        self.lib.orbo_undistort(_p(xy), len(xy), _p(K4), _p(dist), dist.size, _p(out))
        return out

    def atan2(self, y, x):
        y = np.ascontiguousarray(y, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(y)
        self.lib.orbo_atan2_n(_p(y), _p(x), _p(out), y.size)
        return out

    def sincosf(self, a):
        a = np.ascontiguousarray(a, np.float32)
        s = np.empty_like(a)
        c = np.empty_like(a)
        self.lib.orbo_sincosf_n(_p(a), _p(s), _p(c), a.size)
        return s, c

    def pattern(self):
        return np.ctypeslib.as_array(C.cast(self.lib.orbo_pattern(), C.POINTER(C.c_int8)), (1024,)).copy()

    def distribute(self, keys, minX, maxX, minY, maxY, N):
        keys = np.ascontiguousarray(keys, KP_DTYPE)
        cap = max(16, N * 4 + 64)
        out = np.zeros(cap, KP_DTYPE)
        n = self.lib.orbo_distribute(_p(keys), len(keys), minX, maxX, minY, maxY, N, _p(out), cap)
        assert n <= cap
        return out[:n].copy()

    def ic_angle(self, img, x, y, umax):
        img = np.ascontiguousarray(img, np.uint8)
        umax = np.ascontiguousarray(umax, np.int32)
        center = img.ctypes.data + y * img.strides[0] + x
        return float(self.lib.orbo_ic_angle(C.c_void_p(center), img.strides[0], _p(umax)))

    def descriptor(self, img, x, y, angle_deg):
        img = np.ascontiguousarray(img, np.uint8)
        out = np.empty(32, np.uint8)
        center = img.ctypes.data + y * img.strides[0] + x
        self.lib.orbo_descriptor(C.c_float(angle_deg), C.c_void_p(center), img.strides[0], _p(out))
        return out

    def distance(self, a, b):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        return int(self.lib.orbo_distance(_p(a), _p(b)))

    def three_maxima(self, sizes):
        sizes = np.ascontiguousarray(sizes, np.int32)
        out = np.empty(3, np.int32)
        self.lib.orbo_three_maxima(_p(sizes), len(sizes), _p(out))
        return tuple(int(v) for v in out)

    def rotation_bin(self, a1, a2):
        return int(self.lib.orbo_rotation_bin(C.c_float(a1), C.c_float(a2)))

    def extractor(self, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7):
        return OracleExtractor(self, nfeatures, scale, nlevels, ini, mn)

    def frame(self, keys_un, desc, bounds):
        return OracleFrame(self, keys_un, desc, bounds)

    def bow_transform(self, voc, desc, levelsup=4, lib=None, fn="orbo_bow_transform"):
        """TemplatedVocabulary::transform(features, v, fv, levelsup) on a flat vocabulary (tests/bow_cases.py::make_vocab).
        Returns dict(word, weight, node per feature; bow_word, bow_value; fv_node, fv_start, fv_idx)."""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        f = getattr(lib or self.lib, fn)
        f.restype = None
        f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_int] + [C.c_void_p] * 9
        pw = np.empty(n, np.int32); pwt = np.empty(n, np.float64); pn = np.empty(n, np.int32)
        bw = np.empty(n + 1, np.int32); bv = np.empty(n + 1, np.float64)
        fn_ = np.empty(n + 1, np.int32); fs = np.empty(n + 2, np.int32); fi = np.empty(n + 1, np.int32)
        cnt = np.zeros(3, np.int32)
        f(voc["n_nodes"], voc["L"], _p(voc["desc"]), _p(voc["child_start"]), _p(voc["children"]), _p(voc["word_id"]),
          _p(voc["weight"]), _p(d), n, levelsup, _p(pw), _p(pwt), _p(pn), _p(bw), _p(bv), _p(fn_), _p(fs), _p(fi), _p(cnt))
        return dict(word=pw, weight=pwt, node=pn, bow_word=bw[:cnt[0]].copy(), bow_value=bv[:cnt[0]].copy(),
                    fv_node=fn_[:cnt[1]].copy(), fv_start=fs[:cnt[1] + 1].copy(), fv_idx=fi[:cnt[2]].copy())

    def gemm3(self, A, b, c=None):
        """cv::Mat A(3x3) * b(3x1) [+ c] for CV_32F, as the matcher's projections evaluate it"""
        A = np.ascontiguousarray(A, np.float32); b = np.ascontiguousarray(b, np.float32).ravel()
        cc = None if c is None else np.ascontiguousarray(c, np.float32).ravel()
        d = np.empty(3, np.float32)
        self.lib.orbo_gemm3.argtypes = [C.c_void_p] * 4
        self.lib.orbo_gemm3(_p(A), _p(b), _p(cc), _p(d))
        return d

    def project(self, Rcw, tcw, K4, bounds, xyz):
        """ORBmatcher.cc:1376-1393 for n world points: (u, v, invz, valid)"""
        R = np.ascontiguousarray(Rcw, np.float32); t = np.ascontiguousarray(tcw, np.float32).ravel()
        k = np.ascontiguousarray(K4, np.float32); bd = np.ascontiguousarray(bounds, np.float32)
        p = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        n = len(p)
        u = np.empty(n, np.float32); v = np.empty(n, np.float32); iz = np.empty(n, np.float32); ok = np.empty(n, np.int32)
        self.lib.orbo_project.argtypes = [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 4
        self.lib.orbo_project(_p(R), _p(t), _p(k), _p(bd), _p(p), n, _p(u), _p(v), _p(iz), _p(ok))
        return u, v, iz, ok

    def distinctive(self, desc, start):
        """MapPoint::ComputeDistinctiveDescriptors for len(start)-1 map points (CSR runs of `desc`): (best, median)"""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        st = np.ascontiguousarray(start, np.int32)
        best = np.empty(len(st) - 1, np.int32); med = np.empty(len(st) - 1, np.int32)
        self.lib.orbo_distinctive(_p(d), _p(st), len(st) - 1, _p(best), _p(med))
        return best, med

    def stereo(self, keys_l, desc_l, keys_r, desc_r, levels_l, levels_r, scale, inv_scale, mb, mbf):
        """Frame::ComputeStereoMatches; levels_* = lists of padded (h+38, w+38) uint8 arrays (mvImagePyramid).
        Returns (mvuRight, mvDepth, sad, kept)."""
        kl = np.ascontiguousarray(keys_l, KP_DTYPE); kr = np.ascontiguousarray(keys_r, KP_DTYPE)
        dl = np.ascontiguousarray(desc_l, np.uint8); dr = np.ascontiguousarray(desc_r, np.uint8)
        nl = len(levels_l)
        la = [np.ascontiguousarray(a, np.uint8) for a in levels_l]
        ra = [np.ascontiguousarray(a, np.uint8) for a in levels_r]
        pl = (C.c_void_p * nl)(*[a.ctypes.data for a in la])
        pr = (C.c_void_p * nl)(*[a.ctypes.data for a in ra])
        cols = np.array([a.shape[1] - 38 for a in la], np.int32)
        rows = np.array([a.shape[0] - 38 for a in la], np.int32)
        sf = np.ascontiguousarray(scale, np.float32); isf = np.ascontiguousarray(inv_scale, np.float32)
        ur = np.empty(len(kl), np.float32); depth = np.empty(len(kl), np.float32); sad = np.empty(len(kl), np.int32)
        kept = self.lib.orbo_stereo(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), pl, pr, _p(cols), _p(rows), nl,
                                    _p(sf), _p(isf), C.c_float(mb), C.c_float(mbf), _p(ur), _p(depth), _p(sad))
        return ur, depth, sad, kept

    def bruteforce(self, q, qa, t, ta, ratio=0.9, check_ori=True):
        q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
        qa = np.ascontiguousarray(qa, np.float32); ta = np.ascontiguousarray(ta, np.float32)
        nq, nt = len(q), len(t)
        best = np.empty(nq, np.int32); second = np.empty(nq, np.int32)
        idx = np.empty(nq, np.int32); m12 = np.empty(nq, np.int32)
        n = self.lib.orbo_bruteforce(_p(q), _p(qa), nq, _p(t), _p(ta), nt, ratio, int(check_ori), _p(best), _p(second),
                                     _p(idx), _p(m12))
        return n, best, second, idx, m12

    def kf_pair(self, d1, a1, d2, a2, ratio=0.75, check_ori=True):
        d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
        a1 = np.ascontiguousarray(a1, np.float32); a2 = np.ascontiguousarray(a2, np.float32)
        m12 = np.empty(len(d1), np.int32)
        n = self.lib.orbo_kf_pair(_p(d1), _p(a1), len(d1), _p(d2), _p(a2), len(d2), ratio, int(check_ori), _p(m12))
        return n, m12


class OracleExtractor:
    def __init__(self, oracle, nfeatures, scale, nlevels, ini, mn):
        self.o = oracle
        self.lib = oracle.lib
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = C.c_void_p(self.lib.orbo_create(nfeatures, scale, nlevels, ini, mn))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orbo_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        f = [np.empty(n, np.float32) for _ in range(4)]
        per = np.empty(n, np.int32)
        umax = np.empty(16, np.int32)
        self.lib.orbo_tables(self.h, *[_p(a) for a in f], _p(per), _p(umax))
        return dict(scale=f[0], inv_scale=f[1], sigma2=f[2], inv_sigma2=f[3], per_level=per, umax=umax)

    def extract(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = self.nfeatures * 2 + 64 * self.nlevels
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.lib.orbo_extract(self.h, _p(img), w, h, img.strides[0], _p(kps), _p(desc), cap)
        if n < 0:
            raise ValueError("unsupported geometry")
        assert n <= cap
        return kps[:n].copy(), desc[:n].copy()

    def level_size(self, l):
        w = C.c_int(); h = C.c_int()
        self.lib.orbo_level_size(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_padded(self, l):
        w, h = self.level_size(l)
        out = np.empty((h + 38, w + 38), np.uint8)
        self.lib.orbo_level_padded(self.h, l, _p(out))
        return out

    def level_blurred(self, l):
        w, h = self.level_size(l)
        out = np.empty((h, w), np.uint8)
        ok = self.lib.orbo_level_blurred(self.h, l, _p(out))
        return out if ok else None

    def level_candidates(self, l):
        w, h = self.level_size(l)
        cap = max(16, (w * h) // 4)
        out = np.zeros(cap, KP_DTYPE)
        n = self.lib.orbo_level_candidates(self.h, l, _p(out), cap)
        return out[:n].copy()

    def level_selected(self, l):
        cap = self.nfeatures * 2 + 64
        out = np.zeros(cap, KP_DTYPE)
        n = self.lib.orbo_level_selected(self.h, l, _p(out), cap)
        return out[:n].copy()

    def stage_ms(self):
        out = np.empty(3, np.float64)
        self.lib.orbo_stage_ms(self.h, _p(out))
        return out


class OracleFrame:
    """Frame/KeyFrame arrays + 64x48 grid (Frame.cc:574-589)."""

    def __init__(self, oracle, keys_un, desc, bounds):
        self.o = oracle
        self.lib = oracle.lib
        self.keys = np.ascontiguousarray(keys_un, KP_DTYPE)
        self.desc = np.ascontiguousarray(desc, np.uint8)
        self.n = len(self.keys)
        self.bounds = tuple(float(b) for b in bounds)  # minX, minY, maxX, maxY
        self.h = C.c_void_p(self.lib.orbo_frame_create(_p(self.keys), _p(self.desc), self.n, *self.bounds))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orbo_frame_destroy(self.h)
            self.h = None

    def grid(self):
        start = np.empty(64 * 48 + 1, np.int32)
        idx = np.empty(max(self.n, 1), np.int32)
        self.lib.orbo_frame_grid(self.h, _p(start), _p(idx))
        return start, idx[:start[-1]].copy()

    def area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.empty(max(self.n, 1), np.int32)
        n = self.lib.orbo_frame_area(self.h, x, y, r, min_level, max_level, _p(out), len(out))
        return out[:n].copy()

    def search_init(self, other, prev_xy, window=100, ratio=0.9, check_ori=True):
        prev = np.ascontiguousarray(prev_xy, np.float32).copy()
        m12 = np.empty(self.n, np.int32)
        n = self.lib.orbo_search_init(self.h, other.h, _p(prev), _p(m12), window, ratio, int(check_ori))
        return n, m12, prev

    def search_projection(self, scale_factors, queries, qdesc, th, mode=0, occupied=None, u_right=None, mbf=0.0,
                          check_ori=True, max_distance=100):
        sf = np.ascontiguousarray(scale_factors, np.float32)
        q = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        occ = np.zeros(self.n, np.uint8) if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        match = np.empty(self.n, np.int32)
        n = self.lib.orbo_search_projection_ex(self.h, _p(sf), _p(ur), mbf, _p(q), _p(qd), len(q), th, mode, max_distance,
                                               _p(occ), _p(match), int(check_ori))
        return n, match

    def search_projection_world(self, scale_factors, Rcw, tcw, K4, queries, qdesc, th, mode=0, occupied=None, u_right=None,
                                mbf=0.0, check_ori=True, max_distance=100):
        """ORBmatcher.cc:1341-1498 from world points: project (:1376-1393, match_oracle.cpp::project_points; the image
        bounds test is the search's own), then the windowed search."""
        wq = np.ascontiguousarray(queries, WORLD_QUERY_DTYPE)
        xyz = np.stack([wq["x"], wq["y"], wq["z"]], axis=1)
        big = np.array([-np.inf, -np.inf, np.inf, np.inf], np.float32)
        u, v, iz, _ = self.o.project(Rcw, tcw, K4, big, xyz)
        q = np.zeros(len(wq), PROJ_QUERY_DTYPE)
        q["u"], q["v"], q["invz"] = u, v, iz
        q["octave"], q["obsPositive"], q["angle"] = wq["octave"], wq["obsPositive"], wq["angle"]
        q["valid"] = (wq["valid"] != 0) & ~(iz < 0)
        return self.search_projection(scale_factors, q, qdesc, th, mode, occupied, u_right, mbf, check_ori, max_distance)

    def search_points(self, scale_factors, queries, qdesc, th, ratio, occupied=None, u_right=None):
        sf = np.ascontiguousarray(scale_factors, np.float32)
        q = np.ascontiguousarray(queries, MP_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        occ = np.zeros(self.n, np.uint8) if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        match = np.empty(self.n, np.int32)
        n = self.lib.orbo_search_points(self.h, _p(sf), _p(ur), _p(q), _p(qd), len(q), th, ratio, _p(occ), _p(match))
        return n, match

    def search_triangulation(self, other, fv1, fv2, F12, ex, ey, sf2, sigma2_2, has1=None, has2=None, ur1=None,
                             ur2=None, only_stereo=False, check_ori=False):
        def fv(v):
            return [np.ascontiguousarray(a, np.int32) for a in v]
        n1, s1, i1 = fv(fv1)
        n2, s2, i2 = fv(fv2)
        has1 = np.zeros(self.n, np.uint8) if has1 is None else np.ascontiguousarray(has1, np.uint8)
        has2 = np.zeros(other.n, np.uint8) if has2 is None else np.ascontiguousarray(has2, np.uint8)
        ur1 = None if ur1 is None else np.ascontiguousarray(ur1, np.float32)
        ur2 = None if ur2 is None else np.ascontiguousarray(ur2, np.float32)
        F12 = np.ascontiguousarray(F12, np.float32)
        sf2 = np.ascontiguousarray(sf2, np.float32)
        sg2 = np.ascontiguousarray(sigma2_2, np.float32)
        m12 = np.empty(self.n, np.int32)
        n = self.lib.orbo_search_triangulation(self.h, other.h, len(n1), _p(n1), _p(s1), _p(i1), len(n2), _p(n2),
                                               _p(s2), _p(i2), _p(has1), _p(has2), _p(ur1), _p(ur2), _p(F12), ex, ey,
                                               _p(sf2), _p(sg2), int(only_stereo), int(check_ori), _p(m12))
        return n, m12

    def search_bow(self, other, fv1, fv2, valid1=None, valid2=None, ratio=0.7, check_ori=True, strict_low=False):
        """SearchByBoW (ORBmatcher.cc:159-288 strict_low=False, 522-655 strict_low=True) -> (n, matches12, matches21)"""
        def fv(v):
            return [np.ascontiguousarray(a, np.int32) for a in v]
        n1, s1, i1 = fv(fv1)
        n2, s2, i2 = fv(fv2)
        v1 = None if valid1 is None else np.ascontiguousarray(valid1, np.uint8)
        v2 = None if valid2 is None else np.ascontiguousarray(valid2, np.uint8)
        m12 = np.empty(self.n, np.int32)
        m21 = np.empty(other.n, np.int32)
        self.lib.orbo_search_bow.argtypes = ([C.c_void_p, C.c_void_p] + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] * 2
                                             + [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p])
        n = self.lib.orbo_search_bow(self.h, other.h, len(n1), _p(n1), _p(s1), _p(i1), len(n2), _p(n2), _p(s2), _p(i2),
                                     _p(v1), _p(v2), ratio, int(check_ori), int(strict_low), _p(m12), _p(m21))
        return n, m12, m21

    def search_best(self, queries, qdesc, chi2=False, u_right=None, inv_sigma2=None):
        """Independent projected search of Fuse / SearchBySim3 -> (best_idx, best_dist)"""
        q = np.ascontiguousarray(queries, BEST_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        s2 = None if inv_sigma2 is None else np.ascontiguousarray(inv_sigma2, np.float32)
        bi = np.empty(len(q), np.int32); bd = np.empty(len(q), np.int32)
        self.lib.orbo_search_best.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
        self.lib.orbo_search_best(self.h, _p(q), _p(qd), len(q), int(chi2), _p(ur), _p(s2), _p(bi), _p(bd))
        return bi, bd


BEST_QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("ur", "<f4"), ("level", "<i4"), ("valid", "<i4")])
