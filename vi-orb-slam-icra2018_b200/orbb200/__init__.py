"""ctypes harness over liborbb200.so (include/orbb200.h).

Python is only the test/bench harness here: the product is the C-ABI library built from csrc/*.cu for sm_100a, and the
C++ adapter in adapter/.  There is no CPU fallback: if the library is missing, or a call fails, this raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_ROOT, "liborbb200.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
PROJ_QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("invz", "<f4"), ("octave", "<i4"), ("valid", "<i4"),
                             ("obs_positive", "<i4"), ("angle", "<f4")])
WORLD_QUERY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("octave", "<i4"), ("valid", "<i4"),
                              ("obs_positive", "<i4"), ("angle", "<f4")])
POSE_DTYPE = np.dtype([("Rcw", "<f4", (9,)), ("tcw", "<f4", (3,)), ("fx", "<f4"), ("fy", "<f4"), ("cx", "<f4"), ("cy", "<f4")])
POINT_QUERY_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"),
                              ("level", "<i4"), ("in_view", "<i4"), ("obs_positive", "<i4")])

CAMERA_DTYPE = np.dtype([(k, "<f4") for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3")])

ORB_OK = 0


def camera(fx, fy, cx, cy, k1=0.0, k2=0.0, p1=0.0, p2=0.0, k3=0.0):
    """orb_camera: mK and mDistCoef of a Frame"""
    c = np.zeros(1, CAMERA_DTYPE)
    c[0] = (fx, fy, cx, cy, k1, k2, p1, p2, k3)
    return c


class OrbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("orbb200 status %d: %s" % (status, message))
        self.status = status


_lib = None


def lib():
    """The loaded C-ABI library. Raises if it has not been built (python vi-orb-slam-icra2018_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OrbError(-1, "liborbb200.so is not built: run `python vi-orb-slam-icra2018_b200/build.py` "
                               "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _abi.declare(_lib)
    return _lib


def _check(st):
    if st != ORB_OK:
        raise OrbError(st, lib().orb_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _dp(t):
    """device pointer of a torch tensor (or an int / None)"""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


def device_count():
    return lib().orb_device_count()


class Extractor:
    """ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)  -- ORBextractor.h:69-79"""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th_fast=20, min_th_fast=7,
                 max_width=752, max_height=480, max_batch=1, device=0):
        self.L = lib()
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.max_batch = max_batch
        h = C.c_void_p()
        _check(self.L.orbx_create(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, max_width, max_height,
                                  max_batch, device, C.byref(h)))
        self.h = h
        cap = C.c_int()
        _check(self.L.orbx_keypoint_capacity(self.h, C.byref(cap)))
        self.capacity = cap.value

    def close(self):
        if getattr(self, "h", None):
            self.L.orbx_destroy(self.h)
            self.h = None

    __del__ = close

    def __call__(self, image):
        """operator()(image, mask, keypoints, descriptors): returns (keypoints[KP_DTYPE], descriptors[n,32])."""
        image = np.asarray(image)
        if image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2 and image.strides[1] == 1
        h, w = image.shape
        kps = np.zeros(self.capacity, KP_DTYPE)
        desc = np.zeros((self.capacity, 32), np.uint8)
        n = C.c_int()
        _check(self.L.orbx_extract(self.h, _p(image), w, h, image.strides[0], _p(kps), _p(desc), self.capacity,
                                   C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images):
        """images: (F, H, W) uint8 host array -> list of (keypoints, descriptors)."""
        images = np.ascontiguousarray(images, np.uint8)
        f, h, w = images.shape
        kps = np.zeros((f, self.capacity), KP_DTYPE)
        desc = np.zeros((f, self.capacity, 32), np.uint8)
        n = np.zeros(f, np.int32)
        _check(self.L.orbx_extract_batch(self.h, _p(images), f, w, h, w, C.c_size_t(h * w), _p(kps), _p(desc),
                                         self.capacity, _p(n)))
        return [(kps[i, :n[i]].copy(), desc[i, :n[i]].copy()) for i in range(f)]

    def extract_batch_device(self, d_images, d_kps, d_desc, d_n, stream=None):
        """torch tensors on the handle's device: images (F,H,W) uint8, kps (F,cap,7) int32/float32 words,
        desc (F,cap,32) uint8, n (F,) int32. Enqueues only."""
        f, h, w = d_images.shape
        _check(self.L.orbx_extract_batch_device(self.h, _dp(d_images), f, w, h, d_images.stride(1),
                                                C.c_size_t(d_images.stride(0)), _dp(d_kps), _dp(d_desc),
                                                self.capacity, _dp(d_n), _dp(stream)))

    def synchronize(self):
        _check(self.L.orbx_synchronize(self.h))

    def stereo_matches(self, right, keys_l, desc_l, keys_r, desc_r, mb, mbf, frame_l=0, frame_r=0):
        """Frame::ComputeStereoMatches on the pyramids of the last extract calls of self (left) and `right`:
        returns (mvuRight, mvDepth, number of matches)."""
        kl = np.ascontiguousarray(keys_l, KP_DTYPE); kr = np.ascontiguousarray(keys_r, KP_DTYPE)
        dl = np.ascontiguousarray(desc_l, np.uint8); dr = np.ascontiguousarray(desc_r, np.uint8)
        ur = np.empty(len(kl), np.float32)
        depth = np.empty(len(kl), np.float32)
        n = C.c_int()
        _check(self.L.orbx_compute_stereo_matches(self.h, frame_l, right.h, frame_r, _p(kl), _p(dl), len(kl), _p(kr), _p(dr),
                                                  len(kr), C.c_float(mb), C.c_float(mbf), _p(ur), _p(depth), C.byref(n)))
        return ur, depth, n.value

    def stereo_matches_device(self, right, frame_l, frame_r, d_keys_l, d_desc_l, d_n_l, d_keys_r, d_desc_r, d_n_r, mb, mbf,
                              d_u_right, d_depth, d_sad, d_kept, stream=None):
        """device-resident ComputeStereoMatches over one (left, right) frame of two extract_batch_device calls"""
        _check(self.L.orbx_compute_stereo_matches_device(self.h, frame_l, right.h, frame_r, _dp(d_keys_l), _dp(d_desc_l),
                                                         _dp(d_n_l), self.capacity, _dp(d_keys_r), _dp(d_desc_r),
                                                         _dp(d_n_r), right.capacity, C.c_float(mb), C.c_float(mbf),
                                                         _dp(d_u_right), _dp(d_depth), _dp(d_sad), _dp(d_kept), _dp(stream)))

    def tables(self):
        n = self.nlevels
        f = [np.empty(n, np.float32) for _ in range(4)]
        per = np.empty(n, np.int32)
        umax = np.empty(16, np.int32)
        _check(self.L.orbx_get_scale_tables(self.h, *[_p(a) for a in f], _p(per), _p(umax)))
        return dict(scale=f[0], inv_scale=f[1], sigma2=f[2], inv_sigma2=f[3], per_level=per, umax=umax)

    def level(self, level, frame=0):
        w, h = C.c_int(), C.c_int()
        _check(self.L.orbx_get_level(self.h, frame, level, None, C.byref(w), C.byref(h)))
        out = np.empty((h.value + 38, w.value + 38), np.uint8)
        _check(self.L.orbx_get_level(self.h, frame, level, _p(out), C.byref(w), C.byref(h)))
        return out

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        _check(self.L.orbx_get_level(self.h, 0, level, None, C.byref(w), C.byref(h)))
        return w.value, h.value

    def blurred(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        _check(self.L.orbx_debug_blurred(self.h, frame, level, _p(out)))
        return out

    def candidates(self, level, frame=0):
        w, h = self.level_size(level)
        cap = max(16, (w * h) // 4)
        out = np.zeros(cap, KP_DTYPE)
        n = C.c_int()
        _check(self.L.orbx_debug_candidates(self.h, frame, level, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def stage_times(self):
        out = np.empty(3, np.float64)
        _check(self.L.orbx_stage_times(self.h, _p(out)))
        return out

    def launch_count(self):
        n = C.c_int()
        _check(self.L.orbx_last_launch_count(self.h, C.byref(n)))
        return n.value

    def last_device_outputs(self, frame=0):
        """Device pointers (ints) of what the last HOST extract call returned for `frame`:
        (d_keys, d_descriptors, d_count, capacity, stream)"""
        k, d, c, st = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        cap = C.c_int()
        _check(self.L.orbx_last_device_outputs(self.h, frame, C.byref(k), C.byref(d), C.byref(c), C.byref(cap), C.byref(st)))
        return k.value, d.value, c.value, cap.value, st.value

    def set_profiling(self, on):
        _check(self.L.orbx_set_profiling(self.h, int(on)))

    def kernel_times(self):
        """(ms summed per kernel group [pyramid, fast, quadtree, blur, brief], number of calls) since last query"""
        ms = np.zeros(5, np.float64)
        n = C.c_int()
        _check(self.L.orbx_kernel_times(self.h, _p(ms), C.byref(n)))
        return ms, n.value


class Frame:
    """Device-resident Frame/KeyFrame arrays + the 64x48 grid (Frame.cc:574-589)."""

    def __init__(self, matcher, keys_un, desc, bounds):
        self.m = matcher
        self.L = matcher.L
        keys = np.ascontiguousarray(keys_un, KP_DTYPE)
        d = np.ascontiguousarray(desc, np.uint8)
        self.n = len(keys)
        self.keys, self.desc = keys, d
        f = C.c_void_p()
        _check(self.L.orbm_frame_create(matcher.h, _p(keys), _p(d), self.n, *[C.c_float(b) for b in bounds], C.byref(f)))
        self.f = f

    @classmethod
    def from_device(cls, matcher, d_keys, d_desc, d_count, capacity, bounds, cam=None, stream=None):
        """Frame post-extraction on the device (orbm_frame_create_device): d_* are device pointers (ints) or torch
        tensors of ONE frame as orbx_extract_batch_device wrote them."""
        self = cls.__new__(cls)
        self.m, self.L = matcher, matcher.L
        f = C.c_void_p()
        _check(self.L.orbm_frame_create_device(matcher.h, _dp(d_keys), _dp(d_desc), _dp(d_count), capacity, _p(cam),
                                               *[C.c_float(b) for b in bounds], _dp(stream), C.byref(f)))
        self.f = f
        n = C.c_int()
        _check(self.L.orbm_frame_size(self.f, C.byref(n)))
        self.n = n.value
        self.keys = self.desc = None
        return self

    def download(self):
        """(mvKeysUn, mDescriptors) of the device-resident frame"""
        keys = np.zeros(self.n, KP_DTYPE)
        desc = np.zeros((self.n, 32), np.uint8)
        _check(self.L.orbm_frame_download(self.f, _p(keys), _p(desc)))
        return keys, desc

    def close(self):
        if getattr(self, "f", None):
            self.L.orbm_frame_destroy(self.f)
            self.f = None

    __del__ = close

    def grid(self):
        start = np.empty(64 * 48 + 1, np.int32)
        idx = np.empty(max(self.n, 1), np.int32)
        _check(self.L.orbm_frame_grid(self.f, _p(start), _p(idx)))
        return start, idx[:start[-1]].copy()

    def area(self, xyr, min_level=-1, max_level=-1, cap=None):
        xyr = np.ascontiguousarray(xyr, np.float32).reshape(-1, 3)
        cap = cap or max(self.n, 1)
        idx = np.empty((len(xyr), cap), np.int32)
        cnt = np.empty(len(xyr), np.int32)
        _check(self.L.orbm_features_in_area(self.f, _p(xyr), len(xyr), min_level, max_level, _p(idx), cap, _p(cnt)))
        return [idx[i, :min(cnt[i], cap)].copy() for i in range(len(xyr))]


class Matcher:
    """ORBmatcher search loops on arrays -- ORBmatcher.h:41-83"""

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        _check(self.L.orbm_create(device, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.orbm_destroy(self.h)
            self.h = None

    __del__ = close

    def synchronize(self):
        _check(self.L.orbm_synchronize(self.h))

    def launch_count(self):
        n = C.c_int()
        _check(self.L.orbm_last_launch_count(self.h, C.byref(n)))
        return n.value

    def frame(self, keys_un, desc, bounds):
        return Frame(self, keys_un, desc, bounds)

    def undistort_points(self, cam, xy):
        """cv::undistortPoints(pts, pts, mK, mDistCoef, Mat(), mK)  (Frame.cc:767)"""
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        out = np.empty_like(xy)
        _check(self.L.orbm_undistort_points(self.h, _p(cam), _p(xy), len(xy), _p(out)))
        return out

    def image_bounds(self, cam, width, height):
        """Frame::ComputeImageBounds: (mnMinX, mnMinY, mnMaxX, mnMaxY)"""
        b = np.empty(4, np.float32)
        _check(self.L.orbm_image_bounds(self.h, _p(cam), width, height, _p(b)))
        return b

    def distance(self, a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        out = np.empty(len(a), np.int32)
        _check(self.L.orbm_distance(self.h, _p(a), _p(b), len(a), _p(out)))
        return out

    def bruteforce(self, q, qa, t, ta, ratio=0.9, check_ori=True):
        """q: (P,nq,32) or (nq,32); t likewise. Returns dict of arrays shaped (P,nq) (+ nmatches (P,))."""
        q = np.ascontiguousarray(q, np.uint8)
        t = np.ascontiguousarray(t, np.uint8)
        single = q.ndim == 2
        if single:
            q, t = q[None], t[None]
        p, nq, nt = q.shape[0], q.shape[1], t.shape[1]
        qa = np.ascontiguousarray(qa, np.float32).reshape(p, nq)
        ta = np.ascontiguousarray(ta, np.float32).reshape(p, nt)
        best = np.empty((p, nq), np.int32); second = np.empty((p, nq), np.int32)
        idx = np.empty((p, nq), np.int32); m12 = np.empty((p, nq), np.int32)
        nm = np.empty(p, np.int32)
        _check(self.L.orbm_bruteforce(self.h, _p(q), _p(qa), nq, _p(t), _p(ta), nt, p, C.c_float(ratio), int(check_ori),
                                      _p(best), _p(second), _p(idx), _p(m12), _p(nm)))
        r = dict(best=best, second=second, idx=idx, matches12=m12, nmatches=nm)
        return {k: v[0] for k, v in r.items()} if single else r

    def bruteforce_device(self, dq, dqa, dt, dta, ratio, check_ori, d_best, d_second, d_idx, d_m12, d_n, stream=None):
        p, nq, nt = dq.shape[0], dq.shape[1], dt.shape[1]
        _check(self.L.orbm_bruteforce_device(self.h, _dp(dq), _dp(dqa), nq, _dp(dt), _dp(dta), nt, p, C.c_float(ratio),
                                             int(check_ori), _dp(d_best), _dp(d_second), _dp(d_idx), _dp(d_m12),
                                             _dp(d_n), _dp(stream)))

    def allpairs_device(self, d_table, d_angles, q_begin, q_end, db_begin, db_end, ratio, check_ori, d_counts,
                        stream=None):
        n_kf, n_desc = d_table.shape[0], d_table.shape[1]
        _check(self.L.orbm_allpairs_device(self.h, _dp(d_table), _dp(d_angles), n_kf, n_desc, q_begin, q_end, db_begin,
                                           db_end, C.c_float(ratio), int(check_ori), _dp(d_counts), _dp(stream)))

    def allpairs_sharded(self, comm, d_local_desc, d_local_angles, kf_per_rank, ratio, check_ori, d_counts, chunk_kf=128,
                         q_count=-1, stream=None):
        """Config 5 through the C ABI: this rank's (n_local x n_kf) tile; chunked ncclAllGather overlapped with matching."""
        n_desc = d_local_desc.shape[1]
        per = np.ascontiguousarray(kf_per_rank, np.int32)
        _check(self.L.orbm_allpairs_sharded(self.h, comm.h, _dp(d_local_desc), _dp(d_local_angles), _p(per), n_desc, q_count, chunk_kf,
                                            C.c_float(ratio), int(check_ori), _dp(d_counts), C.c_void_p(stream or 0)))

    def distinctive_descriptors(self, desc, start):
        """MapPoint::ComputeDistinctiveDescriptors for len(start)-1 map points: (best index inside each run, median)"""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        st = np.ascontiguousarray(start, np.int32)
        best = np.empty(len(st) - 1, np.int32)
        med = np.empty(len(st) - 1, np.int32)
        _check(self.L.orbm_distinctive_descriptors(self.h, _p(d), _p(st), len(st) - 1, _p(best), _p(med)))
        return best, med

    def vocabulary(self, voc):
        """upload a flat DBoW2 tree (dict with n_nodes, L, desc, child_start, children, word_id, weight) -> handle"""
        v = C.c_void_p()
        keep = [np.ascontiguousarray(voc["desc"], np.uint8), np.ascontiguousarray(voc["child_start"], np.int32),
                np.ascontiguousarray(voc["children"], np.int32), np.ascontiguousarray(voc["word_id"], np.int32),
                np.ascontiguousarray(voc["weight"], np.float64)]
        _check(self.L.orbm_vocabulary_create(self.h, int(voc["n_nodes"]), int(voc["L"]), *[_p(a) for a in keep], C.byref(v)))
        return v

    def vocabulary_destroy(self, v):
        _check(self.L.orbm_vocabulary_destroy(v))

    def bow_transform(self, v, desc, levelsup=4):
        """TemplatedVocabulary::transform(features, BowVector, FeatureVector, levelsup): dict like the oracle's"""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        pw = np.empty(n, np.int32); pwt = np.empty(n, np.float64); pn = np.empty(n, np.int32)
        bw = np.empty(n + 1, np.int32); bv = np.empty(n + 1, np.float64)
        fn = np.empty(n + 1, np.int32); fs = np.empty(n + 2, np.int32); fi = np.empty(n + 1, np.int32)
        nw, nn = C.c_int(), C.c_int()
        _check(self.L.orbm_bow_transform(self.h, v, _p(d), n, levelsup, _p(pw), _p(pwt), _p(pn), _p(bw), _p(bv), C.byref(nw),
                                         _p(fn), _p(fs), _p(fi), C.byref(nn)))
        return dict(word=pw, weight=pwt, node=pn, bow_word=bw[:nw.value].copy(), bow_value=bv[:nw.value].copy(),
                    fv_node=fn[:nn.value].copy(), fv_start=fs[:nn.value + 1].copy(), fv_idx=fi[:fs[nn.value]].copy())

    def bow_transform_device(self, v, d_desc, n, levelsup, d_word, d_weight, d_node, stream=None):
        _check(self.L.orbm_bow_transform_device(self.h, v, _dp(d_desc), n, levelsup, _dp(d_word), _dp(d_weight), _dp(d_node),
                                                _dp(stream)))

    def popc_peak(self):
        v = C.c_double()
        _check(self.L.orbm_popc_peak(self.h, C.byref(v)))
        return v.value

    def search_for_initialization(self, f1, f2, prev_xy, window=100, ratio=0.9, check_ori=True):
        prev = np.ascontiguousarray(prev_xy, np.float32).copy()
        m12 = np.empty(f1.n, np.int32)
        n = C.c_int()
        _check(self.L.orbm_search_for_initialization(self.h, f1.f, f2.f, _p(prev), _p(m12), window, C.c_float(ratio),
                                                     int(check_ori), C.byref(n)))
        return n.value, m12, prev

    def search_for_initialization_batch(self, pairs, window=100, ratio=0.9, check_ori=True):
        """pairs: [(f1, f2, prev_xy)]; returns ([(n, matches12, prev_xy)], candidate compares) -- orbm_search_for_initialization_batch"""
        jobs = (_abi.InitJob * len(pairs))()
        keep = []
        for j, (f1, f2, prev_xy) in enumerate(pairs):
            prev = np.ascontiguousarray(prev_xy, np.float32).copy()
            m12 = np.empty(f1.n, np.int32)
            keep.append((prev, m12))
            jobs[j].f1, jobs[j].f2 = f1.f, f2.f
            jobs[j].prev_xy, jobs[j].matches12 = prev.ctypes.data, m12.ctypes.data
        cand = C.c_longlong()
        _check(self.L.orbm_search_for_initialization_batch(self.h, jobs, len(pairs), window, C.c_float(ratio), int(check_ori),
                                                           C.byref(cand)))
        return [(jobs[j].nmatches, keep[j][1], keep[j][0]) for j in range(len(pairs))], cand.value

    def search_by_projection_batch(self, jobs_in, scale_factors, th, mode=0, mbf=0.0, check_ori=True, max_distance=100):
        """jobs_in: [(cur, queries, qdesc, occupied or None, u_right or None)]; returns ([(n, cur_match)], candidate
        compares) -- orbm_search_by_projection_batch"""
        sf = np.ascontiguousarray(scale_factors, np.float32)
        jobs = (_abi.ProjectionJob * len(jobs_in))()
        keep = []
        for j, (cur, queries, qdesc, occupied, u_right) in enumerate(jobs_in):
            q = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE)
            qd = np.ascontiguousarray(qdesc, np.uint8)
            occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
            ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
            match = np.empty(cur.n, np.int32)
            keep.append((q, qd, occ, ur, match))
            jobs[j].cur, jobs[j].queries, jobs[j].query_desc, jobs[j].nq = cur.f, q.ctypes.data, qd.ctypes.data, len(q)
            jobs[j].u_right = None if ur is None else ur.ctypes.data
            jobs[j].occupied = None if occ is None else occ.ctypes.data
            jobs[j].cur_match = match.ctypes.data
        cand = C.c_longlong()
        _check(self.L.orbm_search_by_projection_batch(self.h, jobs, len(jobs_in), _p(sf), len(sf), C.c_float(mbf), C.c_float(th), mode,
                                                      max_distance, int(check_ori), C.byref(cand)))
        return [(jobs[j].nmatches, keep[j][4]) for j in range(len(jobs_in))], cand.value

    def search_by_projection(self, cur, scale_factors, queries, qdesc, th, mode=0, occupied=None, u_right=None,
                             mbf=0.0, check_ori=True, max_distance=100):
        sf = np.ascontiguousarray(scale_factors, np.float32)
        q = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        occ = np.zeros(cur.n, np.uint8) if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        match = np.empty(cur.n, np.int32)
        n = C.c_int()
        _check(self.L.orbm_search_by_projection_ex(self.h, cur.f, _p(sf), len(sf), _p(ur), C.c_float(mbf), _p(q), _p(qd),
                                                   len(q), C.c_float(th), mode, max_distance, _p(occ), _p(match),
                                                   int(check_ori), C.byref(n)))
        return n.value, match

    def project_points(self, Rcw, tcw, K4, xyz):
        """ORBmatcher.cc:1376-1388 on the device: (u, v, invz) of n world points"""
        pose = np.zeros(1, POSE_DTYPE)
        pose["Rcw"][0] = np.asarray(Rcw, np.float32).ravel()
        pose["tcw"][0] = np.asarray(tcw, np.float32).ravel()
        pose["fx"], pose["fy"], pose["cx"], pose["cy"] = [np.float32(v) for v in K4]
        p = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        u = np.empty(len(p), np.float32); v = np.empty(len(p), np.float32); iz = np.empty(len(p), np.float32)
        _check(self.L.orbm_project_points(self.h, _p(pose), _p(p), len(p), _p(u), _p(v), _p(iz)))
        return u, v, iz

    def search_by_projection_world(self, cur, scale_factors, Rcw, tcw, K4, queries, qdesc, th, mode=0, occupied=None,
                                   u_right=None, mbf=0.0, check_ori=True, max_distance=100):
        """SearchByProjection(Current, Last, th, bMono) with the projection (ORBmatcher.cc:1376-1393) on the device:
        queries = world points (WORLD_QUERY_DTYPE), Rcw / tcw = the current pose, K4 = (fx, fy, cx, cy)."""
        sf = np.ascontiguousarray(scale_factors, np.float32)
        q = np.ascontiguousarray(queries, WORLD_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        pose = np.zeros(1, POSE_DTYPE)
        pose["Rcw"][0] = np.asarray(Rcw, np.float32).ravel()
        pose["tcw"][0] = np.asarray(tcw, np.float32).ravel()
        pose["fx"], pose["fy"], pose["cx"], pose["cy"] = [np.float32(v) for v in K4]
        occ = np.zeros(cur.n, np.uint8) if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        match = np.empty(cur.n, np.int32)
        n = C.c_int()
        _check(self.L.orbm_search_by_projection_world(self.h, cur.f, _p(sf), len(sf), _p(ur), C.c_float(mbf), _p(pose), _p(q),
                                                      _p(qd), len(q), C.c_float(th), mode, max_distance, _p(occ), _p(match),
                                                      int(check_ori), C.byref(n)))
        return n.value, match

    def search_by_projection_points(self, f, scale_factors, queries, qdesc, th, ratio, occupied=None, u_right=None):
        sf = np.ascontiguousarray(scale_factors, np.float32)
        q = np.ascontiguousarray(queries, POINT_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        occ = np.zeros(f.n, np.uint8) if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        match = np.empty(f.n, np.int32)
        n = C.c_int()
        _check(self.L.orbm_search_by_projection_points(self.h, f.f, _p(sf), len(sf), _p(ur), _p(q), _p(qd), len(q),
                                                       C.c_float(th), C.c_float(ratio), _p(occ), _p(match),
                                                       C.byref(n)))
        return n.value, match

    def search_for_triangulation(self, k1, k2, fv1, fv2, F12, ex, ey, sf2, sigma2_2, has1=None, has2=None, ur1=None,
                                 ur2=None, only_stereo=False, check_ori=False):
        def fv(v):
            return [np.ascontiguousarray(a, np.int32) for a in v]
        n1, s1, i1 = fv(fv1)
        n2, s2, i2 = fv(fv2)
        has1 = np.zeros(k1.n, np.uint8) if has1 is None else np.ascontiguousarray(has1, np.uint8)
        has2 = np.zeros(k2.n, np.uint8) if has2 is None else np.ascontiguousarray(has2, np.uint8)
        ur1 = None if ur1 is None else np.ascontiguousarray(ur1, np.float32)
        ur2 = None if ur2 is None else np.ascontiguousarray(ur2, np.float32)
        F12 = np.ascontiguousarray(F12, np.float32)
        sf2 = np.ascontiguousarray(sf2, np.float32)
        sg2 = np.ascontiguousarray(sigma2_2, np.float32)
        m12 = np.empty(k1.n, np.int32)
        n = C.c_int()
        _check(self.L.orbm_search_for_triangulation(self.h, k1.f, k2.f, len(n1), _p(n1), _p(s1), _p(i1), len(n2),
                                                    _p(n2), _p(s2), _p(i2), _p(has1), _p(has2), _p(ur1), _p(ur2),
                                                    _p(F12), C.c_float(ex), C.c_float(ey), _p(sf2), _p(sg2), len(sf2),
                                                    int(only_stereo), int(check_ori), _p(m12), C.byref(n)))
        return n.value, m12

    def search_by_bow(self, k1, k2, fv1, fv2, valid1=None, valid2=None, ratio=0.7, check_ori=True, strict_low=False):
        """SearchByBoW (ORBmatcher.cc:159-288 strict_low=False; 522-655 strict_low=True) -> (n, matches12, matches21)"""
        def fv(v):
            return [np.ascontiguousarray(a, np.int32) for a in v]
        n1, s1, i1 = fv(fv1)
        n2, s2, i2 = fv(fv2)
        v1 = None if valid1 is None else np.ascontiguousarray(valid1, np.uint8)
        v2 = None if valid2 is None else np.ascontiguousarray(valid2, np.uint8)
        m12 = np.empty(k1.n, np.int32)
        m21 = np.empty(k2.n, np.int32)
        n = C.c_int()
        _check(self.L.orbm_search_by_bow(self.h, k1.f, k2.f, len(n1), _p(n1), _p(s1), _p(i1), len(n2), _p(n2), _p(s2),
                                         _p(i2), _p(v1), _p(v2), C.c_float(ratio), int(check_ori), int(strict_low),
                                         _p(m12), _p(m21), C.byref(n)))
        return n.value, m12, m21

    def search_projected_best(self, kf, queries, qdesc, chi2=False, u_right=None, inv_sigma2=None):
        """Projected search of Fuse / SearchBySim3 (ORBmatcher.cc:892-944, 1051-1075, 1191-1215) -> (best_idx, best_dist)"""
        q = np.ascontiguousarray(queries, BEST_QUERY_DTYPE)
        qd = np.ascontiguousarray(qdesc, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        s2 = None if inv_sigma2 is None else np.ascontiguousarray(inv_sigma2, np.float32)
        bi = np.empty(len(q), np.int32)
        bd = np.empty(len(q), np.int32)
        _check(self.L.orbm_search_projected_best(self.h, kf.f, _p(q), _p(qd), len(q), int(chi2), _p(ur), _p(s2),
                                                 0 if s2 is None else len(s2), _p(bi), _p(bd)))
        return bi, bd


BEST_QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("ur", "<f4"), ("level", "<i4"), ("valid", "<i4")])


class Comm:
    """NCCL communicator of the sharded all-pairs path (orbm_comm_*): one per process / GPU."""

    @staticmethod
    def unique_id():
        buf = np.zeros(128, np.uint8)
        _check(lib().orbm_comm_unique_id(_p(buf)))
        return buf

    def __init__(self, unique_id, rank, world, device):
        self.L = lib()
        self.h = C.c_void_p()
        uid = np.ascontiguousarray(unique_id, np.uint8)
        _check(self.L.orbm_comm_create(_p(uid), rank, world, device, C.byref(self.h)))
        self.rank, self.world = rank, world

    def close(self):
        if self.h:
            self.L.orbm_comm_destroy(self.h)
            self.h = C.c_void_p()

    def nccl_version(self):
        v = C.c_int()
        _check(self.L.orbm_comm_info(self.h, None, None, C.byref(v)))
        return v.value

    def last_gather(self):
        ms, by, ch = C.c_double(), C.c_double(), C.c_int()
        _check(self.L.orbm_comm_last_gather(self.h, C.byref(ms), C.byref(by), C.byref(ch)))
        return {"ms": ms.value, "bytes_received": by.value, "chunks": ch.value}
