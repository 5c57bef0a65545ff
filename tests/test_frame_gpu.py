"""GPU parity of the Frame constructor's post-extraction work: cv::undistortPoints as
Frame::UndistortKeyPoints / ComputeImageBounds call it (Frame.cc:748-808) and the device-resident hand-over from the
extractor to the matcher (no host round trip of keypoints or descriptors).  Undistorted coordinates are bit-exact
against the oracle model (itself pinned to cv2 4.13 by tests/test_cvprims.py) and against cv2 directly."""
import numpy as np
import pytest

import orbb200
from orbb200.synth import shifted_pair
from test_cvprims import CAMERAS, EXTREME_CAMERAS

pytestmark = pytest.mark.gpu


def _cam(K4, dist):
    d = list(dist) + [0.0] * (5 - len(dist))
    return orbb200.camera(*K4, *d)


@pytest.mark.parametrize("K4,dist,size", CAMERAS)
def test_undistort_points(matcher, oracle, K4, dist, size):
    rng = np.random.default_rng(11)
    w, h = size
    pts = (rng.random((20000, 2)) * [w, h]).astype(np.float32)
    got = matcher.undistort_points(_cam(K4, dist), pts)
    ref = oracle.undistort(pts, K4, dist)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    cv2 = pytest.importorskip("cv2")
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
    cref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, np.array(dist, np.float32).reshape(-1, 1), None, K).reshape(-1, 2)
    assert np.array_equal(got.view(np.uint32), cref.view(np.uint32))


@pytest.mark.parametrize("K4,dist,size", EXTREME_CAMERAS)
def test_undistort_points_sign_flip_guard(matcher, oracle, K4, dist, size):
    """cameras whose radial factor changes sign inside the image: OpenCV's icdist < 0 guard is taken"""
    rng = np.random.default_rng(3)
    w, h = size
    pts = (rng.random((20000, 2)) * [w, h]).astype(np.float32)
    got = matcher.undistort_points(_cam(K4, dist), pts)
    ref = oracle.undistort(pts, K4, dist)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_undistort_is_a_copy_without_k1(matcher):
    pts = np.array([[1.5, 2.25], [700.0, 400.0]], np.float32)
    assert np.array_equal(matcher.undistort_points(None, pts), pts)
    # Frame.cc:750 tests only mDistCoef(0): k2/p1/p2 alone do not undistort
    assert np.array_equal(matcher.undistort_points(orbb200.camera(450, 450, 376, 240, 0.0, 0.1, 0.01, 0.01), pts), pts)
    assert matcher.undistort_points(None, np.zeros((0, 2), np.float32)).shape == (0, 2)


@pytest.mark.parametrize("K4,dist,size", CAMERAS)
def test_image_bounds(matcher, oracle, K4, dist, size):
    w, h = size
    c = oracle.undistort(np.array([[0, 0], [w, 0], [0, h], [w, h]], np.float32), K4, dist)
    ref = np.array([min(c[0, 0], c[2, 0]), min(c[0, 1], c[1, 1]), max(c[1, 0], c[3, 0]), max(c[2, 1], c[3, 1])], np.float32)
    assert np.array_equal(matcher.image_bounds(_cam(K4, dist), w, h), ref)
    assert np.array_equal(matcher.image_bounds(None, w, h), np.array([0, 0, w, h], np.float32))


@pytest.mark.parametrize("distorted", [False, True])
def test_device_resident_frame(matcher, oracle, distorted):
    """extract on the device -> orbm_frame_create_device -> SearchForInitialization, against the same chain through the
    host (oracle extractor, oracle undistortion, oracle grid and search)."""
    torch = pytest.importorskip("torch")
    K4, dist, (w, h) = CAMERAS[0]
    cam = _cam(K4, dist) if distorted else None
    a, b = shifted_pair(5, w, h)
    ex = orbb200.Extractor(2000, max_width=w, max_height=h, max_batch=2)
    cap = ex.capacity
    d_img = torch.from_numpy(np.stack([a, b])).cuda()
    d_kps = torch.zeros((2, cap, 7), dtype=torch.int32, device="cuda")
    d_desc = torch.zeros((2, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(2, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d_img, d_kps, d_desc, d_n)
    ex.synchronize()
    bounds = matcher.image_bounds(cam, w, h)
    frames = [orbb200.Frame.from_device(matcher, d_kps[i], d_desc[i], d_n[i:i + 1], cap, bounds, cam) for i in range(2)]
    assert matcher.launch_count() == 4     # import + scan + fill + sort, no copies of keypoints through the host

    oe = oracle.extractor(2000)
    oframes, okeys = [], []
    for img, f in zip((a, b), frames):
        rk, rd = oe.extract(img)
        ru = rk.copy()
        if distorted:
            xy = oracle.undistort(np.stack([rk["x"], rk["y"]], 1), K4, dist)
            ru["x"], ru["y"] = xy[:, 0], xy[:, 1]
        k, d = f.download()
        assert f.n == len(rk)
        assert k.tobytes() == ru.tobytes()          # mvKeysUn: every field, bit for bit
        assert np.array_equal(d, rd)
        of = oracle.frame(ru, rd, tuple(bounds))
        s0, i0 = f.grid()
        s1, i1 = of.grid()
        assert np.array_equal(s0, s1) and np.array_equal(i0, i1)
        oframes.append(of)
        okeys.append(ru)
    prev = np.stack([okeys[0]["x"], okeys[0]["y"]], 1).copy()
    n_gpu, m_gpu, p_gpu = matcher.search_for_initialization(frames[0], frames[1], prev.copy(), 100, 0.9, True)
    n_ref, m_ref, p_ref = oframes[0].search_init(oframes[1], prev.copy(), 100, 0.9, True)
    assert n_gpu == n_ref and n_ref > 20
    assert np.array_equal(m_gpu, m_ref) and np.array_equal(p_gpu, p_ref)
    for f in frames:
        f.close()
    ex.close()


def test_host_extract_then_device_frame(matcher, oracle):
    """The live loop without a second upload: orbx_extract (host keypoints out, as ORBextractor::operator() must) leaves its
    outputs on the device; orbx_last_device_outputs + orbm_frame_create_device build the Frame from them, and the search on
    two such frames equals the host chain.  A device call or a stale index is refused."""
    K4, dist, (w, h) = CAMERAS[0]
    cam = _cam(K4, dist)
    a, b = shifted_pair(6, w, h)
    bounds = matcher.image_bounds(cam, w, h)
    ex = orbb200.Extractor(1000, max_width=w, max_height=h, max_batch=1)
    frames, host = [], []
    for img in (a, b):
        k, d = ex(img)
        dk, dd, dc, cap, st = ex.last_device_outputs(0)
        frames.append(orbb200.Frame.from_device(matcher, dk, dd, dc, cap, bounds, cam, st))
        host.append((k, d))
    with pytest.raises(orbb200.OrbError):
        ex.last_device_outputs(1)
    oframes = []
    for (k, d), f in zip(host, frames):
        ru = k.copy()
        xy = oracle.undistort(np.stack([k["x"], k["y"]], 1), K4, dist)
        ru["x"], ru["y"] = xy[:, 0], xy[:, 1]
        fk, fd = f.download()
        assert fk.tobytes() == ru.tobytes() and np.array_equal(fd, d)
        oframes.append(oracle.frame(ru, d, tuple(bounds)))
    prev = np.stack([oframes[0].keys["x"], oframes[0].keys["y"]], 1).copy()
    n_gpu, m_gpu, p_gpu = matcher.search_for_initialization(frames[0], frames[1], prev.copy(), 100, 0.9, True)
    n_ref, m_ref, p_ref = oframes[0].search_init(oframes[1], prev.copy(), 100, 0.9, True)
    assert n_gpu == n_ref and n_ref > 20 and np.array_equal(m_gpu, m_ref) and np.array_equal(p_gpu, p_ref)
    for f in frames:
        f.close()
    ex.close()


def test_device_frame_rejects_bad_arguments(matcher):
    torch = pytest.importorskip("torch")
    d_k = torch.zeros((4, 7), dtype=torch.int32, device="cuda")
    d_d = torch.zeros((4, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.tensor([9], dtype=torch.int32, device="cuda")
    with pytest.raises(orbb200.OrbError) as e:
        orbb200.Frame.from_device(matcher, d_k, d_d, d_n, 4, (0, 0, 752, 480))
    assert e.value.status == 2      # ORB_ERR_CAPACITY: count on the device exceeds the capacity
    with pytest.raises(orbb200.OrbError):
        orbb200.Frame.from_device(matcher, d_k, d_d, d_n, 4, (0, 0, 0, 480))
    d_n0 = torch.zeros(1, dtype=torch.int32, device="cuda")
    f = orbb200.Frame.from_device(matcher, d_k, d_d, d_n0, 4, (0, 0, 752, 480))   # a frame without keypoints is legal
    assert f.n == 0 and f.grid()[0][-1] == 0
    f.close()
