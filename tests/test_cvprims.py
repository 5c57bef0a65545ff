"""The oracle's OpenCV primitive models against the installed OpenCV (cv2 4.13.0), bit for bit.
These models carry all the arithmetic the reference delegates to OpenCV, so this is what pins them."""
import numpy as np
import pytest

from orbb200.synth import synth_frame

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module", autouse=True)
def _single_thread():
    cv2.setNumThreads(1)


@pytest.mark.parametrize("w,h", [(752, 480), (1241, 376), (640, 480), (333, 251)])
def test_resize_chain(oracle, w, h):
    cur = synth_frame(1, w, h)
    for l in range(1, 8):
        dw, dh = int(round(w / 1.2 ** l)), int(round(h / 1.2 ** l))
        ref = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(oracle.resize(cur, dw, dh), ref), (w, h, l)
        cur = ref


def test_resize_adversarial(oracle):
    rng = np.random.default_rng(0)
    for _ in range(40):
        sw, sh = rng.integers(8, 200, 2)
        dw, dh = max(2, int(sw / rng.uniform(1.0, 1.5))), max(2, int(sh / rng.uniform(1.0, 1.5)))
        img = rng.choice([0, 1, 127, 128, 254, 255], (sh, sw)).astype(np.uint8)
        assert np.array_equal(oracle.resize(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


def test_border_and_blur(oracle):
    for seed, (w, h) in enumerate([(752, 480), (210, 134), (62, 62), (9, 8)]):
        img = synth_frame(seed, max(w, 64), max(h, 64))[:h, :w].copy()
        if min(w, h) > 19:
            assert np.array_equal(oracle.border(img), cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
        assert np.array_equal(oracle.blur(img), cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
    rng = np.random.default_rng(1)
    ext = rng.choice([0, 255], (90, 120)).astype(np.uint8)   # saturating pattern: rounding of the single >>16
    assert np.array_equal(oracle.blur(ext), cv2.GaussianBlur(ext, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))


def _cv_fast(img, th):
    det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kp = det.detect(img, None)
    return np.array([(k.pt[0], k.pt[1], k.response, k.size, k.angle, k.octave, k.class_id) for k in kp]).reshape(-1, 7)


def _oracle_fast(oracle, img, th):
    k = oracle.fast(img, th)
    return np.stack([k["x"], k["y"], k["response"], k["size"], k["angle"], k["octave"], k["class_id"]], 1).astype(np.float64)


@pytest.mark.parametrize("th", [20, 7, 1, 100])
def test_fast_full_image(oracle, th):
    for seed, noise in ((0, False), (3, True)):
        img = synth_frame(seed, 400, 300, noise_only=noise)
        a, b = _cv_fast(img, th), _oracle_fast(oracle, img, th)
        assert a.shape == b.shape and np.array_equal(a, b)   # same set, same order, same responses


def test_fast_cell_rois(oracle):
    """What the extractor really calls: FAST on ~36x36 views of a larger image, some thinner than 7 px."""
    img = synth_frame(0)
    rng = np.random.default_rng(0)
    for _ in range(300):
        x, y = rng.integers(0, 700), rng.integers(0, 440)
        w, h = rng.integers(7, 40), rng.integers(1, 40)
        roi = img[y:y + h, x:x + w]
        for th in (20, 7):
            a, b = _cv_fast(roi, th), _oracle_fast(oracle, roi, th)
            assert a.shape == b.shape and np.array_equal(a, b)


def test_fast_equal_neighbours_suppress_each_other(oracle):
    """Two adjacent pixels with equal maximal score: strict '>' drops both, so a cell can be empty at threshold 20 and
    trigger the 7 fallback (ORBextractor.cc:811-818)."""
    img = np.full((24, 24), 50, np.uint8)
    img[10:13, 10:14] = 200                      # a 4x3 bright blob: symmetric corners with equal scores
    for th in (20, 7):
        a, b = _cv_fast(img, th), _oracle_fast(oracle, img, th)
        assert a.shape == b.shape and np.array_equal(a, b)


def test_fast_atan2(oracle):
    rng = np.random.default_rng(0)
    y = rng.normal(0, 1000, 20000).astype(np.float32)
    x = rng.normal(0, 1000, 20000).astype(np.float32)
    y[:10] = 0; x[5:15] = 0; y[20:30] = x[20:30]; y[30:40] = -x[30:40]
    y[40:1000] = np.rint(y[40:1000]); x[40:1000] = np.rint(x[40:1000])     # integer moments, as IC_Angle produces
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    got = oracle.atan2(y, x)
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))
    assert oracle.atan2(np.zeros(1, np.float32), np.zeros(1, np.float32))[0] == 0.0


# Cameras of the reference's own settings files: Examples/Monocular/EuRoC.yaml:8-16, TUM1.yaml (5 coefficients), plus a
# strong barrel/pincushion pair so that the five iterations do not converge early.
CAMERAS = [
    ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05), (752, 480)),
    ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314), (640, 480)),
    ((300.0, 310.0, 320.0, 240.0), (-0.45, 0.21, 0.003, -0.002), (640, 480)),
    ((700.0, 705.0, 600.0, 180.0), (0.12, 0.05, -0.001, 0.0007, -0.01), (1241, 376)),
]


@pytest.mark.parametrize("K4,dist,size", CAMERAS)
def test_undistort_points(oracle, K4, dist, size):
    """cv::undistortPoints as Frame::UndistortKeyPoints / ComputeImageBounds call it (Frame.cc:767, :793)"""
    rng = np.random.default_rng(7)
    w, h = size
    pts = (rng.random((50000, 2)) * [w, h]).astype(np.float32)
    pts = np.concatenate([pts, np.array([[0, 0], [w, 0], [0, h], [w, h], [K4[2], K4[3]]], np.float32)])
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
    D = np.array(dist, np.float32).reshape(-1, 1)
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    got = oracle.undistort(pts, K4, dist)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


# Cameras far outside anything real: the radial denominator changes sign inside the image, so OpenCV's guard
# (icdist < 0 -> keep the normalised input, its regression test 14583) is taken for a large share of the points.
EXTREME_CAMERAS = [
    ((100.0, 100.0, 320.0, 240.0), (-5.0, 0.0, 0.0, 0.0), (640, 480)),
    ((150.0, 140.0, 320.0, 240.0), (-1.2, 0.3, 0.01, -0.02, 0.0), (640, 480)),
    ((120.0, 120.0, 300.0, 200.0), (2.5, -8.0, 0.0, 0.0, 3.0), (640, 480)),
]


@pytest.mark.parametrize("K4,dist,size", EXTREME_CAMERAS)
def test_undistort_points_sign_flip_guard(oracle, K4, dist, size):
    rng = np.random.default_rng(3)
    w, h = size
    pts = (rng.random((20000, 2)) * [w, h]).astype(np.float32)
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, np.array(dist, np.float32).reshape(-1, 1), None, K).reshape(-1, 2)
    got = oracle.undistort(pts, K4, dist)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.isclose(ref, pts, atol=1e-3).all(1).sum() > 1000          # the guard branch really was exercised


def test_gemm_3x3_projection(oracle):
    """x3Dc = Rcw * x3Dw + tcw (ORBmatcher.cc:1377): cv evaluates the expression as one gemm whose small-matrix path is
    float arithmetic in source order -- NOT a double accumulation.  Pinned on random poses, with and without the + tcw."""
    rng = np.random.default_rng(0)
    for i in range(4000):
        R = rng.normal(size=(3, 3)).astype(np.float32)
        P = (rng.normal(size=(3, 1)) * 5).astype(np.float32)
        t = rng.normal(size=(3, 1)).astype(np.float32)
        assert np.array_equal(oracle.gemm3(R, P, t).view(np.uint32), cv2.gemm(R, P, 1.0, t, 1.0).ravel().view(np.uint32))
        if i % 4 == 0:
            assert np.array_equal(oracle.gemm3(R, P).view(np.uint32), cv2.gemm(R, P, 1.0, None, 0.0).ravel().view(np.uint32))
    # the whole projection on a plausible pose: u, v, invz from float operations in source order
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = cv2.Rodrigues(np.array([0.02, -0.05, 0.01], np.float32))[0].astype(np.float32)
    T[:3, 3] = [0.1, -0.03, 0.2]
    K4 = np.array([458.654, 457.296, 367.215, 248.375], np.float32)
    xyz = (rng.normal(size=(2000, 3)) * [2, 1.5, 1] + [0, 0, 4]).astype(np.float32)
    u, v, iz, ok = oracle.project(T[:3, :3], T[:3, 3], K4, [0, 0, 752, 480], xyz)
    for i in range(len(xyz)):
        c = cv2.gemm(np.ascontiguousarray(T[:3, :3]), xyz[i].reshape(3, 1), 1.0, np.ascontiguousarray(T[:3, 3]).reshape(3, 1), 1.0).ravel()
        invz = np.float32(1.0 / np.float64(c[2]))
        uu = np.float32(np.float32(np.float32(K4[0] * c[0]) * invz) + K4[2])
        vv = np.float32(np.float32(np.float32(K4[1] * c[1]) * invz) + K4[3])
        assert u[i] == uu and v[i] == vv and iz[i] == invz
        assert bool(ok[i]) == bool(invz >= 0 and 0 <= uu <= 752 and 0 <= vv <= 480)
    assert 200 < ok.sum() < 2000
