"""The oracle against the committed golden vectors (outputs of the reference's own ORBextractor.cc built in place, see
tools/gen_golden.py), against oracle/_ref when it is present, and the known-answer tests SURVEY.md section 8c lists."""
import hashlib
import os

import numpy as np
import pytest

from orbb200.synth import synth_frame

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_image(g):
    if g["image"].size:
        return g["image"]
    img = synth_frame(int(g["seed"]), int(g["w"]), int(g["h"]), noise_only=bool(g["noise"]))
    if sha(img) != str(g["image_sha256"]):
        pytest.skip("synthetic generator differs on this host (numpy/scipy version), golden input not reproducible")
    return img


@pytest.mark.parametrize("name", ["euroc_s0", "small_s5", "qvga_noise", "kitti_s2"])
def test_extract_golden(oracle, name):
    g = np.load(os.path.join(GOLD, "extract_%s.npz" % name))
    img = golden_image(g)
    ex = oracle.extractor(int(g["nfeatures"]))
    kps, desc = ex.extract(img)
    assert kps.tobytes() == g["kps"].tobytes()
    assert np.array_equal(desc, g["desc"])
    for l in range(8):
        assert sha(ex.level_padded(l)) == str(g["level_sha256"][l])


def test_octree_golden_kats(oracle):
    g = np.load(os.path.join(GOLD, "octree_kats.npz"))
    win = g["win"]
    names = sorted(k[:-3] for k in g.files if k.endswith("_in"))
    assert len(names) >= 8
    for n in names:
        out = oracle.distribute(g[n + "_in"], int(win[0]), int(win[1]), int(win[2]), int(win[3]), int(g[n + "_n"]))
        assert out.tobytes() == g[n + "_out"].tobytes(), n


def test_against_reference_binary(oracle):
    """Where oracle/_ref exists (built from /root/reference in the build container and shipped with the snapshot), run
    the reference itself on fresh inputs."""
    ref_runner = pytest.importorskip("ref_runner")
    if not ref_runner.ref_binary("orb_ref"):
        pytest.skip("oracle/_ref not built")
    frames = [synth_frame(40 + i, 480, 360) for i in range(3)]
    try:
        res = ref_runner.ref_extract(frames, nfeatures=700, dump_levels=True)
    except Exception as e:  # pragma: no cover
        pytest.skip("oracle/_ref does not run here: %s" % e)
    ex = oracle.extractor(700)
    for img, r in zip(frames, res):
        kps, desc = ex.extract(img)
        assert kps.tobytes() == r["kps"].tobytes() and np.array_equal(desc, r["desc"])
        assert all(np.array_equal(ex.level_padded(l), r["levels"][l]) for l in range(8))


def test_constructor_tables(oracle):
    t = oracle.extractor(1000).tables()
    assert list(t["per_level"]) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(oracle.extractor(2000).tables()["per_level"]) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert list(t["umax"]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    want = np.array([1, 1.2000000477, 1.4400000572, 1.7280001640, 2.0736002922, 2.4883203507, 2.9859845638, 3.5831816196])
    assert np.allclose(t["scale"], want, rtol=0, atol=1e-9 * 4)
    assert np.array_equal(t["inv_scale"], np.float32(1) / t["scale"])
    sizes = [oracle.extractor(1000).level_size(0)]
    ex = oracle.extractor(1000)
    ex.extract(synth_frame(0))
    assert [ex.level_size(l) for l in range(8)] == [(752, 480), (627, 400), (522, 333), (435, 278), (363, 231), (302, 193),
                                                   (252, 161), (210, 134)]


def test_pattern_table(oracle):
    p = oracle.pattern()
    assert hashlib.sha256(p.astype("<i4").tobytes()).hexdigest() == \
        "7e645581387b82784797e8adddb9b6f0c12611859fda09ca8a9bec96d767a05f"
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = open(os.path.join(here, "oracle", "brief_pattern.inc")).read()
    b = open(os.path.join(here, "vi-orb-slam-icra2018_b200", "csrc", "brief_pattern.inc")).read()
    assert a == b
    assert p.min() == -13 and p.max() == 13 or p.max() == 12


def test_hamming_kats(oracle):
    z, o = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    assert oracle.distance(z, o) == 256 and oracle.distance(z, z) == 0
    rng = np.random.default_rng(0)
    for _ in range(200):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        want = sum(int(x).bit_count() for x in (a ^ b))
        assert oracle.distance(a, b) == want
        c = a.copy(); c[rng.integers(32)] ^= 1 << rng.integers(8)
        assert oracle.distance(a, c) == 1


def test_rotation_histogram_kats(oracle):
    # factor = 1/30 (upstream quirk): bins are 30 degrees wide, only 0..12 are reachable
    assert oracle.rotation_bin(15.0, 0.0) == 1          # round(0.5) away from zero
    assert oracle.rotation_bin(14.9, 0.0) == 0
    assert oracle.rotation_bin(345.0, 0.0) == 12
    assert oracle.rotation_bin(0.0, 15.0) == 12         # negative wrap: -15 + 360
    assert oracle.rotation_bin(359.9, 0.0) == 12
    assert max(oracle.rotation_bin(a, 0.0) for a in np.arange(0, 360, 0.25)) == 12
    sizes = np.zeros(30, np.int32); sizes[[2, 5, 7, 9]] = [100, 9, 50, 10]
    assert oracle.three_maxima(sizes) == (2, 7, 9)
    sizes[9] = 9                                         # third below 10% of the first (and 5 ties it, first wins)
    assert oracle.three_maxima(sizes) == (2, 7, -1)
    sizes[7] = 9
    assert oracle.three_maxima(sizes) == (2, -1, -1)
    assert oracle.three_maxima(np.zeros(30, np.int32)) == (-1, -1, -1)


def test_grid_kats(oracle):
    from oracle_py import KP_DTYPE
    k = np.zeros(6, KP_DTYPE)
    # 752x480 bounds: cell width 11.75, height 10
    k["x"] = [0.0, 5.874, 5.876, 751.9, 746.0, 100.0]
    k["y"] = [0.0, 0.0, 0.0, 100.0, 100.0, 475.1]
    f = oracle.frame(k, np.zeros((6, 32), np.uint8), (0.0, 0.0, 752.0, 480.0))
    start, idx = f.grid()
    cell = {int(i): int(np.searchsorted(start, p, side="right") - 1) for p, i in enumerate(idx)}
    assert cell[0] == 0 and cell[1] == 0 and cell[2] == 1 * 48        # round() boundary between column 0 and 1
    assert 3 not in cell                                              # posX == 64 -> dropped (Frame.cc:730-733)
    assert cell[4] == 63 * 48 + 10
    assert 5 not in cell                                              # posY == 48 -> dropped
    assert list(f.area(3.0, 1.0, 4.0)) == [0, 1, 2]
    assert list(f.area(3.0, 1.0, 3.0)) == [1, 2]                      # |dx| < r is strict: key 0 is exactly 3 px away
    assert list(f.area(3.0, 1.0, 2.875)) == [1]
    assert list(f.area(-500.0, 0.0, 10.0)) == []


def test_search_init_steals_matches(oracle):
    """Two level-0 queries want the same train keypoint; the later, closer one steals it (ORBmatcher.cc:463-467)."""
    from oracle_py import KP_DTYPE
    rng = np.random.default_rng(2)
    d = rng.integers(0, 256, (3, 32), dtype=np.uint8)
    q = np.stack([d[0], d[0]]).copy()
    q[0, 0] ^= 0x0f                                   # distance 4 to the target
    q[1, 0] ^= 0x01                                   # distance 1
    k1 = np.zeros(2, KP_DTYPE); k1["x"] = [50, 52]; k1["y"] = [50, 50]
    k2 = np.zeros(3, KP_DTYPE); k2["x"] = [51, 300, 400]; k2["y"] = [50, 300, 300]
    f1 = oracle.frame(k1, q, (0, 0, 752, 480)); f2 = oracle.frame(k2, d, (0, 0, 752, 480))
    n, m12, prev = f1.search_init(f2, np.stack([k1["x"], k1["y"]], 1), 100, 0.9, False)
    assert n == 1 and list(m12) == [-1, 0]
    assert list(prev[1]) == [51.0, 50.0] and list(prev[0]) == [50.0, 50.0]
