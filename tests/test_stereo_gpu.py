"""GPU parity of Frame::ComputeStereoMatches (Frame.cc:810-984) against the oracle
restatement: mvuRight and mvDepth bit for bit (float32 with the reference's operation order), the same matches kept."""
import os

import numpy as np
import pytest

import orbb200
from datagen import stereo_pair

pytestmark = pytest.mark.gpu

from stereo_cases import STEREO_CASES, images

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stereo_ref_vectors.npz")
RIGS = {"euroc": (752, 480, 1200, 47.90639384423901 / 435.2046959714599, 47.90639384423901)}


def _oracle_stereo(oracle, left, right, nfeat, mb, mbf):
    el, er = oracle.extractor(nfeat), oracle.extractor(nfeat)
    kl, dl = el.extract(left)
    kr, dr = er.extract(right)
    t = el.tables()
    LL = [el.level_padded(i) for i in range(8)]
    RR = [er.level_padded(i) for i in range(8)]
    ur, depth, sad, kept = oracle.stereo(kl, dl, kr, dr, LL, RR, t["scale"], t["inv_scale"], mb, mbf)
    return kl, dl, kr, dr, ur, depth, sad, kept


@pytest.mark.parametrize("name", sorted(STEREO_CASES))
def test_compute_stereo_matches(oracle, name):
    """extraction + stereo on the GPU == oracle == the reference's own Frame::ComputeStereoMatches (committed golden)"""
    w, h, nfeat, mb, mbf, seed, disparities, same_rows = STEREO_CASES[name]
    left, right = images(name)
    kl, dl, kr, dr, ur_ref, depth_ref, sad_ref, kept_ref = _oracle_stereo(oracle, left, right, nfeat, mb, mbf)
    exl = orbb200.Extractor(nfeat, max_width=w, max_height=h)
    exr = orbb200.Extractor(nfeat, max_width=w, max_height=h)
    gkl, gdl = exl(left)
    gkr, gdr = exr(right)
    assert gkl.tobytes() == kl.tobytes() and gkr.tobytes() == kr.tobytes()
    ur, depth, n = exl.stereo_matches(exr, gkl, gdl, gkr, gdr, mb, mbf)
    assert exl.launch_count() == 2
    assert n == kept_ref and kept_ref > 100
    assert np.array_equal(ur.view(np.uint32), ur_ref.view(np.uint32))
    assert np.array_equal(depth.view(np.uint32), depth_ref.view(np.uint32))
    assert int((ur >= 0).sum()) == n
    g = np.load(GOLDEN)
    assert n == int(g[name + "/kept"])
    assert np.array_equal(ur.view(np.uint32), g[name + "/u_right"].view(np.uint32))
    assert np.array_equal(depth.view(np.uint32), g[name + "/depth"].view(np.uint32))
    if name == "euroc_zero":
        assert (depth == np.float32(mbf) / np.float32(0.01)).any()     # the disparity <= 0 branch was taken
    else:   # the disparities are the planted ones (sub-pixel refined) wherever mbf/mb admits them
        ok = ur >= 0
        disp = kl["x"][ok] - ur[ok]
        band = np.minimum((kl["y"][ok] * 4 // h).astype(int), 3)
        for b, d in enumerate(disparities):
            if d < mbf / mb - 1 and (band == b).sum() > 10:
                assert abs(np.median(disp[band == b]) - d) < 0.6
    exl.close()
    exr.close()


def test_stereo_device_resident(oracle):
    """the same through device pointers: extraction outputs are consumed where the extractors wrote them"""
    torch = pytest.importorskip("torch")
    w, h, nfeat, mb, mbf = RIGS["euroc"]
    pairs = [stereo_pair(10 + i, w, h) for i in range(2)]
    exl = orbb200.Extractor(nfeat, max_width=w, max_height=h, max_batch=2)
    exr = orbb200.Extractor(nfeat, max_width=w, max_height=h, max_batch=2)
    cap = exl.capacity
    out = {}
    for name, ex, imgs in (("l", exl, [p[0] for p in pairs]), ("r", exr, [p[1] for p in pairs])):
        d_img = torch.from_numpy(np.stack(imgs)).cuda()
        d_k = torch.zeros((2, cap, 7), dtype=torch.int32, device="cuda")
        d_d = torch.zeros((2, cap, 32), dtype=torch.uint8, device="cuda")
        d_n = torch.zeros(2, dtype=torch.int32, device="cuda")
        ex.extract_batch_device(d_img, d_k, d_d, d_n)
        out[name] = (d_img, d_k, d_d, d_n)
    exl.synchronize()
    exr.synchronize()
    for f in range(2):
        d_u = torch.zeros(cap, dtype=torch.float32, device="cuda")
        d_z = torch.zeros(cap, dtype=torch.float32, device="cuda")
        d_s = torch.zeros(cap, dtype=torch.int32, device="cuda")
        d_kept = torch.zeros(1, dtype=torch.int32, device="cuda")
        exl.stereo_matches_device(exr, f, f, out["l"][1][f], out["l"][2][f], out["l"][3][f:f + 1], out["r"][1][f],
                                  out["r"][2][f], out["r"][3][f:f + 1], mb, mbf, d_u, d_z, d_s, d_kept)
        exl.synchronize()
        kl, dl, kr, dr, ur_ref, depth_ref, sad_ref, kept_ref = _oracle_stereo(oracle, pairs[f][0], pairs[f][1], nfeat, mb, mbf)
        n = int(out["l"][3][f])
        assert n == len(kl)
        assert int(d_kept) == kept_ref
        assert np.array_equal(d_u[:n].cpu().numpy().view(np.uint32), ur_ref.view(np.uint32))
        assert np.array_equal(d_z[:n].cpu().numpy().view(np.uint32), depth_ref.view(np.uint32))
        assert np.array_equal(d_s[:n].cpu().numpy(), sad_ref)
    exl.close()
    exr.close()


@pytest.mark.parametrize("name", ["euroc_s1", "small_wide"])
def test_stereo_on_perturbed_keypoints(oracle, name):
    """inputs the extractor would not produce (the oracle is pinned to the reference's text on exactly these by
    tests/test_stereo_ref.py): sub-pixel jitter, octaves shifted by +-1, shuffled order, other baselines"""
    w, h, nfeat, mb, mbf = STEREO_CASES[name][:5]
    left, right = images(name)
    exl = orbb200.Extractor(nfeat, max_width=w, max_height=h)
    exr = orbb200.Extractor(nfeat, max_width=w, max_height=h)
    kl, dl = exl(left)
    kr, dr = exr(right)
    el, er = oracle.extractor(nfeat), oracle.extractor(nfeat)
    el.extract(left)
    er.extract(right)
    t = el.tables()
    LL = [el.level_padded(i) for i in range(8)]
    RR = [er.level_padded(i) for i in range(8)]
    for seed in range(8):
        rng = np.random.default_rng(seed)
        a, b = kl.copy(), kr.copy()
        for k in (a, b):
            k["x"] += rng.uniform(-1.5, 1.5, len(k)).astype(np.float32)
            k["y"] += rng.uniform(-1.5, 1.5, len(k)).astype(np.float32)
        b["octave"] = np.clip(b["octave"] + rng.integers(-1, 2, len(b)), 0, 7)
        pa, pb = rng.permutation(len(a)), rng.permutation(len(b))
        a, da, b, db = a[pa], dl[pa], b[pb], dr[pb]
        mbf2 = float(mbf * rng.uniform(0.3, 2.0))
        ur, depth, n = exl.stereo_matches(exr, a, da, b, db, mb, mbf2)
        rur, rdepth, rsad, rkept = oracle.stereo(a, da, b, db, LL, RR, t["scale"], t["inv_scale"], mb, mbf2)
        assert n == rkept and n > 20
        assert np.array_equal(ur.view(np.uint32), rur.view(np.uint32))
        assert np.array_equal(depth.view(np.uint32), rdepth.view(np.uint32))
    exl.close()
    exr.close()


def test_stereo_pair_as_one_batch(oracle):
    """left and right image as ONE batch of two on ONE handle (one launch set instead of two): frames 0 and 1 of the same
    arena, same result as with two extractors"""
    name = "euroc_s2"
    w, h, nfeat, mb, mbf = STEREO_CASES[name][:5]
    left, right = images(name)
    ex = orbb200.Extractor(nfeat, max_width=w, max_height=h, max_batch=2)
    (kl, dl), (kr, dr) = ex.extract_batch(np.stack([left, right]))
    ur, depth, n = ex.stereo_matches(ex, kl, dl, kr, dr, mb, mbf, frame_l=0, frame_r=1)
    g = np.load(GOLDEN)
    assert n == int(g[name + "/kept"])
    assert np.array_equal(ur.view(np.uint32), g[name + "/u_right"].view(np.uint32))
    assert np.array_equal(depth.view(np.uint32), g[name + "/depth"].view(np.uint32))
    with pytest.raises(orbb200.OrbError):
        ex.stereo_matches(ex, kl, dl, kr, dr, mb, mbf, frame_l=0, frame_r=2)     # outside the last batch
    ex.close()


def test_stereo_edge_cases(oracle):
    w, h, nfeat, mb, mbf = RIGS["euroc"]
    left, right = stereo_pair(4, w, h)
    exl = orbb200.Extractor(nfeat, max_width=w, max_height=h)
    exr = orbb200.Extractor(nfeat, max_width=w, max_height=h)
    kl, dl = exl(left)
    kr, dr = exr(right)
    # no right keypoints, no left keypoints
    ur, depth, n = exl.stereo_matches(exr, kl, dl, kr[:0], dr[:0], mb, mbf)
    assert n == 0 and (ur == -1).all() and (depth == -1).all()
    ur, depth, n = exl.stereo_matches(exr, kl[:0], dl[:0], kr, dr, mb, mbf)
    assert n == 0 and len(ur) == 0
    # mismatched extractors are refused
    other = orbb200.Extractor(nfeat, max_width=640, max_height=480)
    other(left[:480, :640].copy())
    with pytest.raises(orbb200.OrbError):
        exl.stereo_matches(other, kl, dl, kr, dr, mb, mbf)
    for e in (exl, exr, other):
        e.close()
