#!/usr/bin/env python3
"""GPU checks that are CPU-pinned but were not run on the device in round 1 (its GPU budget was spent); DESIGN.md
section 8, item 7.  Kept OUT of tests/ until they have passed once on a B200 -- then move the asserts into
tests/test_frame_gpu.py / tests/test_stereo_gpu.py.  Run: gpurun -- 'python tools/carryover_checks.py'."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("vi-orb-slam-icra2018_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import orbb200  # noqa: E402
from oracle_py import Oracle  # noqa: E402
from stereo_cases import STEREO_CASES, images  # noqa: E402
from test_cvprims import EXTREME_CAMERAS  # noqa: E402

o = Oracle()
m = orbb200.Matcher(0)
fails = 0

# 1. undistortion guard branch (radial factor changes sign) through the CUDA path
for K4, dist, (w, h) in EXTREME_CAMERAS:
    rng = np.random.default_rng(3)
    pts = (rng.random((20000, 2)) * [w, h]).astype(np.float32)
    d = list(dist) + [0.0] * (5 - len(dist))
    got = m.undistort_points(orbb200.camera(*K4, *d), pts)
    ref = o.undistort(pts, K4, dist)
    bad = int((got.view(np.uint32) != ref.view(np.uint32)).any(1).sum())
    print("undistort extreme", K4, dist, "mismatching points:", bad)
    fails += bad > 0

# 2. stereo on keypoints the extractor would not produce (jitter, shifted octaves, shuffled order, other baselines)
for name in ("euroc_s1", "small_wide"):
    w, h, nfeat, mb, mbf = STEREO_CASES[name][:5]
    left, right = images(name)
    exl = orbb200.Extractor(nfeat, max_width=w, max_height=h)
    exr = orbb200.Extractor(nfeat, max_width=w, max_height=h)
    kl, dl = exl(left)
    kr, dr = exr(right)
    el, er = o.extractor(nfeat), o.extractor(nfeat)
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
        rur, rdepth, rsad, rkept = o.stereo(a, da, b, db, LL, RR, t["scale"], t["inv_scale"], mb, mbf2)
        ok = n == rkept and np.array_equal(ur.view(np.uint32), rur.view(np.uint32)) and \
            np.array_equal(depth.view(np.uint32), rdepth.view(np.uint32))
        print("stereo perturbed", name, seed, "kept", n, "ok" if ok else "MISMATCH")
        fails += not ok
    exl.close()
    exr.close()
m.close()
print("carry-over checks:", "all passed" if fails == 0 else "%d FAILED" % fails)
sys.exit(1 if fails else 0)
