#!/usr/bin/env python3
"""Calls SearchForInitialization and SearchByProjection a few times (run under ncu to get per-kernel times)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-orb-slam-icra2018_b200"))
import orbb200
from orbb200.synth import shifted_pair
m = orbb200.Matcher(0)
sf = np.array([1.2 ** i for i in range(8)], np.float32)
for name, (w, h) in (("euroc", (752, 480)), ("kitti", (1241, 376))):
    a, b = shifted_pair(3, w, h)
    ex = orbb200.Extractor(2000, max_width=w, max_height=h)
    ka, da = ex(a); kb, db = ex(b)
    bounds = (0.0, 0.0, float(w), float(h))
    f1, f2 = m.frame(ka, da, bounds), m.frame(kb, db, bounds)
    prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    q = np.zeros(len(ka), orbb200.PROJ_QUERY_DTYPE)
    q["u"], q["v"], q["invz"], q["octave"], q["valid"], q["obs_positive"], q["angle"] = ka["x"] + 7, ka["y"] + 3, 0.1, ka["octave"], 1, 1, ka["angle"]
    for it in range(3):
        t0 = time.perf_counter(); m.search_for_initialization(f1, f2, prev, 100, 0.9, True); t1 = time.perf_counter()
        m.search_by_projection(f2, sf, q, da, 15.0, 0, None, None, 0.0, True); t2 = time.perf_counter()
        print(name, "init %.3f ms  proj %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
