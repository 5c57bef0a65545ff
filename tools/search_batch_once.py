"""One orbm_search_by_projection_batch call of 256 KITTI-shaped jobs (for ncu launch lists / host timing)."""
import os, sys, time
import numpy as np
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vi-orb-slam-icra2018_b200"))
import orbb200
from orbb200.synth import shifted_pair
w, h = 1241, 376
m = orbb200.Matcher(0)
ex = orbb200.Extractor(2000, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=2)
sf = np.array([1.2 ** i for i in range(8)], np.float32)
jobs = []
for k in range(16):
    a, b = shifted_pair(100 + k, w, h)
    (ka, da), (kb, db) = ex.extract_batch(np.stack([a, b]))
    f2 = m.frame(kb, db, (0.0, 0.0, float(w), float(h)))
    q = np.zeros(len(ka), orbb200.PROJ_QUERY_DTYPE)
    q["u"], q["v"], q["invz"], q["octave"], q["valid"], q["obs_positive"], q["angle"] = ka["x"] + 7, ka["y"] + 3, 0.1, ka["octave"], 1, 1, ka["angle"]
    jobs.append((f2, q, da, None, None))
jobs = jobs * 16
for it in range(3):
    t0 = time.perf_counter()
    res, cand = m.search_by_projection_batch(jobs, sf, 15.0)
    print("batch of %d: %.2f ms, %d candidates" % (len(jobs), (time.perf_counter() - t0) * 1e3, cand))
