#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel, meant to be run under compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-orb-slam-icra2018_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orbb200
from orbb200.synth import shifted_pair, synth_frame
from datagen import planted_descriptors

a, b = shifted_pair(3, 400, 300)
ex = orbb200.Extractor(600, max_width=400, max_height=300, max_batch=2)
(ka, da), (kb, db) = ex.extract_batch(np.stack([a, b]))
kn, dn = ex(synth_frame(5, 400, 300, noise_only=True))       # many candidates: quadtree spill path on level 0
# a batch of 9 takes the staged pyramid / blur kernels (bulk-async copies, mbarriers) and the wide level-0 copy (400 = 25 * 16)
ex9 = orbb200.Extractor(300, max_width=400, max_height=300, max_batch=9)
r9 = ex9.extract_batch(np.stack([synth_frame(20 + i, 400, 300) for i in range(9)]))
n9 = sum(len(k) for k, _ in r9)
ex9.close()
m = orbb200.Matcher(0)
bounds = (0.0, 0.0, 400.0, 300.0)
f1, f2 = m.frame(ka, da, bounds), m.frame(kb, db, bounds)
sf = np.array([1.2 ** i for i in range(8)], np.float32)
prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
n_init = m.search_for_initialization(f1, f2, prev, 100, 0.9, True)[0]
q = np.zeros(len(ka), orbb200.PROJ_QUERY_DTYPE)
q["u"], q["v"], q["invz"], q["octave"], q["valid"], q["obs_positive"], q["angle"] = ka["x"] + 7, ka["y"] + 3, 0.1, ka["octave"], 1, 1, ka["angle"]
n_proj = m.search_by_projection(f2, sf, q, da, 15.0)[0]
pq = np.zeros(len(ka), orbb200.POINT_QUERY_DTYPE)
pq["proj_x"], pq["proj_y"], pq["view_cos"], pq["level"], pq["in_view"], pq["obs_positive"] = ka["x"] + 7, ka["y"] + 3, 1.0, ka["octave"], 1, 1
n_pts = m.search_by_projection_points(f2, sf, pq, da, 3.0, 0.8)[0]
node = lambda k, dx, dy: ((k["y"] + dy) // 48).astype(np.int64) * 100 + ((k["x"] + dx) // 48).astype(np.int64)
def fv(k, dx, dy):
    nd = node(k, dx, dy); ids = np.unique(nd); start = [0]; idx = []
    for i in ids:
        idx.extend(np.nonzero(nd == i)[0].tolist()); start.append(len(idx))
    return ids.astype(np.int32), np.array(start, np.int32), np.array(idx, np.int32)
F12 = np.array([[0, 0, -3.0], [0, 0, 7.0], [3.0, -7.0, 0]], np.float32) * 1e-2
n_tri = m.search_for_triangulation(f1, f2, fv(ka, 0, 0), fv(kb, 7, 3), F12, 3000.0, 200.0, sf, sf * sf, check_ori=True)[0]
n_bow = m.search_by_bow(f1, f2, fv(ka, 0, 0), fv(kb, -7, -3), ratio=0.75, check_ori=True, strict_low=True)[0]
n_proj_kf = m.search_by_projection(f2, sf, q, da, 10.0, mode=3, check_ori=False, max_distance=50)[0] if "max_distance" in m.search_by_projection.__code__.co_varnames else -1
bq = np.zeros(len(ka), orbb200.BEST_QUERY_DTYPE)
bq["u"], bq["v"], bq["radius"], bq["ur"], bq["level"], bq["valid"] = ka["x"] + 7, ka["y"] + 3, 12.0, -1.0, ka["octave"], 1
best_idx, best_dist = m.search_projected_best(f2, bq, da, chi2=True, inv_sigma2=1.0 / (sf * sf))
# round-2 entry points: projection on the device, batched windowed searches
wq = np.zeros(len(ka), orbb200.WORLD_QUERY_DTYPE)
wq["x"], wq["y"], wq["z"] = (ka["x"] + 7 - 200.0) / 300.0 * 5.0, (ka["y"] + 3 - 150.0) / 300.0 * 5.0, 5.0
wq["octave"], wq["valid"], wq["obs_positive"], wq["angle"] = ka["octave"], 1, 1, ka["angle"]
n_world = m.search_by_projection_world(f2, sf, np.eye(3, dtype=np.float32), np.zeros(3, np.float32), (300.0, 300.0, 200.0, 150.0), wq, da, 15.0)[0]
jobs = [(f2, q, da, None, None), (f1, q[:100], da[:100], None, None), (f2, q[:0], da[:0], None, None)]
n_pbatch = sum(r[0] for r in m.search_by_projection_batch(jobs, sf, 15.0)[0])
n_ibatch = sum(r[0] for r in m.search_for_initialization_batch([(f1, f2, prev), (f2, f1, np.stack([kb["x"], kb["y"]], 1).astype(np.float32))], 100, 0.9, True)[0])
rng = np.random.default_rng(0)
qd, qa, td, ta = planted_descriptors(rng, 300, 270)
n_bf = int(m.bruteforce(qd, qa, td, ta, 0.9, True)["nmatches"])
d = m.distance(qd[:64], td[:64])
import torch
tab = torch.from_numpy(np.stack([qd, qd[::-1].copy(), planted_descriptors(rng, 300, 300)[0]])).cuda()
torch.manual_seed(0)
ang = torch.rand((3, 300), device="cuda") * 360
cnt = torch.zeros((3, 3), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
m.allpairs_device(tab, ang, 0, 3, 0, 3, 0.75, True, cnt)
m.synchronize()
print("ok", n9, n_world, n_pbatch, n_ibatch, len(ka), len(kn), n_init, n_proj, n_pts, n_tri, n_bow, n_proj_kf, int((best_idx >= 0).sum()), n_bf, int(d.sum()), cnt.cpu().numpy().tolist())
