#!/usr/bin/env python3
"""One invocation of every SURVEY section 8f entry point (device-resident frame, stereo, distinctive descriptors, BoW transform) on the
bench inputs, for `ncu` launch lists (tools/gpu_frame_side_prof.sh).  Not a benchmark: numbers printed under ncu are
never bench values."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-orb-slam-icra2018_b200"))
import torch  # noqa: E402
import orbb200  # noqa: E402
from orbb200.synth import stereo_pair  # noqa: E402

w, h = 752, 480
left, right = stereo_pair(1, w, h)
mb, mbf = 47.90639384423901 / 435.2046959714599, 47.90639384423901
exl = orbb200.Extractor(1200, max_width=w, max_height=h)
exr = orbb200.Extractor(1200, max_width=w, max_height=h)
m = orbb200.Matcher(0)
kl, dl = exl(left)
kr, dr = exr(right)
ur, depth, kept = exl.stereo_matches(exr, kl, dl, kr, dr, mb, mbf)
cam = orbb200.camera(458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)
cap = exl.capacity
d_img = torch.from_numpy(left[None]).cuda()
d_k = torch.zeros((1, cap, 7), dtype=torch.int32, device="cuda")
d_d = torch.zeros((1, cap, 32), dtype=torch.uint8, device="cuda")
d_n = torch.zeros(1, dtype=torch.int32, device="cuda")
exl.extract_batch_device(d_img, d_k, d_d, d_n)
exl.synchronize()
f = orbb200.Frame.from_device(m, d_k[0], d_d[0], d_n, cap, m.image_bounds(cam, w, h), cam)
rng = np.random.default_rng(3)
sizes = rng.integers(2, 41, 300 if "--small" in sys.argv else 20000)
start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
desc = rng.integers(0, 256, (int(start[-1]), 32), dtype=np.uint8)
best, med = m.distinctive_descriptors(desc, start)
from orbb200.synth import make_vocab  # noqa: E402
voc = make_vocab(seed=1, k=10, L=4)
v = m.vocabulary(voc)
bow = m.bow_transform(v, dl, 4)
m.vocabulary_destroy(v)
print("bow words", len(bow["bow_word"]), "nodes", len(bow["fv_node"]))
print("stereo kept", kept, "frame n", f.n, "distinctive", len(best))
