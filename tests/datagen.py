"""Seeded inputs shared by the parity tests."""
import numpy as np


def planted_descriptors(rng, nq, nt, frac=0.6, max_flip=60):
    """Random 256-bit descriptors; a fraction of the train rows are queries with k ~ U{0..max_flip} flipped bits,
    so thresholds, the ratio test and ties are all exercised (config 4)."""
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    n = int(min(nq, nt) * frac)
    src = rng.permutation(nq)[:n]
    dst = rng.permutation(nt)[:n]
    for s, d in zip(src, dst):
        row = np.unpackbits(q[s])
        k = rng.integers(0, max_flip + 1)
        flip = rng.permutation(256)[:k]
        row[flip] ^= 1
        t[d] = np.packbits(row)
    qa = (rng.random(nq) * 360).astype(np.float32)
    ta = (rng.random(nt) * 360).astype(np.float32)
    # planted rows share a dominant rotation so the histogram keeps them
    ta[dst] = np.mod(qa[src] - 30.0 + rng.normal(0, 4, n), 360).astype(np.float32)
    return q, qa, t, ta


def stereo_pair(seed, w=752, h=480, disparities=(6, 14, 27, 41), sigma=2.0):
    """Rectified stereo pair of one synthetic scene: horizontal bands with different integer disparities (a point at
    uL in the left image sits at uL - d in the right one), fresh sensor noise per view."""
    from orbb200.synth import synth_frame
    rng = np.random.default_rng(seed + 77)
    dmax = max(disparities)
    big = synth_frame(seed, w + dmax, h).astype(np.float64)
    left = big[:, :w].copy()
    right = np.empty_like(left)
    edges = np.linspace(0, h, len(disparities) + 1).astype(int)
    for d, y0, y1 in zip(disparities, edges[:-1], edges[1:]):
        right[y0:y1] = big[y0:y1, d:d + w]
    left = np.clip(np.rint(left + rng.normal(0, sigma, left.shape)), 0, 255).astype(np.uint8)
    right = np.clip(np.rint(right + rng.normal(0, sigma, right.shape)), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(left), np.ascontiguousarray(right)
