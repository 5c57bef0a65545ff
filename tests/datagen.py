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


from orbb200.synth import stereo_pair  # noqa: E402,F401  (lives with the other synthetic-input generators)
