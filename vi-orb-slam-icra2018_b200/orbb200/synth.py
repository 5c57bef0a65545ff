"""Seeded synthetic frames (no datasets are available offline) -- SURVEY.md section 8d.

Contrast-modulated blurred noise over a smooth base: about a third of the image is near-flat so that the
FAST 20->7 threshold fallback and empty cells both occur, and the quadtree enters its careful phase.
`synth_frame` (numpy/scipy, deterministic) is what tests and golden fixtures use; `synth_frames_torch`
generates big batches on the GPU for bench.py (same recipe, different RNG stream).
"""
import numpy as np


def _upsampled_grid(rng, w, h, pitch):
    from scipy import ndimage
    gw, gh = w // pitch + 3, h // pitch + 3
    g = rng.random((gh, gw))
    up = ndimage.zoom(g, pitch, order=3, mode="nearest")
    return up[pitch:pitch + h, pitch:pitch + w]


def synth_frame(seed, w=752, h=480, noise_only=False):
    """uint8 (h, w) frame. noise_only=True is the stress case (sigma 1.2 blurred noise, 16-20k L0 candidates)."""
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    u = rng.integers(0, 256, (h, w)).astype(np.float64)
    if noise_only:
        n = ndimage.gaussian_filter(u, 1.2, mode="reflect")
        n = (n - n.mean()) / n.std()
        return np.clip(np.rint(128 + n * 55), 0, 255).astype(np.uint8)
    n = ndimage.gaussian_filter(u, 1.6, mode="reflect")
    n = (n - n.mean()) / n.std()
    contrast = np.clip((_upsampled_grid(rng, w, h, 60) - 0.35) * 2.2, 0, 1) ** 2
    base = _upsampled_grid(rng, w, h, 120) * 120 + 60
    return np.clip(np.rint(base + n * contrast * 55), 0, 255).astype(np.uint8)


def shifted_pair(seed, w=752, h=480, dx=7, dy=3, sigma=2.0):
    """Two views of one scene: the second is the first shifted by (dx, dy) px with fresh sensor noise."""
    rng = np.random.default_rng(seed + 1_000_003)
    big = synth_frame(seed, w + dx, h + dy).astype(np.float64)
    a = big[dy:, dx:]
    b = big[:h, :w]
    a = np.clip(np.rint(a + rng.normal(0, sigma, a.shape)), 0, 255).astype(np.uint8)
    b = np.clip(np.rint(b + rng.normal(0, sigma, b.shape)), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def stereo_pair(seed, w=752, h=480, disparities=(6, 14, 27, 41), sigma=2.0):
    """Rectified stereo pair of one synthetic scene: horizontal bands with different integer disparities (a point at
    uL in the left image sits at uL - d in the right one), fresh sensor noise per view."""
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


def synth_frames_torch(n, w=752, h=480, seed=0, device="cuda"):
    """(n, h, w) uint8 tensor on `device`, same recipe as synth_frame evaluated with torch ops."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    u = torch.randint(0, 256, (n, 1, h, w), generator=g, device=device).float()
    r = 6
    x = torch.arange(-r, r + 1, device=device).float()
    k = torch.exp(-x * x / (2 * 1.6 * 1.6))
    k = k / k.sum()
    nz = F.conv2d(F.pad(u, (r, r, 0, 0), mode="reflect"), k.view(1, 1, 1, -1))
    nz = F.conv2d(F.pad(nz, (0, 0, r, r), mode="reflect"), k.view(1, 1, -1, 1))
    nz = (nz - nz.mean(dim=(2, 3), keepdim=True)) / nz.std(dim=(2, 3), keepdim=True)

    def grid(pitch):
        gw, gh = w // pitch + 3, h // pitch + 3
        t = torch.rand((n, 1, gh, gw), generator=g, device=device)
        up = F.interpolate(t, scale_factor=pitch, mode="bicubic", align_corners=False)
        return up[:, :, pitch:pitch + h, pitch:pitch + w]

    contrast = torch.clamp((grid(60) - 0.35) * 2.2, 0, 1) ** 2
    base = grid(120) * 120 + 60
    img = torch.clamp(torch.round(base + nz * contrast * 55), 0, 255).to(torch.uint8)
    return img[:, 0].contiguous()


# ---- synthetic DBoW2 vocabulary trees (the reference's Vocabulary/ORBvoc.bin, k = 10, L = 6, is not in the checkout;
# parity does not depend on how a tree was trained, only on its arrays)
def make_vocab(seed, k=10, L=4, ragged=False, stop_fraction=0.02):
    """Flat arrays of a DBoW2 tree built like HKmeansStep numbers it (children of a node get consecutive ids in creation
    order, depth first per level); child descriptors are noisy copies of the parent's so that descents are informative.
    ragged: some inner nodes keep fewer than k children and some branches end early (leaves above level L).
    stop_fraction of the words get weight 0 (DBoW2 'stopped' words: transform skips them)."""
    rng = np.random.default_rng(seed)
    desc = [np.zeros(32, np.uint8)]
    children = [[]]
    level = [0]
    frontier = [0]
    for lv in range(1, L + 1):
        nxt = []
        for parent in frontier:
            if ragged and lv > 1 and rng.random() < 0.08:
                continue                                # the branch ends here: `parent` stays a leaf
            nk = k if not ragged else int(rng.integers(2, k + 1))
            base = np.unpackbits(desc[parent]) if parent else rng.integers(0, 2, 256, dtype=np.uint8)
            for _ in range(nk):
                bits = base.copy() if parent else rng.integers(0, 2, 256, dtype=np.uint8)
                flip = rng.permutation(256)[:max(4, 96 >> lv)]
                bits[flip] ^= 1
                nid = len(desc)
                desc.append(np.packbits(bits))
                children.append([])
                level.append(lv)
                children[parent].append(nid)
                nxt.append(nid)
        frontier = nxt
    n = len(desc)
    word_id = np.zeros(n, np.int32)
    weight = np.zeros(n, np.float64)
    w = 0
    for i in range(n):
        if i and not children[i]:
            word_id[i] = w
            w += 1
            weight[i] = 0.0 if rng.random() < stop_fraction else float(np.log(rng.uniform(1.5, 400.0)))
    start = np.zeros(n + 1, np.int32)
    start[1:] = np.cumsum([len(c) for c in children])
    flat = np.array([c for cs in children for c in cs], np.int32)
    return dict(n_nodes=n, L=L, k=k, desc=np.ascontiguousarray(np.stack(desc)), child_start=start, children=flat,
                word_id=word_id, weight=weight, n_words=w)


def features_for(voc, seed, n):
    """descriptors near random tree nodes (so different branches are visited) plus pure noise"""
    rng = np.random.default_rng(seed)
    pick = rng.integers(1, voc["n_nodes"], n)
    bits = np.unpackbits(voc["desc"][pick], axis=1)
    for i in range(n):
        k = int(rng.integers(0, 50))
        bits[i, rng.permutation(256)[:k]] ^= 1
    out = np.packbits(bits, axis=1)
    out[::17] = rng.integers(0, 256, (len(out[::17]), 32), dtype=np.uint8)
    return np.ascontiguousarray(out)
