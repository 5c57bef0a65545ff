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
