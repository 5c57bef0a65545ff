"""Arithmetic identities the round-2 kernels rest on, restated with numpy and checked on the CPU (no GPU, no library call).

Each case mirrors a rewrite that replaced the reference-shaped arithmetic inside a kernel by a cheaper form; the GPU parity tests
prove the kernels, these pin the identities themselves so that an edit of either side shows up without a device:

  * csrc/octree.cu::child_of         -- quadrant of a key from the packed box words (ORBextractor.cc:483-539, DivideNode)
  * csrc/pyramid.cu::py_vsum         -- vertical pass of cv::resize INTER_LINEAR with 32-bit products (was: two multiply-high)
  * csrc/blur.cu::blur_hrow          -- 7-tap row sums of four adjacent pixels as 8 four-way dot products
  * csrc/brief.cu (brief_staged)     -- cvRound by the 1.5 * 2^23 bias instead of a float -> int conversion
  * csrc/fast_warp.cu ticket_cells   -- runs of 8 cells + single cells at the end hand out every cell exactly once
"""
import numpy as np


def test_quadtree_child_of_on_packed_words():
    rng = np.random.default_rng(1)
    n = 400_000
    x = rng.integers(0, 2048, n)
    z = np.minimum(x + rng.integers(0, 2048, n), 4095)
    y = rng.integers(0, 2048, n)
    w = np.minimum(y + rng.integers(0, 2048, n), 4095)
    mx = x + ((z - x + 1) >> 1)            # the reference's midpoints: UL.x + ceil((UR.x - UL.x) / 2)
    my = y + ((w - y + 1) >> 1)
    near = rng.random(n) < 0.6             # most keys sit within two pixels of a midpoint
    kx = np.where(near, np.clip(mx + rng.integers(-2, 3, n), 0, 4095), rng.integers(0, 4096, n))
    ky = np.where(near, np.clip(my + rng.integers(-2, 3, n), 0, 4095), rng.integers(0, 4096, n))
    key = (kx.astype(np.uint64) << 20) | (ky.astype(np.uint64) << 8) | rng.integers(0, 256, n).astype(np.uint64)
    ref = (kx >= mx).astype(int) + 2 * (ky >= my).astype(int)
    w0 = (x | (y << 16)).astype(np.uint64)
    w1 = (z | (w << 16)).astype(np.uint64)
    s = (w0 + w1 + 0x00010001) & 0xFFFFFFFF
    tx = ((s & 0xFFFE) << 19) & 0xFFFFFFFF
    ty = (s >> 9) & 0xFFFFFF00
    got = (key >= tx).astype(int) + 2 * ((key & 0xFFF00) >= ty).astype(int)
    assert np.array_equal(ref, got)


def test_pyramid_vertical_pass_with_32_bit_products():
    rng = np.random.default_rng(2)
    n = 500_000
    # horizontal sums T = p0 * a0 + p1 * a1 with a0 + a1 = 2048 (INTER_RESIZE_COEF_SCALE), p <= 255
    a0 = rng.integers(0, 2049, n)
    T0 = rng.integers(0, 256, n) * a0 + rng.integers(0, 256, n) * (2048 - a0)
    a1 = rng.integers(0, 2049, n)
    T1 = rng.integers(0, 256, n) * a1 + rng.integers(0, 256, n) * (2048 - a1)
    b0 = rng.integers(0, 2049, n)
    b1 = 2048 - b0
    # OpenCV's VResizeLinear for 8-bit: ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
    ref = (((b0 * (T0 >> 4)) >> 16) + ((b1 * (T1 >> 4)) >> 16) + 2) >> 2
    # round-1/2 form: hi32((T & ~15) * (b << 12))
    hi = lambda t, b: ((t & ~15).astype(np.uint64) * (b.astype(np.uint64) << 12)) >> 32
    assert np.array_equal(ref, (hi(T0, b0) + hi(T1, b1) + 2) >> 2)
    # final form: 32-bit products of T >> 4 and b, PRMT 0x7632 picks the two high halves, IDP.2A adds them and the 2
    p0 = (T0 >> 4) * b0
    p1 = (T1 >> 4) * b1
    assert p0.max() < 2 ** 32 and p1.max() < 2 ** 32
    packed = (p0 >> 16) | ((p1 >> 16) << 16)
    s = (packed & 0xFFFF) * 1 + (packed >> 16) * 1 + 2
    assert s.max() <= 1023
    assert np.array_equal(ref, s >> 2)


def _dp4a(word, taps):
    """unsigned four-way dot product of the bytes of `word` with the bytes of the constant `taps` (byte 0 first)"""
    acc = np.zeros(word.shape, np.int64)
    for k in range(4):
        acc += ((word >> (8 * k)) & 0xFF) * ((taps >> (8 * k)) & 0xFF)
    return acc


def test_blur_row_sums_as_eight_dot_products():
    rng = np.random.default_rng(3)
    n = 200_000
    bytes12 = rng.integers(0, 256, (n, 12)).astype(np.int64)     # window: pixels x-4 .. x+7 of one input row
    word = lambda i: sum(bytes12[:, 4 * i + k] << (8 * k) for k in range(4))
    w0, w1, w2 = word(0), word(1), word(2)
    taps = np.array([18, 34, 48, 56, 48, 34, 18], np.int64)      # 7x7 sigma 2 kernel scaled to 256 (cv::getGaussianKernel, 8 bit)
    ref = [(bytes12[:, k + 1:k + 8] * taps).sum(axis=1) for k in range(4)]   # pixel k is byte 4 + k: bytes k+1 .. k+7
    funnel = lambda lo, hi, sh: ((lo >> sh) | (hi << (32 - sh))) & 0xFFFFFFFF
    X, Y = funnel(w0, w1, 16), funnel(w1, w2, 16)                # bytes 2..5, 6..9
    TAP_LO, TAP_HI = 0x38302212, 0x00122230
    h = [
        _dp4a(w0, 0x30221200) + _dp4a(w1, 0x12223038),
        _dp4a(X, TAP_LO) + _dp4a(Y, TAP_HI),
        _dp4a(X, 0x30221200) + _dp4a(Y, 0x12223038),
        _dp4a(w1, TAP_LO) + _dp4a(w2, TAP_HI),
    ]
    for k in range(4):
        assert np.array_equal(ref[k], h[k]), k
    assert all((r % 2 == 0).all() for r in ref)                  # every tap is even (noted in experiments/README.md)


def test_cvround_by_the_rounding_bias():
    rng = np.random.default_rng(4)
    v = np.concatenate([
        (rng.random(300_000, dtype=np.float32) * 64 - 32),                      # what a rotated pattern coordinate can be
        (np.arange(-4096, 4096, dtype=np.float32) * np.float32(0.5)),           # every tie in range: round half to even
        np.array([0.0, -0.0, 0.49999997, 0.5, 0.50000006, 1.5, 2.5, -0.5, -1.5, -2.5, 4194303.5, -4194303.5], np.float32),
    ]).astype(np.float32)
    biased = (v + np.float32(12582912.0)).astype(np.float32)     # one float32 add, round to nearest even
    got = biased.view(np.uint32).astype(np.int64) - 0x4B400000
    ref = np.rint(v.astype(np.float64)).astype(np.int64)         # cvRound == lrint: nearest, ties to even
    assert np.array_equal(ref, got)
    # the kernel never subtracts the bias per coordinate: r * W + c with both biased is off by 0x4B400000 * (W + 1) mod 2^32
    W = 80
    r, c = got[:1000], got[1000:2000]
    rb, cb = biased[:1000].view(np.uint32).astype(np.uint64), biased[1000:2000].view(np.uint32).astype(np.uint64)
    lhs = (rb * W + cb - 0x4B400000 * (W + 1)) & 0xFFFFFFFF
    assert np.array_equal(lhs, (r * W + c) & 0xFFFFFFFF)


def test_fast_run_tickets_cover_every_cell_once():
    RUN, TAIL = 8, 4

    def cells_of_ticket(t, total, n_warps):
        tail = min(total, n_warps * TAIL)
        n_runs = (total - tail) // RUN
        if t < n_runs:
            return range(t * RUN, t * RUN + RUN)
        first = n_runs * RUN + (t - n_runs)
        return range(first, first + 1) if first < total else range(0)

    for total, n_warps in ((0, 3552), (1, 8), (7, 8), (982, 3552), (14208, 3552), (14209, 3552), (15712, 3552), (100_003, 3552),
                           (4096 * 982, 3552), (31, 1), (64, 2)):
        seen = 0
        t = 0
        nxt = 0
        while True:
            cells = cells_of_ticket(t, total, n_warps)
            if len(cells) == 0:
                break
            assert cells[0] == nxt and cells[-1] < total        # consecutive, in range: exactly once, in order
            nxt = cells[-1] + 1
            seen += len(cells)
            t += 1
            if total > 1_000_000 and t > 2000 and cells_of_ticket(t + 10, total, n_warps):
                # long launches: check the run part analytically and jump to the single-cell tail
                tail = min(total, n_warps * TAIL)
                n_runs = (total - tail) // RUN
                if t < n_runs:
                    seen += (n_runs - t) * RUN
                    nxt = n_runs * RUN
                    t = n_runs
        assert seen == total and nxt == total
