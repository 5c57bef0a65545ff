"""GPU parity of the extractor (kernels 1-6) against the CPU oracle, stage by stage and end to end.

Bar: bit-exact for images, candidate lists (order included), selected set and order, descriptors, coordinates,
octaves, responses; angles within 1e-4 degrees (north star) -- and in practice bit-exact too, asserted separately.
"""
import numpy as np
import pytest

import orbb200
from orbb200.synth import synth_frame

pytestmark = pytest.mark.gpu

CASES = [
    # seed, w, h, nfeatures, noise_only
    (0, 752, 480, 1000, False),     # EuRoC mono (config 1)
    (1, 752, 480, 2000, False),     # EuRoC initialisation extractor (2 x nFeatures, Tracking.cc:822)
    (2, 1241, 376, 2000, False),    # KITTI (config 2)
    (3, 752, 480, 1000, True),      # pure-noise stress: ~20k level-0 candidates
    (4, 640, 480, 500, False),
    (5, 400, 300, 1500, False),     # more features wanted than the small levels can give
    (6, 333, 251, 300, False),      # odd sizes
]


def _assert_same_keypoints(got, ref):
    assert len(got) == len(ref)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(got[f], ref[f]), f
    assert np.allclose(got["angle"], ref["angle"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("seed,w,h,nf,noise", CASES)
def test_extract_matches_oracle(oracle, seed, w, h, nf, noise):
    img = synth_frame(seed, w, h, noise_only=noise)
    ex = orbb200.Extractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    oe = oracle.extractor(nf, 1.2, 8, 20, 7)
    kps, desc = ex(img)
    rk, rd = oe.extract(img)
    # constructor tables (ORBextractor.cc:412-472)
    t, rt = ex.tables(), oe.tables()
    for k in rt:
        assert np.array_equal(t[k], rt[k]), k
    # stage by stage
    for l in range(8):
        assert ex.level_size(l) == oe.level_size(l)
        assert np.array_equal(ex.level(l), oe.level_padded(l)), "pyramid level %d" % l
        c, rc = ex.candidates(l), oe.level_candidates(l)
        assert len(c) == len(rc), "candidate count level %d" % l
        assert c.tobytes() == rc.tobytes(), "candidates level %d" % l
        rb = oe.level_blurred(l)
        if rb is not None:
            assert np.array_equal(ex.blurred(l), rb), "blur level %d" % l
    # end to end
    _assert_same_keypoints(kps, rk)
    assert np.array_equal(kps["angle"].view(np.uint32), rk["angle"].view(np.uint32)), "angles not bit-exact"
    assert np.array_equal(desc, rd)
    assert len(kps) <= ex.capacity
    assert ex.launch_count() == 5      # fused pyramid, FAST, quadtree, blur, orientation + BRIEF
    ex.close()


@pytest.mark.parametrize("nf,scale,nlevels,ini,mn,w,h", [
    (300, 1.5, 4, 12, 5, 640, 480),      # other constructor arguments than the shipped YAMLs
    (1200, 1.2, 8, 12, 7, 752, 480),     # EuRoC stereo / KITTI04-12 settings (iniThFAST 12)
    (500, 2.0, 3, 20, 7, 800, 600),      # scale 2: the resize window is wider than the fast path assumes
    (400, 1.2, 1, 20, 7, 320, 240),      # a single level
    (1000, 1.1, 12, 30, 10, 752, 480),   # many shallow levels
])
def test_other_constructor_arguments(oracle, nf, scale, nlevels, ini, mn, w, h):
    img = synth_frame(11, w, h)
    ex = orbb200.Extractor(nf, scale, nlevels, ini, mn, max_width=w, max_height=h)
    oe = oracle.extractor(nf, scale, nlevels, ini, mn)
    kps, desc = ex(img)
    rk, rd = oe.extract(img)
    t, rt = ex.tables(), oe.tables()
    for k in rt:
        assert np.array_equal(t[k], rt[k]), k
    for l in range(nlevels):
        assert np.array_equal(ex.level(l), oe.level_padded(l)), "pyramid level %d" % l
    assert kps.tobytes() == rk.tobytes()
    assert np.array_equal(desc, rd)
    ex.close()


def test_size_sweep(oracle):
    """Odd image sizes and aspect ratios: FAST cells of every width modulo 4 and height parity, level widths that are
    not multiples of 4 or 16 (pyramid / blur column groups that straddle the border), tall and wide images, a contrast
    so low that most cells fall back to minThFAST, and a cell size near the 60-px maximum."""
    rng = np.random.default_rng(2024)
    sizes = [(97, 131), (131, 97), (255, 257), (331, 203), (402, 119), (260, 402), (513, 383), (640, 367), (89, 89),
             (178, 119)]   # 178x119 -> 59-px cells on level 0
    for i, (w, h) in enumerate(sizes):
        img = synth_frame(700 + i, w, h)
        if i % 3 == 1:
            img = (img.astype(np.int32) // 6 + 100).astype(np.uint8)      # low contrast: threshold fallback everywhere
        nf = int(rng.integers(50, 600))
        nlev = min(8, int(np.log(min(w, h) / 63.0) / np.log(1.2)) + 1)    # the top level must stay >= 62 px
        ex = orbb200.Extractor(nf, 1.2, nlev, 20, 7, max_width=w, max_height=h)
        oe = oracle.extractor(nf, 1.2, nlev, 20, 7)
        kps, desc = ex(img)
        rk, rd = oe.extract(img)
        for l in range(nlev):
            assert np.array_equal(ex.level(l), oe.level_padded(l)), (w, h, "pyramid level %d" % l)
            rb = oe.level_blurred(l)
            if rb is not None:
                assert np.array_equal(ex.blurred(l), rb), (w, h, "blur level %d" % l)
            assert ex.candidates(l).tobytes() == oe.level_candidates(l).tobytes(), (w, h, "FAST candidates level %d" % l)
        assert kps.tobytes() == rk.tobytes(), (w, h)
        assert np.array_equal(desc, rd), (w, h)
        ex.close()


def test_staged_batch_kernels_size_sweep(oracle):
    """Batches of 8 frames and more take the bulk-async staged pyramid and blur kernels (one band of rows per CTA, staged
    through shared memory): every padded pyramid level, every blurred level and the final keypoints / descriptors of the
    first and of the last frame of a 9-frame batch equal the oracle, for EuRoC, KITTI and odd sizes (level widths that are
    not multiples of 4 or 16, heights that are not multiples of the 16- and 32-row bands, odd band heights)."""
    for i, (w, h, nf) in enumerate([(752, 480, 1000), (1241, 376, 2000), (331, 203, 300), (255, 257, 200), (640, 367, 500),
                                    (513, 383, 400)]):
        nlev = min(8, int(np.log(min(w, h) / 63.0) / np.log(1.2)) + 1)
        frames = np.stack([synth_frame(900 + 10 * i + j, w, h) for j in range(9)])
        ex = orbb200.Extractor(nf, 1.2, nlev, 20, 7, max_width=w, max_height=h, max_batch=9)
        oe = oracle.extractor(nf, 1.2, nlev, 20, 7)
        res = ex.extract_batch(frames)
        for f in (0, 8):
            rk, rd = oe.extract(frames[f])
            for l in range(nlev):
                assert np.array_equal(ex.level(l, f), oe.level_padded(l)), (w, h, f, "pyramid level %d" % l)
                rb = oe.level_blurred(l)
                if rb is not None:
                    assert np.array_equal(ex.blurred(l, f), rb), (w, h, f, "blur level %d" % l)
            _assert_same_keypoints(res[f][0], rk)
            assert np.array_equal(res[f][1], rd)
        ex.close()


def test_host_pipeline_chunks_take_the_staged_kernels(oracle, monkeypatch):
    """The end-to-end path of bench.py: orbx_extract_batch cuts a host batch into pipeline chunks that work in arena slots
    [frameBase, frameBase + n).  With 16-frame chunks every chunk takes the staged pyramid / blur kernels and the TMA loads of
    FAST and BRIEF address the arena by frameBase + frame: 40 frames (chunks at frameBase 0, 16, ...), every frame's
    keypoints and descriptors equal the oracle's on the 5 distinct images the batch repeats."""
    monkeypatch.setenv("ORBB_PIPE_CHUNK", "16")
    # two of the five images are pure noise: their cells overflow the FAST kernel's shared-memory queue into the per-warp
    # global scratch, of which the two compute streams of the pipeline each own a set
    base = [synth_frame(950 + i, 752, 480, noise_only=(i in (1, 3))) for i in range(5)]
    frames = np.stack([base[i % 5] for i in range(40)])
    ex = orbb200.Extractor(1000, max_width=752, max_height=480, max_batch=40)
    oe = oracle.extractor(1000)
    ref = [oe.extract(b) for b in base]
    res = ex.extract_batch(frames)
    for i, (k, d) in enumerate(res):
        rk, rd = ref[i % 5]
        _assert_same_keypoints(k, rk)
        assert np.array_equal(d, rd), i
    ex.close()


def test_batch_equals_single_and_oracle(oracle):
    frames = np.stack([synth_frame(100 + i, 752, 480) for i in range(5)] + [synth_frame(200, 752, 480, noise_only=True)])
    ex = orbb200.Extractor(1000, max_width=752, max_height=480, max_batch=4)   # 6 frames -> chunks of 4 + 2
    oe = oracle.extractor(1000)
    res = ex.extract_batch(frames)
    for i, (k, d) in enumerate(res):
        rk, rd = oe.extract(frames[i])
        _assert_same_keypoints(k, rk)
        assert np.array_equal(d, rd)
    ex.close()


def test_strided_input_and_size_change(oracle):
    big = synth_frame(7, 800, 500)
    view = big[10:490, 20:772]                     # 752x480 view with stride 800
    ex = orbb200.Extractor(1000, max_width=752, max_height=480)
    oe = oracle.extractor(1000)
    k, d = ex(view)
    rk, rd = oe.extract(np.ascontiguousarray(view))
    _assert_same_keypoints(k, rk)
    assert np.array_equal(d, rd)
    small = synth_frame(8, 512, 384)               # a different size on the same handle
    k, d = ex(small)
    rk, rd = oe.extract(small)
    _assert_same_keypoints(k, rk)
    assert np.array_equal(d, rd)
    ex.close()


def test_empty_and_invalid(oracle):
    ex = orbb200.Extractor(1000, max_width=752, max_height=480)
    k, d = ex(np.zeros((0, 0), np.uint8))          # empty image: no keypoints, no error (ORBextractor.cc:1048)
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = ex(np.full((480, 752), 128, np.uint8))  # flat image: zero keypoints (descriptors released, :1080)
    assert len(k) == 0
    with pytest.raises(orbb200.OrbError):
        ex(np.zeros((40, 40), np.uint8))           # top level below 62 px: the reference divides by zero here
    with pytest.raises(orbb200.OrbError):
        orbb200.Extractor(0)
    ex.close()


def test_device_resident_batch(oracle):
    torch = pytest.importorskip("torch")
    frames = np.stack([synth_frame(300 + i, 752, 480) for i in range(3)])
    ex = orbb200.Extractor(1000, max_width=752, max_height=480, max_batch=3)
    d_img = torch.from_numpy(frames).cuda()
    cap = ex.capacity
    d_kps = torch.zeros((3, cap, 7), dtype=torch.int32, device="cuda")
    d_desc = torch.zeros((3, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(3, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d_img, d_kps, d_desc, d_n)
    ex.synchronize()
    n = d_n.cpu().numpy()
    kraw = d_kps.cpu().numpy()
    oe = oracle.extractor(1000)
    for i in range(3):
        rk, rd = oe.extract(frames[i])
        k = kraw[i, :n[i]].copy().view(orbb200.KP_DTYPE).reshape(-1)
        _assert_same_keypoints(k, rk)
        assert np.array_equal(d_desc[i, :n[i]].cpu().numpy(), rd)
    ex.close()


def test_full_size_batch_is_periodic(oracle):
    """BASELINE.json config 3 at full size: 4096 EuRoC frames in one device-resident call. The arena then spans more
    than 4 GB per buffer, so every frame offset must be 64-bit. Property: the batch repeats 8 distinct frames, hence
    the outputs must repeat with period 8, and the first period must equal the oracle."""
    torch = pytest.importorskip("torch")
    if torch.cuda.mem_get_info()[0] < 60e9:
        pytest.skip("needs ~35 GB of device memory")
    base = np.stack([synth_frame(400 + i, 752, 480) for i in range(8)])
    nF = 4096
    ex = orbb200.Extractor(1000, max_width=752, max_height=480, max_batch=nF)
    d_img = torch.from_numpy(base).cuda().repeat(nF // 8, 1, 1).contiguous()
    cap = ex.capacity
    d_kps = torch.zeros((nF, cap, 7), dtype=torch.int32, device="cuda")
    d_desc = torch.zeros((nF, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(nF, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d_img, d_kps, d_desc, d_n)
    ex.synchronize()
    n = d_n.view(nF // 8, 8)
    assert bool((n == n[0:1]).all())
    k = d_kps.view(nF // 8, 8, cap, 7)
    d = d_desc.view(nF // 8, 8, cap, 32)
    valid = (torch.arange(cap, device="cuda")[None, :] < n[0][:, None])           # (8, cap)
    assert bool(((k == k[0:1]) | ~valid[None, :, :, None]).all())
    assert bool(((d == d[0:1]) | ~valid[None, :, :, None]).all())
    oe = oracle.extractor(1000)
    n0 = n[0].cpu().numpy()
    for i in (0, 7):
        rk, rd = oe.extract(base[i])
        got = d_kps[nF - 8 + i, :n0[i]].cpu().numpy().copy().view(orbb200.KP_DTYPE).reshape(-1)   # from the LAST period
        _assert_same_keypoints(got, rk)
        assert np.array_equal(d_desc[nF - 8 + i, :n0[i]].cpu().numpy(), rd)
    ex.close()
