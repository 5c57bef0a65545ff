"""MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:257-322): the oracle restatement
against the reference's own text of that function (compiled in place, oracle/_ref/libdistinctive_ref.so), and the batched
CUDA kernel against the oracle.  Integer work: exact."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_distinctive  # noqa: E402


def observation_sets(seed, sizes):
    """per map point: noisy copies of one 256-bit descriptor (plus outliers and exact duplicates), as CSR"""
    rng = np.random.default_rng(seed)
    rows, start = [], [0]
    for n in sizes:
        base = rng.integers(0, 2, 256, dtype=np.uint8)
        for i in range(n):
            r = base.copy()
            k = rng.integers(0, 70) if rng.random() < 0.85 else rng.integers(90, 128)
            r[rng.permutation(256)[:k]] ^= 1
            rows.append(np.packbits(r))
        if n >= 4 and rng.random() < 0.5:       # exact duplicates: equal medians, the first row must win
            rows[-1] = rows[-3].copy()
        start.append(start[-1] + n)
    desc = np.stack(rows) if rows else np.zeros((0, 32), np.uint8)
    return desc, np.array(start, np.int32)


SIZES = [1, 2, 3, 4, 5, 8, 13, 32, 33, 64, 100, 257, 0, 7, 2, 2, 6, 31, 9, 500]


def test_median_rule_known_answers(oracle):
    z = np.zeros((1, 32), np.uint8)
    assert oracle.distinctive(z, [0, 1])[0][0] == 0
    assert oracle.distinctive(np.zeros((0, 32), np.uint8), [0, 0])[0][0] == -1
    # three rows a, b, c with d(a,b)=1, d(a,c)=3, d(b,c)=2: sorted rows (0,1,3) (0,1,2) (0,2,3), rank (size_t)(0.5*2) = 1
    # -> medians 1, 1, 2: rows a and b tie, the first wins
    d = np.zeros((3, 32), np.uint8)
    d[1, 0] = 0b1
    d[2, 0] = 0b111
    best, med = oracle.distinctive(d, [0, 3])
    assert best[0] == 0 and med[0] == 1
    # two rows: rank (size_t)(0.5*1) = 0 -> the median is the self-distance 0 for both, row 0 wins
    best, med = oracle.distinctive(d[1:], [0, 2])
    assert best[0] == 0 and med[0] == 0


@pytest.mark.skipif(not ref_distinctive.available() and not os.path.isdir("/root/reference"),
                    reason="oracle/_ref/libdistinctive_ref.so is built only where /root/reference is mounted")
def test_oracle_equals_reference(oracle):
    if not ref_distinctive.available():
        ref_distinctive.build()
    for seed in range(4):
        desc, start = observation_sets(seed, SIZES)
        best, med = oracle.distinctive(desc, start)
        for p in range(len(start) - 1):
            assert ref_distinctive.distinctive(desc[start[p]:start[p + 1]]) == best[p], (seed, p)
    # bad keyframes are skipped by the reference; the C ABI expects them removed
    desc, start = observation_sets(9, [40])
    bad = (np.arange(40) % 3 == 0).astype(np.uint8)
    good = desc[bad == 0]
    assert ref_distinctive.distinctive(desc, bad) == oracle.distinctive(good, [0, len(good)])[0][0]


@pytest.mark.gpu
def test_gpu_equals_oracle(matcher, oracle):
    for seed in range(4):
        desc, start = observation_sets(seed, SIZES)
        best, med = matcher.distinctive_descriptors(desc, start)
        rbest, rmed = oracle.distinctive(desc, start)
        assert np.array_equal(best, rbest)
        assert np.array_equal(med[rbest >= 0], rmed[rbest >= 0])
    assert matcher.launch_count() == 1


@pytest.mark.gpu
def test_gpu_batch_of_map_points(matcher, oracle):
    """a LocalMapping-sized batch: 20 000 map points with 2..40 observations each in ONE launch"""
    rng = np.random.default_rng(3)
    sizes = rng.integers(2, 41, 20000)
    start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    desc = rng.integers(0, 256, (int(start[-1]), 32), dtype=np.uint8)
    proto = rng.integers(0, 256, (len(sizes), 32), dtype=np.uint8)
    owner = np.repeat(np.arange(len(sizes)), sizes)
    desc = proto[owner] ^ (desc & rng.integers(0, 256, desc.shape, dtype=np.uint8) & rng.integers(0, 256, desc.shape, dtype=np.uint8))
    best, med = matcher.distinctive_descriptors(desc, start)
    rbest, rmed = oracle.distinctive(desc, start)
    assert np.array_equal(best, rbest) and np.array_equal(med, rmed)


@pytest.mark.gpu
def test_gpu_rejects_bad_runs(matcher):
    import orbb200
    with pytest.raises(orbb200.OrbError):
        matcher.distinctive_descriptors(np.zeros((2, 32), np.uint8), [0, 2, 1])
    with pytest.raises(orbb200.OrbError):
        matcher.distinctive_descriptors(np.zeros((2, 32), np.uint8), [1, 2])
    best, med = matcher.distinctive_descriptors(np.zeros((0, 32), np.uint8), [0])
    assert len(best) == 0
