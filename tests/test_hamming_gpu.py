"""GPU parity of the Hamming matcher (kernel 7) against the oracle restatement of ORBmatcher.cc:1675-1691, 432-512."""
import numpy as np
import pytest

from datagen import planted_descriptors

pytestmark = pytest.mark.gpu


def test_distance_kat(matcher, oracle):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (4096, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (4096, 32), dtype=np.uint8)
    a[0] = 0; b[0] = 255          # zeros vs ones = 256
    a[1] = b[1]                   # identical = 0
    a[2] = b[2]; a[2, 7] ^= 0x10  # single-bit flip = 1
    d = matcher.distance(a, b)
    assert d[0] == 256 and d[1] == 0 and d[2] == 1
    ref = np.array([oracle.distance(x, y) for x, y in zip(a[:512], b[:512])])
    assert np.array_equal(d[:512], ref)
    pc = np.unpackbits(a ^ b, axis=1).sum(1)
    assert np.array_equal(d, pc)


@pytest.mark.parametrize("nq,nt,ratio,ori", [(2000, 2000, 0.9, True), (513, 1025, 0.6, True), (7, 3, 0.9, False),
                                            (1, 1, 0.9, True), (300, 1, 0.9, True), (1000, 777, 0.75, False)])
def test_bruteforce_matches_oracle(matcher, oracle, nq, nt, ratio, ori):
    rng = np.random.default_rng(nq * 7919 + nt)
    q, qa, t, ta = planted_descriptors(rng, nq, nt)
    if nt > 4:   # exact duplicates: ties on best and second-best must resolve to the lowest index
        t[3] = t[1]
        q[0] = t[1]
    r = matcher.bruteforce(q, qa, t, ta, ratio, ori)
    n, best, second, idx, m12 = oracle.bruteforce(q, qa, t, ta, ratio, ori)
    assert np.array_equal(r["best"], best)
    assert np.array_equal(r["second"], second)
    assert np.array_equal(r["idx"], idx)
    assert np.array_equal(r["matches12"], m12)
    assert int(r["nmatches"]) == n
    if nq >= 300 and nt >= 300:
        assert n > 0


def test_bruteforce_batched_pairs(matcher, oracle):
    rng = np.random.default_rng(5)
    P, nq, nt = 9, 640, 700
    sets = [planted_descriptors(rng, nq, nt) for _ in range(P)]
    q = np.stack([s[0] for s in sets]); qa = np.stack([s[1] for s in sets])
    t = np.stack([s[2] for s in sets]); ta = np.stack([s[3] for s in sets])
    r = matcher.bruteforce(q, qa, t, ta, 0.9, True)
    for p in range(P):
        n, best, second, idx, m12 = oracle.bruteforce(q[p], qa[p], t[p], ta[p], 0.9, True)
        assert np.array_equal(r["best"][p], best) and np.array_equal(r["second"][p], second)
        assert np.array_equal(r["idx"][p], idx) and np.array_equal(r["matches12"][p], m12)
        assert r["nmatches"][p] == n


def test_bruteforce_full_size_properties(matcher):
    """Config-4 size (2000x2000, many pairs): size-independent properties instead of the slow oracle."""
    rng = np.random.default_rng(11)
    P, n = 16, 2000
    q = rng.integers(0, 256, (P, n, 32), dtype=np.uint8)
    perm = np.stack([rng.permutation(n) for _ in range(P)])
    t = np.stack([q[p][perm[p]] for p in range(P)])          # trains are a permutation of the queries
    a = np.zeros((P, n), np.float32)
    r = matcher.bruteforce(q, a, t, a, 0.9, True)
    inv = np.argsort(perm, axis=1)
    assert np.array_equal(r["idx"], inv)                     # every query finds its own copy ...
    assert (r["best"] == 0).all()                            # ... at distance 0
    assert (r["second"] > 60).all()                          # random 256-bit strings are far apart
    assert np.array_equal(r["matches12"], inv) and (r["nmatches"] == n).all()


def test_allpairs_counts(matcher, oracle):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(21)
    nkf, nd = 12, 500
    table = rng.integers(0, 256, (nkf, nd, 32), dtype=np.uint8)
    ang = (rng.random((nkf, nd)) * 360).astype(np.float32)
    # keyframes 2k and 2k+1 see the same scene: planted near-duplicates incl. exact ties and repeated structure
    for k in range(0, nkf, 2):
        q, qa, t, ta = planted_descriptors(rng, nd, nd, frac=0.5, max_flip=40)
        t[5] = t[9]
        table[k], ang[k], table[k + 1], ang[k + 1] = q, qa, t, ta
    d_table = torch.from_numpy(table).cuda()
    d_ang = torch.from_numpy(ang).cuda()
    for ratio, ori in ((0.75, True), (0.9, False)):
        counts = torch.full((nkf, nkf), -7, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()      # the matcher enqueues on its own stream
        matcher.allpairs_device(d_table, d_ang, 0, nkf, 0, nkf, ratio, ori, counts)
        matcher.synchronize()
        torch.cuda.synchronize()
        got = counts.cpu().numpy()
        ref = np.array([[oracle.kf_pair(table[i], ang[i], table[j], ang[j], ratio, ori)[0] for j in range(nkf)]
                        for i in range(nkf)])
        assert np.array_equal(got, ref)
        assert ref.max() > 100
    # sub-ranges (what a rank computes in the sharded run) agree with the full matrix
    part = torch.full((4, nkf), -7, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    matcher.allpairs_device(d_table, d_ang, 4, 8, 0, nkf, 0.9, False, part)
    matcher.synchronize()
    assert np.array_equal(part.cpu().numpy(), ref[4:8])


def test_allpairs_sharded_single_rank(matcher, oracle):
    """orbm_allpairs_sharded through the C ABI with a one-rank NCCL communicator: the own-block path, q_count, and the
    segment launches agree with the oracle (the N > 1 gathers are checked against the single-GPU kernel by
    bench.py --gpus N, allpairs.check, and by tools/gpu_multi_ap.sh)."""
    import torch
    import orbb200
    rng = np.random.default_rng(23)
    nkf, nd = 9, 200
    table = np.zeros((nkf, nd, 32), np.uint8)
    ang = np.zeros((nkf, nd), np.float32)
    for k in range(0, nkf - 1, 2):
        q, qa, t, ta = planted_descriptors(rng, nd, nd, frac=0.5, max_flip=40)
        table[k], ang[k], table[k + 1], ang[k + 1] = q, qa, t, ta
    table[8] = rng.integers(0, 256, (nd, 32), dtype=np.uint8)
    d_table, d_ang = torch.from_numpy(table).cuda(), torch.from_numpy(ang).cuda()
    ref = np.array([[oracle.kf_pair(table[i], ang[i], table[j], ang[j], 0.75, True)[0] for j in range(nkf)] for i in range(nkf)])
    comm = orbb200.Comm(orbb200.Comm.unique_id(), 0, 1, 0)
    try:
        assert comm.nccl_version() >= 21000
        for q_count in (-1, 4):
            rows = nkf if q_count < 0 else q_count
            counts = torch.full((rows, nkf), -7, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            matcher.allpairs_sharded(comm, d_table, d_ang, [nkf], 0.75, True, counts, chunk_kf=4, q_count=q_count)
            matcher.synchronize()
            assert np.array_equal(counts.cpu().numpy(), ref[:rows])
        assert comm.last_gather()["chunks"] == 0          # one rank: nothing to gather
    finally:
        comm.close()


def test_popc_peak_plausible(matcher):
    v = matcher.popc_peak()
    assert 1e12 < v < 2e13
