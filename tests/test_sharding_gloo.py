"""Host-side logic of the multi-GPU path on CPU: two ranks over gloo.

The CUDA kernel is replaced by an injected host stand-in so that only the plumbing is exercised here: block ownership,
the all-gather of the descriptor table (including uneven blocks), the own-block-first column schedule, and the assembly
of the per-rank tile.  The kernel itself is covered by tests/test_hamming_gpu.py::test_allpairs_counts.
"""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

from orbb200 import shard


def test_block_ranges_cover_everything():
    for n in (0, 1, 7, 8, 4096, 8191):
        for world in (1, 2, 3, 8):
            blocks = [shard.block_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1
            for r in range(world):
                sched = shard.column_schedule(n, r, world)
                assert sched[0] == blocks[r]                      # own block first: no communication needed
                assert sorted(sched) == sorted(blocks)            # every column block exactly once


def _popcount_rows(a, b):
    return np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(2)


def _worker(rank, world, port, n_kf, n_desc, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(99)                                # every rank can rebuild the global table
    table = rng.integers(0, 256, (n_kf, n_desc, 32), dtype=np.uint8)
    angles = rng.random((n_kf, n_desc)).astype(np.float32)
    b, e = shard.block_range(n_kf, rank, world)
    calls = []

    def stand_in(tab, ang, q0, q1, d0, d1, out):
        # "match count" stand-in: number of descriptor pairs closer than 110 bits; order of calls is recorded
        calls.append((d0, d1))
        t = tab.numpy()
        for qi in range(q0, q1):
            for dj in range(d0, d1):
                out[qi - q0, dj] = int((_popcount_rows(t[qi], t[dj]) < 110).sum())

    counts = shard.allpairs_sharded(None, torch.from_numpy(table[b:e].copy()), torch.from_numpy(angles[b:e].copy()), n_kf,
                                    dist, compute=stand_in)
    np.save(os.path.join(out_dir, "counts_%d.npy" % rank), counts.numpy())
    np.save(os.path.join(out_dir, "calls_%d.npy" % rank), np.array(calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_kf", [6, 7])    # 7: uneven blocks, the gather runs on padded blocks
def test_allpairs_sharded_two_ranks(tmp_path, n_kf):
    world, n_desc = 2, 24
    port = 29500 + os.getpid() % 2000 + n_kf
    mp.spawn(_worker, args=(world, port, n_kf, n_desc, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(99)
    table = rng.integers(0, 256, (n_kf, n_desc, 32), dtype=np.uint8)
    full = np.array([[int((_popcount_rows(table[i], table[j]) < 110).sum()) for j in range(n_kf)] for i in range(n_kf)])
    got = np.concatenate([np.load(tmp_path / ("counts_%d.npy" % r)) for r in range(world)])
    assert np.array_equal(got, full)
    for r in range(world):
        calls = np.load(tmp_path / ("calls_%d.npy" % r))
        assert len(calls) == world
        max_local = -(-n_kf // world)
        assert calls[0][0] == r * max_local                       # own block first


def test_cost_balanced_blocks():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        for costs in (rng.integers(1, 1600, 5000), np.ones(7), np.array([100.0, 1, 1, 1]), np.zeros(5), np.zeros(0)):
            blocks = shard.cost_balanced_blocks(costs, world)
            assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == len(costs)
            assert all(blocks[i][1] == blocks[i + 1][0] and blocks[i][0] <= blocks[i][1] for i in range(world - 1))
            if len(costs) >= 1000:                                    # many small units: every rank is within 5 % of its share
                share = [float(np.sum(costs[b:e])) for b, e in blocks]
                assert max(share) < 1.05 * sum(share) / world


def _distinctive_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from oracle_py import Oracle
    o = Oracle()
    rng = np.random.default_rng(4)                                   # every rank rebuilds the same map
    sizes = rng.integers(0, 30, 301)
    start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    desc = rng.integers(0, 256, (int(start[-1]), 32), dtype=np.uint8)
    best, med, blocks = shard.distinctive_sharded(None, desc, start, dist, compute=lambda d, s: o.distinctive(d, s))
    np.save(os.path.join(out_dir, "best_%d.npy" % rank), best)
    np.save(os.path.join(out_dir, "blocks_%d.npy" % rank), np.array(blocks))
    dist.barrier()
    dist.destroy_process_group()


def test_distinctive_sharded_two_ranks(tmp_path, oracle):
    """map points shard by cost-balanced contiguous blocks with no data-path collective; every rank ends with all results"""
    world = 2
    port = 29500 + os.getpid() % 2000 + 17
    mp.spawn(_distinctive_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(4)
    sizes = rng.integers(0, 30, 301)
    start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    desc = rng.integers(0, 256, (int(start[-1]), 32), dtype=np.uint8)
    want, _ = oracle.distinctive(desc, start)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("best_%d.npy" % r)), want)
    blocks = np.load(tmp_path / "blocks_0.npy")
    assert blocks[0][0] == 0 and blocks[-1][1] == 301 and blocks[0][1] == blocks[1][0]
    cost = sizes.astype(np.float64) ** 2
    assert abs(cost[: blocks[0][1]].sum() - cost[blocks[0][1]:].sum()) < 0.05 * cost.sum()
