#!/usr/bin/env python3
"""Config 5: all-pairs keyframe descriptor matching sharded by query block, NCCL all-gather of the descriptor table.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/allpairs_sharded.py [--kf 2048] [--desc 1000] [--check]

Each rank generates (deterministically) the descriptors of its own keyframe block, all-gathers the table, and computes
its (n_local x n_kf) tile of match counts; the gather overlaps the matching of the rank's own block.  --check recomputes
the full matrix on rank 0 alone and compares.  Prints one JSON line (rank 0): compares/s aggregate, max over ranks.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-orb-slam-icra2018_b200"))
import orbb200  # noqa: E402
from orbb200 import shard  # noqa: E402


def block_descriptors(b, e, n_desc, device):
    """keyframe k's descriptors depend only on k; neighbours (k, k^1) share planted near-duplicates."""
    out = torch.empty((e - b, n_desc, 32), dtype=torch.uint8, device=device)
    ang = torch.empty((e - b, n_desc), dtype=torch.float32, device=device)
    for k in range(b, e):
        g = torch.Generator(device=device); g.manual_seed(1000 + (k >> 1))
        base = torch.randint(0, 256, (n_desc, 32), dtype=torch.uint8, device=device, generator=g)
        a = torch.rand(n_desc, device=device, generator=g) * 360
        if k & 1:   # the odd twin: half of the rows get a few flipped bits, the rest are fresh
            g2 = torch.Generator(device=device); g2.manual_seed(77 + k)
            noise = (torch.randint(0, 256, (n_desc, 32), dtype=torch.uint8, device=device, generator=g2)
                     & torch.randint(0, 256, (n_desc, 32), dtype=torch.uint8, device=device, generator=g2)
                     & torch.randint(0, 256, (n_desc, 32), dtype=torch.uint8, device=device, generator=g2)
                     & torch.randint(0, 256, (n_desc, 32), dtype=torch.uint8, device=device, generator=g2))
            fresh = torch.randint(0, 256, (n_desc, 32), dtype=torch.uint8, device=device, generator=g2)
            base = torch.where((torch.arange(n_desc, device=device) % 2 == 0)[:, None], base ^ noise, fresh)
            a = (a + 20.0) % 360
        out[k - b] = base
        ang[k - b] = a
    return out, ang


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kf", type=int, default=2048)
    ap.add_argument("--desc", type=int, default=1000)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    m = orbb200.Matcher(local)
    b, e = shard.block_range(a.kf, rank, world)
    d_local, a_local = block_descriptors(b, e, a.desc, dev)
    st = torch.cuda.Stream(device=dev)
    shard.allpairs_sharded(m, d_local, a_local, a.kf, dist, 0.75, True, torch_stream=st)   # warm-up (NCCL channels, attributes)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    counts = shard.allpairs_sharded(m, d_local, a_local, a.kf, dist, 0.75, True, torch_stream=st)
    e1.record(st)
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = None
    if a.check:
        gathered = [torch.empty((shard.block_range(a.kf, r, world)[1] - shard.block_range(a.kf, r, world)[0], a.kf), dtype=torch.int32, device=dev)
                    for r in range(world)]
        dist.all_gather(gathered, counts)
        if rank == 0:
            full_d, full_a = block_descriptors(0, a.kf, a.desc, dev)
            ref = torch.empty((a.kf, a.kf), dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            m.allpairs_device(full_d, full_a, 0, a.kf, 0, a.kf, 0.75, True, ref)
            m.synchronize()
            ok = bool(torch.equal(torch.cat(gathered), ref))
    if rank == 0:
        ms = float(t.item())
        print(json.dumps({"workload": "configs[4] shape: all-pairs keyframe matching, %d keyframes x %d descriptors, sharded by query block, "
                                      "NCCL all-gather of the table overlapped with the own-block matching" % (a.kf, a.desc),
                          "n_gpus": world, "ms": ms, "value": a.kf * a.kf * a.desc * a.desc / (ms * 1e-3), "unit": "compares/s",
                          "matches_total": int(counts.sum().item()), "sharded_equals_single_gpu": ok}))
    m.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
