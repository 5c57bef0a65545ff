#!/usr/bin/env python3
"""MapPoint::ComputeDistinctiveDescriptors sharded over the GPUs of one box (shard.distinctive_sharded: cost-balanced
contiguous blocks of map points, no data-path collective, the two result ints per point all-gathered over NCCL), checked
against the single-GPU result of the same map.  Launch with torch.distributed.run, one rank per GPU."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-orb-slam-icra2018_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import orbb200  # noqa: E402
from orbb200 import shard  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = orbb200.Matcher(local)
rng = np.random.default_rng(3)                      # every rank rebuilds the same map
n_points = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
sizes = rng.integers(2, 41, n_points)
start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
desc = rng.integers(0, 256, (int(start[-1]), 32), dtype=np.uint8)
best, med, blocks = shard.distinctive_sharded(m, desc, start, dist)     # warm-up (allocations, NCCL channels)
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
best, med, blocks = shard.distinctive_sharded(m, desc, start, dist)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], device="cuda")
dist.all_reduce(dt, op=dist.ReduceOp.MAX)
ok = None
if rank == 0:
    t1 = time.perf_counter()
    want, wmed = m.distinctive_descriptors(desc, start)
    single = time.perf_counter() - t1
    ok = bool(np.array_equal(best, want) and np.array_equal(med, wmed))
    print(json.dumps({"workload": "ComputeDistinctiveDescriptors, %d map points, %d observations, sharded by cost-balanced blocks" % (n_points, int(start[-1])),
                      "n_gpus": world, "ms_sharded_host_to_host": float(dt.item()) * 1e3, "ms_single_gpu_host_to_host": single * 1e3,
                      "blocks": [list(map(int, b)) for b in blocks], "sharded_equals_single_gpu": ok}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok in (None, True) else 1)
