"""Multi-GPU plumbing (one process per GPU, torch.distributed): who owns which frames / keyframes, and the one
exchange step of the path -- the all-gather of the keyframe descriptor table for all-pairs matching.

Frame batches and frame pairs are independent units: ranks take contiguous blocks and nothing crosses NVLink.
All-pairs keyframe matching has exactly one exchange: every rank contributes the descriptors of its own keyframe block
and needs everybody else's.  The gather is started first, and the rank matches its query block against its OWN block
(already local) while the gather is in flight; the remaining column blocks follow once it has landed.
"""
import numpy as np


def block_range(n, rank, world):
    """Contiguous block [begin, end) of n units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def column_schedule(n_kf, rank, world):
    """Order in which a rank visits the db keyframe blocks: its own block first (no communication needed), then the
    others in ring order so that ranks do not all hammer the same source at once."""
    order = [(rank + k) % world for k in range(world)]
    return [block_range(n_kf, r, world) for r in order]


def all_gather_table(local_desc, local_angles, n_kf, dist, async_op=True):
    """All-gather the descriptor table. local_desc: (n_local, n_desc, 32) uint8 tensor, local_angles: (n_local, n_desc)
    float32, on this rank's device (or CPU for gloo).  Returns (table, angles, work_handles); wait on the handles before
    reading rows outside this rank's block.  Blocks may differ in size by one keyframe, so the gather is done on padded
    blocks and compacted (the padding rows are never read)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [block_range(n_kf, r, world) for r in range(world)]
    max_local = max(e - b for b, e in sizes)
    n_desc = local_desc.shape[1]
    pad_d = torch.zeros((max_local, n_desc, 32), dtype=torch.uint8, device=local_desc.device)
    pad_a = torch.zeros((max_local, n_desc), dtype=torch.float32, device=local_desc.device)
    pad_d[: local_desc.shape[0]] = local_desc
    pad_a[: local_angles.shape[0]] = local_angles
    gd = torch.empty((world * max_local, n_desc, 32), dtype=torch.uint8, device=local_desc.device)
    ga = torch.empty((world * max_local, n_desc), dtype=torch.float32, device=local_desc.device)
    # the own block is valid immediately: copy it in place so that matching against it can start before the gather ends
    gd[rank * max_local: rank * max_local + local_desc.shape[0]] = local_desc
    ga[rank * max_local: rank * max_local + local_angles.shape[0]] = local_angles
    works = [dist.all_gather_into_tensor(gd, pad_d, async_op=async_op),
             dist.all_gather_into_tensor(ga, pad_a, async_op=async_op)]
    return gd, ga, [w for w in works if w is not None], max_local, sizes


def make_comm(dist, device):
    """orbm_comm for this process: rank 0 creates the ncclUniqueId, torch.distributed only carries its 128 bytes."""
    import torch
    import orbb200
    world, rank = dist.get_world_size(), dist.get_rank()
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.from_numpy(orbb200.Comm.unique_id().copy())
    dev = torch.device("cuda", device)
    uid = uid.to(dev)
    dist.broadcast(uid, 0)
    return orbb200.Comm(uid.cpu().numpy(), rank, world, device)


def allpairs_sharded(matcher, local_desc, local_angles, n_kf, dist, ratio=0.75, check_ori=True, compute=None,
                     torch_stream=None, comm=None, chunk_kf=128):
    """Config 5: this rank's (n_local x n_kf) tile of the keyframe match-count matrix.

    On CUDA this is a thin caller of the C ABI (orbm_allpairs_sharded: chunked ncclAllGather on the communicator's copy
    stream, matching of the landed chunks on `torch_stream`); `comm` is an orbb200.Comm (made here if None).  The call
    returns with the work enqueued.
    With `compute` given (the gloo tests: a host stand-in for the kernel) the same block layout is exercised through
    torch.distributed instead: compute(table, angles, q_begin, q_end, db_begin, db_end, counts).
    """
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [block_range(n_kf, r, world) for r in range(world)]
    qb, qe = sizes[rank]
    if compute is None:
        if not local_desc.is_cuda:
            raise RuntimeError("allpairs_sharded: the CUDA path needs device tensors (there is no CPU fallback)")
        own_comm = comm is None
        if own_comm:
            comm = make_comm(dist, local_desc.device.index)
        if torch_stream is None:
            torch_stream = torch.cuda.Stream(device=local_desc.device)
        torch_stream.wait_stream(torch.cuda.current_stream(local_desc.device))   # the producers of local_desc
        counts = torch.full((qe - qb, n_kf), -1, dtype=torch.int32, device=local_desc.device)
        per = [e - b for b, e in sizes]
        with torch.cuda.stream(torch_stream):
            matcher.allpairs_sharded(comm, local_desc.contiguous(), local_angles.contiguous(), per, ratio, check_ori, counts,
                                     chunk_kf=chunk_kf, stream=torch_stream.cuda_stream)
        if own_comm:
            torch_stream.synchronize()
            comm.close()
        return counts
    gd, ga, works, max_local, sizes = all_gather_table(local_desc, local_angles, n_kf, dist)
    counts = torch.full((qe - qb, n_kf), -1, dtype=torch.int32, device=local_desc.device)
    padded_rows = world * max_local
    # rows of the padded table: rank r's keyframes start at r*max_local
    padded_counts = torch.full((qe - qb, padded_rows), -1, dtype=torch.int32, device=local_desc.device)
    # the gather's output includes this rank's own slot: wait for it before anything reads the table (this path is the host
    # stand-in of the tests; the CUDA path orders its reads against the chunked gathers by events in csrc/shard.cu)
    for w in works:
        w.wait()
    for (b, e) in column_schedule(n_kf, rank, world):
        r = next(i for i, s in enumerate(sizes) if s == (b, e))
        compute(gd, ga, rank * max_local, rank * max_local + (qe - qb), r * max_local, r * max_local + (e - b), padded_counts)
    for r, (b, e) in enumerate(sizes):
        counts[:, b:e] = padded_counts[:, r * max_local: r * max_local + (e - b)]
    return counts


def cost_balanced_blocks(costs, world):
    """Contiguous blocks [begin, end) of units with unequal cost, one per rank, cut where the running cost crosses k/world
    of the total (every rank gets a block, possibly empty; the blocks cover all units in order)."""
    costs = np.asarray(costs, np.float64)
    n = len(costs)
    if n == 0:
        return [(0, 0)] * world
    cum = np.cumsum(costs)
    total = cum[-1]
    cuts = [0]
    for k in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * k / world, side="left")) + 1 if total > 0 else (n * k) // world)
    cuts = [min(max(c, cuts[i - 1] if i else 0), n) for i, c in enumerate(cuts)]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def distinctive_sharded(matcher, desc, start, dist, compute=None):
    """MapPoint::ComputeDistinctiveDescriptors for a map too large for one GPU's turn-around: map points are independent
    units, so ranks take contiguous blocks balanced by cost (n^2 distances per point with n observations) and NOTHING of
    the data path crosses NVLink; only the per-point results (two ints) are all-gathered so that every rank ends with the
    full answer.  desc: (total, 32) uint8 host array, start: CSR run starts (len = points + 1), both known to every rank.

    compute(desc_block, start_block) -> (best, median) defaults to the CUDA kernel through the C ABI; the gloo tests inject
    a host stand-in."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    start = np.asarray(start, np.int64)
    n_points = len(start) - 1
    sizes = np.diff(start)
    blocks = cost_balanced_blocks(sizes.astype(np.float64) ** 2, world)
    b, e = blocks[rank]
    if compute is None:
        def compute(d, s):
            return matcher.distinctive_descriptors(d, s)
    if e > b:
        local_start = (start[b:e + 1] - start[b]).astype(np.int32)
        best, med = compute(np.ascontiguousarray(desc[start[b]:start[e]]), local_start)
    else:
        best, med = np.zeros(0, np.int32), np.zeros(0, np.int32)
    max_local = max(1, max(y - x for x, y in blocks))
    pad = torch.full((2, max_local), -2, dtype=torch.int32)
    pad[0, : e - b] = torch.from_numpy(np.asarray(best, np.int32))
    pad[1, : e - b] = torch.from_numpy(np.asarray(med, np.int32))
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    pad = pad.to(dev)
    gathered = torch.empty((world * 2, max_local), dtype=torch.int32, device=dev)   # rank blocks concatenated along dim 0
    dist.all_gather_into_tensor(gathered, pad)
    gathered = gathered.cpu().numpy().reshape(world, 2, max_local)
    out_best = np.empty(n_points, np.int32)
    out_med = np.empty(n_points, np.int32)
    for r, (x, y) in enumerate(blocks):
        out_best[x:y] = gathered[r, 0, : y - x]
        out_med[x:y] = gathered[r, 1, : y - x]
    return out_best, out_med, blocks
