"""Multi-GPU sharding of the path (one process per GPU, torch.distributed for the plumbing).

The path shards by INDEPENDENT COSMOLOGIES (parameter sweeps, BASELINE config 5): cosmology i of a
sweep goes to rank i mod world, every rank pushes its share through `clpp_perturb_solve_batch` +
transfer + spectra on its own GPU, and only the small results (C_l tables) are gathered.  There is no
data-path collective.  For a single cosmology across GPUs the C ABI also exposes
k-range / q-range entry points; `partition_modes_by_cost` produces the cost-balanced k partition
(the number of NDF steps of a mode grows like k tau_0).
"""
import numpy as np


def shard_indices(n_items, rank, world):
    """Indices of the items (cosmologies) owned by `rank`: round-robin, so every rank gets the same mix."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, n_items, world))


def partition_modes_by_cost(k, world, cost_floor=None):
    """Greedy longest-first partition of the k modes over `world` ranks, cost_i ~ k_i + floor
    (the floor is the k-independent part: tight-coupling and hierarchy phases).
    Returns a list of index arrays (sorted), one per rank; they cover range(len(k)) exactly once."""
    k = np.asarray(k, dtype=np.float64)
    floor = float(np.median(k)) if cost_floor is None else float(cost_floor)
    cost = k + floor
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world)
    owner = np.empty(len(k), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += cost[i]
    return [np.sort(np.nonzero(owner == r)[0]) for r in range(world)]


def gather_tables(local, n_items, rank, world, group=None):
    """All-gather per-cosmology result tables.  `local` maps global cosmology index -> 1-D float64 array
    (all of the same length).  Returns the list of n_items arrays on every rank.  Works with the gloo
    backend on CPU tensors and with NCCL on CUDA tensors (device = that of the current CUDA device)."""
    import torch
    import torch.distributed as dist
    mine = shard_indices(n_items, rank, world)
    assert sorted(local) == mine, "rank %d holds %s, expected %s" % (rank, sorted(local), mine)
    per_rank = (n_items + world - 1) // world
    width = len(next(iter(local.values()))) if local else 0
    if world > 1:
        w = torch.tensor([width], dtype=torch.int64)
        if dist.get_backend(group) == "nccl":
            w = w.cuda()
        dist.all_reduce(w, op=dist.ReduceOp.MAX, group=group)
        width = int(w.item())
    buf = torch.zeros(per_rank, width, dtype=torch.float64)
    for j, i in enumerate(mine):
        buf[j] = torch.from_numpy(np.ascontiguousarray(local[i], dtype=np.float64))
    if world == 1:
        gathered = [buf]
    else:
        if dist.get_backend(group) == "nccl":
            buf = buf.cuda()
        gathered = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(gathered, buf, group=group)
    out = [None] * n_items
    for r in range(world):
        g = gathered[r].cpu().numpy()
        for j, i in enumerate(shard_indices(n_items, r, world)):
            out[i] = g[j].copy()
    return out
