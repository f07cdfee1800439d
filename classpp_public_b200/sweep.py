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


class SweepPipeline:
    """Throughput scheduler for a sweep of independent cosmologies on ONE GPU (BASELINE config 5).

    A batch of B cosmologies goes through `front(set, b)` (host: upstream tables -> PerturbationsModule without solving),
    ONE batched perturbation launch (`PerturbationsModule.solve_batch`) and `back(set, b, pt)` (halofit, transfer, spectra,
    lensing, P(k) on the cosmology's own stream).  The per-cosmology stages are short kernels and host work that leave
    the GPU mostly idle, so the pipeline keeps `n_sets` batches in flight on separate sets of contexts: `front` and `back`
    of one batch run under the batched launch of another.  By default the batched launches themselves do not overlap
    (`stagger=None`): measured on B200, two launches in flight finish no sooner than back to back, because the
    perturbation kernel saturates the SMs' instruction delivery at about half of the warp slots
    (profiles/README.md); `stagger=s` lets launches overlap, at least s seconds apart.  `submit()` returns at once;
    `drain()` waits for everything submitted so far.  No collective, no shared state between the sets: every context
    owns its streams and buffers."""

    def __init__(self, batch, front, back, n_sets=2, stagger=None, workers=None, on_solved=None):
        import os
        import threading
        from concurrent.futures import ThreadPoolExecutor
        self.batch, self.front, self.back, self.n_sets = int(batch), front, back, int(n_sets)
        self.stagger = None if stagger is None else float(stagger)
        self.on_solved = on_solved
        w = workers or min(self.batch, os.cpu_count() or 1)
        self._front_pool = ThreadPoolExecutor(max_workers=w)
        self._back_pool = ThreadPoolExecutor(max_workers=w)  # own queue: the next front never waits behind a back
        self._step_pool = ThreadPoolExecutor(max_workers=self.n_sets)
        self._tasks = [None] * self.n_sets
        self._n = 0
        self._gate = threading.Lock()
        self._t_last_launch = -1e30
        self.solve_seconds = []  # wall time of every batched launch, in order of completion
        self.launch_log = []     # (set, start, end) of every batched launch (time.perf_counter)

    def _run(self, s, args):
        import time
        from .modules import PerturbationsModule
        pts = list(self._front_pool.map(lambda b: self.front(s, b, *args), range(self.batch)))
        if self.stagger is None:  # one batched launch at a time; only `front`/`back` overlap it
            with self._gate:
                t0 = time.perf_counter()
                PerturbationsModule.solve_batch(pts)
        else:  # launches overlap, at least `stagger` seconds apart
            with self._gate:
                wait = self._t_last_launch + self.stagger - time.perf_counter()
                if wait > 0:
                    time.sleep(wait)
                self._t_last_launch = time.perf_counter()
            t0 = time.perf_counter()
            PerturbationsModule.solve_batch(pts)
        self.solve_seconds.append(time.perf_counter() - t0)
        self.launch_log.append((s, t0, time.perf_counter()))
        if self.on_solved is not None:
            self.on_solved(s)
        futs = [self._back_pool.submit(self.back, s, b, pts[b], *args) for b in range(self.batch)]
        return [f.result() for f in futs]

    def submit(self, *args):
        """Queue one batch on the next context set (waits only if that set is still busy with an older batch)."""
        s = self._n % self.n_sets
        self._n += 1
        if self._tasks[s] is not None:
            self._tasks[s].result()
        self._tasks[s] = self._step_pool.submit(self._run, s, args)
        return s

    def drain(self):
        """Wait for every submitted batch; returns the list of `back` results of the last batch of each set."""
        out = []
        for s in range(self.n_sets):
            if self._tasks[s] is not None:
                out.append(self._tasks[s].result())
        return out

    def close(self):
        self.drain()
        for p in (self._front_pool, self._back_pool, self._step_pool):
            p.shutdown(wait=True)
