"""Host-side logic of the N>1 path on CPU: world_size-2 gloo, one process per (pretend) GPU.
The data path has no collective (replicas); what is distributed is the assignment of cosmologies
to ranks and the gathering of the small result tables."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from classpp_public_b200 import sweep


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_cl(i, width):
    # stands for the C_l table of cosmology i (the GPU stages are covered by tests marked gpu)
    return np.arange(width, dtype=np.float64) * (i + 1) + 0.25 * i


def _worker(rank, world, port, n_items, width, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sweep.shard_indices(n_items, rank, world)
        local = {i: _fake_cl(i, width) for i in mine}
        full = sweep.gather_tables(local, n_items, rank, world)
        ok = all(np.array_equal(full[i], _fake_cl(i, width)) for i in range(n_items))
        ret[rank] = (ok, mine)
    finally:
        dist.destroy_process_group()


def _exchange_worker(rank, world, port, k, ret):
    """The exchange step of the one-cosmology-over-N-GPUs path (multigpu.DistExchange) on host tensors over gloo:
    ragged cost-balanced k partitions, all-gather of the source columns, all-reduce of the partial C_l."""
    import torch
    from classpp_public_b200 import multigpu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts = sweep.partition_modes_by_cost(k, world)
        ntp, nt = 3, 11
        full = torch.arange(ntp * len(k) * nt, dtype=torch.float64).reshape(ntp, len(k), nt) * 0.5 + 1.0
        S = torch.zeros_like(full)
        mine = torch.as_tensor(parts[rank], dtype=torch.long)
        S[:, mine, :] = full[:, mine, :]  # what stage 1 of this rank produced
        ex = multigpu.DistExchange(rank, world)
        ex.allgather_columns(S, parts)
        cl = ex.allreduce_sum(np.full((4, 7), float(rank + 1)), None)
        ret[rank] = (bool(torch.equal(S, full)), bool(np.all(cl == sum(range(1, world + 1)))), [len(p) for p in parts])
    finally:
        dist.destroy_process_group()


def test_source_exchange_of_one_cosmology_world2(golden):
    k = np.asarray(golden("lcdm_coarse").arrays["ref.k"], dtype=np.float64)[:-2]  # odd count: ragged partitions
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_exchange_worker, args=(world, port, k, ret), nprocs=world, join=True)
        for r in range(world):
            assert ret[r][0] and ret[r][1], ret[r]
        assert sum(ret[0][2]) == len(k)


@pytest.mark.parametrize("n_items", [5, 8])
def test_sweep_shard_and_gather_world2(n_items):
    world, width = 2, 791  # 113 l values x 7 spectra
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_items, width, ret), nprocs=world, join=True)
        assert all(ret[r][0] for r in range(world))
        owned = sorted(i for r in range(world) for i in ret[r][1])
        assert owned == list(range(n_items))  # every cosmology exactly once


def test_shard_indices_cover_exactly_once():
    for n in (0, 1, 7, 1024):
        for world in (1, 2, 4, 8):
            allidx = sorted(i for r in range(world) for i in sweep.shard_indices(n, r, world))
            assert allidx == list(range(n))
            sizes = [len(sweep.shard_indices(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sweep.shard_indices(4, 2, 2)


def test_cost_balanced_k_partition(golden):
    k = golden("planck18").arrays["ref.k"]
    for world in (2, 4, 8):
        parts = sweep.partition_modes_by_cost(k, world)
        assert sorted(np.concatenate(parts).tolist()) == list(range(len(k)))
        cost = [float(np.sum(k[p] + np.median(k))) for p in parts]
        # the longest chain (k_max) bounds the balance; everything else is spread evenly
        assert max(cost) <= max(1.05 * np.mean(cost), float(k[-1] + np.median(k)) * 1.001)


def test_sweep_pipeline_scheduling_on_cpu(monkeypatch):
    """Host logic of sweep.SweepPipeline without a GPU (the batched launch is stubbed): every batch passes through
    front -> solve -> back on alternating context sets, launches never overlap by default, a set is only reused after
    its previous batch has been consumed, and drain() returns the last results of every set."""
    import threading
    import time
    from classpp_public_b200 import modules as M
    from classpp_public_b200.sweep import SweepPipeline

    log, lock = [], threading.Lock()
    in_launch = [0]

    def fake_solve_batch(pts):
        with lock:
            in_launch[0] += 1
            assert in_launch[0] == 1, "two batched launches in flight"
            log.append(("solve", pts[0][0], pts[0][2]))
        time.sleep(0.02)
        with lock:
            in_launch[0] -= 1

    monkeypatch.setattr(M.PerturbationsModule, "solve_batch", staticmethod(fake_solve_batch))
    busy = {0: False, 1: False}

    def front(s, b, tag):
        if b == 0:
            with lock:
                assert not busy[s], "context set reused while its previous batch is still in flight"
                busy[s] = True
        return (s, b, tag)

    def back(s, b, pt, tag):
        assert pt == (s, b, tag)
        time.sleep(0.005)
        return tag * 100 + b

    solved = []

    def on_solved(s):
        solved.append(s)

    B = 4
    pipe = SweepPipeline(B, front, back, n_sets=2, workers=4, on_solved=on_solved)
    orig_run = pipe._run

    def run_and_release(s, args):
        out = orig_run(s, args)
        with lock:
            busy[s] = False
        return out

    pipe._run = run_and_release
    for tag in range(5):
        assert pipe.submit(tag) == tag % 2
    res = pipe.drain()
    assert sorted(t for _, _, t in log) == [0, 1, 2, 3, 4]
    assert [s for _, s, _ in log].count(0) == 3 and solved.count(1) == 2
    assert res == [[400 + b for b in range(B)], [300 + b for b in range(B)]]
    assert len(pipe.solve_seconds) == 5
    # overlapping mode: launches may be concurrent but are at least `stagger` apart
    pipe.close()
    starts = []

    def fake_solve_batch2(pts):
        starts.append(time.perf_counter())
        time.sleep(0.05)

    monkeypatch.setattr(M.PerturbationsModule, "solve_batch", staticmethod(fake_solve_batch2))
    pipe = SweepPipeline(2, lambda s, b: (s, b), lambda s, b, pt: 0, n_sets=2, stagger=0.03, workers=2)
    pipe.submit(); pipe.submit()
    pipe.drain()
    assert len(starts) == 2 and abs(starts[1] - starts[0]) >= 0.029 and abs(starts[1] - starts[0]) < 0.05
    pipe.close()
