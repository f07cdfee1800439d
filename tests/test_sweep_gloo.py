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
