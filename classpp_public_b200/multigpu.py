"""One cosmology over several GPUs (north star: cost-balanced k ranges, all-gather of S(k,tau) over NVLink before
the transfer stage, q ranges for the line-of-sight integrals, all-reduce of the partial C_l).

One process per GPU; `torch.distributed` (NCCL) is the plumbing, the device buffers are those of libclpp.so wrapped
zero-copy as torch tensors.  The exchange step is a real one (stage 2 needs S at ALL k for the cubic spline in k,
transfer_module.cpp:604), unlike the sweep path (sweep.py), which has no data-path collective.

`DistExchange` holds the two collectives (torch.distributed / NCCL; on host tensors over gloo in
tests/test_sweep_gloo.py::test_source_exchange_of_one_cosmology_world2).  The partition / merge logic around them is also
exercised on ONE GPU with the "ranks" emulated as contexts of the same process and the collectives replaced by direct
copies (tests/test_gpu_parity.py::test_one_cosmology_over_two_ranks_equals_single_gpu).
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from . import modules as M
from . import sweep


class _DevArray:
    """Zero-copy view of a device buffer of libclpp.so for torch (CUDA array interface v3)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def device_sources(pt):
    """torch view [tp][k][tau] of the device-resident source table of a PerturbationsModule."""
    import torch
    ctx, i = pt.ctx, pt.info
    p, n = C.c_void_p(), C.c_long()
    ctx.check(ctx._lib.clpp_perturb_device_sources(ctx.handle, C.byref(p), C.byref(n), ctx.err))
    t = torch.as_tensor(_DevArray(p.value, n.value), device="cuda:%d" % ctx.device)
    return t.view(i.tp_size, i.k_size, i.tau_size)


def q_range(q_size, rank, world):
    """Contiguous q range of `rank` (the cost of a q value is roughly uniform after the neglect tests)."""
    lo = (q_size * rank) // world
    hi = (q_size * (rank + 1)) // world
    return lo, hi


class DistExchange:
    """The two collectives over torch.distributed (NCCL on the GPUs of one box)."""

    def __init__(self, rank, world, group=None):
        self.rank, self.world, self.group = rank, world, group

    def allgather_columns(self, S, parts):
        """S: [tp][k][tau] device tensor with the columns parts[rank] filled; fills all the others."""
        import torch
        import torch.distributed as dist
        width = max(len(p) for p in parts)
        ntp, _, nt = S.shape
        send = torch.zeros(ntp, width, nt, dtype=S.dtype, device=S.device)
        mine = torch.as_tensor(parts[self.rank], device=S.device, dtype=torch.long)
        send[:, : len(mine), :] = S.index_select(1, mine)
        recv = [torch.empty_like(send) for _ in range(self.world)]
        dist.all_gather(recv, send, group=self.group)
        for r, p in enumerate(parts):
            if r != self.rank and len(p):
                idx = torch.as_tensor(p, device=S.device, dtype=torch.long)
                S.index_copy_(1, idx, recv[r][:, : len(p), :])
        if S.is_cuda:  # (host tensors: the gloo test of this logic, tests/test_sweep_gloo.py)
            torch.cuda.synchronize(S.device)

    def allreduce_sum(self, cl, device):
        """Sum of the partial C_l tables of all ranks; `device` None keeps the table on the host (gloo test)."""
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(cl).copy())
        if device is not None:
            t = t.to("cuda:%d" % device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()


def stage1_partition(inputs, ctx, world):
    """Grids (bit-exact, host) and the cost-balanced k partition, identical on every rank."""
    bg = M.BackgroundModule(inputs, ctx)
    th = M.ThermodynamicsModule(inputs, bg)
    pt = M.PerturbationsModule(inputs, bg, th, solve=False)
    parts = sweep.partition_modes_by_cost(pt.k_[0], world)
    return bg, th, pt, parts


def solve_modes(pt, indices):
    idx = np.ascontiguousarray(indices, dtype=np.int32)
    ctx = pt.ctx
    ctx.check(ctx._lib.clpp_perturb_solve_list(ctx.handle, capi.iptr(idx), len(idx), ctx.err))


def finish(inputs, bg, th, pt, primordial, nonlinear, rank, world, exchange):
    """Stages 2 and 3 of this rank's q range, then the all-reduce of the partial C_l."""
    tr = M.TransferModule(inputs, bg, th, pt, nonlinear, compute=False)
    lo, hi = q_range(tr.info.q_size, rank, world)
    nl = None
    if nonlinear is not None and getattr(nonlinear, "nl_corr_density_m", None) is not None:
        nl = np.ascontiguousarray(nonlinear.nl_corr_density_m, dtype=np.float64)
    ctx = pt.ctx
    ctx.check(ctx._lib.clpp_transfer_compute(ctx.handle, capi.dptr(nl), int(lo), int(hi), ctx.err))
    sp = M.SpectraModule(inputs, pt, primordial, nonlinear, tr, q_range=(lo, hi))
    cl = exchange.allreduce_sum(sp.cl_[0], ctx.device)
    return tr, sp, cl


def compute_cl_distributed(inputs, primordial, nonlinear, rank, world, device, group=None):
    """One cosmology on `world` GPUs (call from every rank). Returns (cl table [l_size*ct_size], ct_size)."""
    ctx = M.Context(device)
    bg, th, pt, parts = stage1_partition(inputs, ctx, world)
    solve_modes(pt, parts[rank])
    ex = DistExchange(rank, world, group)
    ex.allgather_columns(device_sources(pt), parts)
    if nonlinear is None and int(inputs.meta.get("nl.method", 0)) != 0:
        # `non linear = halofit`: every rank now holds delta_m at all k and runs the (3 ms) halofit step on its own device
        nonlinear = M.NonlinearModule(inputs, bg, pt, primordial)
    tr, sp, cl = finish(inputs, bg, th, pt, primordial, nonlinear, rank, world, ex)
    ct = sp.ct_size_
    ctx.close()
    return cl, ct
