"""The lane program (classpp_public_b200/csrc/lane.cuh: one GPU thread integrates one k mode) compiled as plain C++ by
tests/hostsim/ and executed on the CPU, against the reference's source functions stored in the golden fixtures.  This is TEST
INFRASTRUCTURE for the device code's logic (layouts, approximation switching, structured Newton solve, register tail): the
product library never runs it -- libclpp.so has no CPU path, tests/test_abi.py::test_version_and_host_only_context."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from classpp_public_b200 import _capi as capi
from classpp_public_b200 import modules as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HS = os.path.join(ROOT, "tests", "hostsim")


@pytest.fixture(scope="module")
def hostsim():
    subprocess.check_call(["make", "-C", HS], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(HS, "libclpp_hostsim.so"))
    lib.hostsim_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p]
    return lib


def solve_columns(lib, inp, k_cols):
    ctx = M.Context(device=-1)
    bg = M.BackgroundModule(inp, ctx)
    th = M.ThermodynamicsModule(inp, bg)
    pt = M.PerturbationsModule(inp, bg, th, solve=False)
    i = pt.info
    src = np.zeros((i.tp_size, i.k_size, i.tau_size))
    ks = (capi.KStat * i.k_size)()
    err = C.create_string_buffer(capi.ERRLEN)
    kl = np.asarray(k_cols, dtype=np.int32)
    assert lib.hostsim_solve(ctx.handle, kl.ctypes.data, len(kl), src.ctypes.data, ks, err) == 0, err.value
    return pt, src, np.frombuffer(ks, dtype=np.dtype(capi.KStat)).copy()


@pytest.mark.parametrize("name,dense", [("lcdm_coarse", False), ("lcdm_coarse", True), ("planck18", False), ("ncdm3_coarse", False)])
def test_lane_program_sources_vs_golden(hostsim, golden, monkeypatch, name, dense):
    """Source functions of the sub-sampled k columns: delta_m, delta_cb and phi+psi at the integrator tolerance (1e-4 of the
    column maximum), the temperature / polarisation sources (cancellations amplify the tolerance) at 1e-2, like the GPU test;
    dense = the dense-LU hub solve instead of the structured one (same step counts)."""
    if dense:
        monkeypatch.setenv("HOSTSIM_DENSE", "1")
    inp = golden(name)
    a = inp.arrays
    k_cols = a["ref.k_cols"].astype(int)
    if name == "ncdm3_coarse":
        k_cols = k_cols[:6]  # the top columns of the 316-equation system take seconds each on one CPU core
    pt, src, kst = solve_columns(hostsim, inp, k_cols)
    ref = a["ref.sources_cols"]  # [tp][tau][col]
    i = pt.info
    for c, ik in enumerate(k_cols):
        assert kst[ik]["status"] == 0 and kst[ik]["steps"] > 100
        for tp in range(i.tp_size):
            r, m = ref[tp, :, c], src[tp, ik, :]
            tol = 1e-4 if tp in (i.index_tp_delta_m, i.index_tp_delta_cb, i.index_tp_phi_plus_psi) else 1e-2
            assert np.max(np.abs(m - r)) <= tol * np.max(np.abs(r)), (name, ik, tp)
