"""Developer harness (test infrastructure): runs the lane program of csrc/lane.cuh on the CPU for a few k modes of a golden
fixture and compares the source functions with the reference's columns stored in the fixture.
usage: python tests/hostsim/run_hostsim.py <fixture> [max_modes]"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from classpp_public_b200 import _capi as capi  # noqa: E402
from classpp_public_b200 import modules as M  # noqa: E402


def solve_columns(inp, k_cols):
    ctx = M.Context(device=-1)
    bg = M.BackgroundModule(inp, ctx)
    th = M.ThermodynamicsModule(inp, bg)
    pt = M.PerturbationsModule(inp, bg, th, solve=False)
    lib = C.CDLL(os.path.join(HERE, "libclpp_hostsim.so"))
    i = pt.info
    src = np.zeros((i.tp_size, i.k_size, i.tau_size))
    ks = (capi.KStat * i.k_size)()
    err = C.create_string_buffer(capi.ERRLEN)
    kl = np.asarray(k_cols, dtype=np.int32)
    lib.hostsim_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p]
    t0 = time.time()
    rc = lib.hostsim_solve(ctx.handle, kl.ctypes.data, len(kl), src.ctypes.data, ks, err)
    dt = time.time() - t0
    assert rc == 0, err.value
    kst = np.frombuffer(ks, dtype=np.dtype(capi.KStat)).copy()
    return pt, src, kst, dt


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "lcdm_coarse"
    inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    a = inp.arrays
    k_cols = a["ref.k_cols"].astype(int)
    if len(sys.argv) > 2:
        k_cols = k_cols[: int(sys.argv[2])]
    pt, src, kst, dt = solve_columns(inp, k_cols)
    ref = a["ref.sources_cols"]  # [tp][tau][col]
    print("fixture %s: %d modes in %.2f s" % (name, len(k_cols), dt))
    for c, ik in enumerate(k_cols):
        s = kst[ik]
        worst = []
        for tp in range(ref.shape[0]):
            r = ref[tp, :, c]
            m = src[tp, ik, :]
            scale = np.max(np.abs(r)) + 1e-300
            worst.append(float(np.max(np.abs(m - r)) / scale))
        worst = " ".join("%.1e" % w for w in worst)
        print("  k[%d]=%.4g status %d intervals %d steps %d failed %d fevals %d jac %d lu %d solves %d  max|dS|/max|S| per type = %s  iv_neq %s iv_steps %s"
              % (ik, pt.k_[0][ik], s["status"], s["intervals"], s["steps"], s["failed"], s["fevals"], s["jacobians"],
                 s["factorizations"], s["solves"], worst, list(s["iv_neq"][: s["intervals"]]), list(s["iv_steps"][: s["intervals"]])))


if __name__ == "__main__":
    main()
