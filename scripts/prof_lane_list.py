"""A short lane-kernel launch for ncu: modes [k0, k1) of a fixture only (clpp_perturb_solve_list).
usage: python scripts/prof_lane_list.py fixture k0 k1"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
from classpp_public_b200 import modules as M

name, k0, k1 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
c = M.Context(0); b = M.BackgroundModule(inp, c); t = M.ThermodynamicsModule(inp, b)
pt = M.PerturbationsModule(inp, b, t, solve=False)
kl = np.arange(k0, k1, dtype=np.int32)
c.check(c._lib.clpp_perturb_solve_list(c.handle, kl.ctypes.data_as(C.POINTER(C.c_int)), len(kl), c.err))
print(name, k0, k1, "perturb ms", c.kernel_ms()["perturb"])
pt._fetch_kstat()
tab = pt._kstat_table()
for ik in (k0, k1 - 1):
    s = tab[ik]
    n = int(s["intervals"])
    print("k[%d]=%.4g steps %d failed %d: " % (ik, pt.k_[0][ik], s["steps"], s["failed"]) +
          "  ".join("neq %d: %d steps, %.0f cyc/step" % (s["iv_neq"][i], s["iv_steps"][i], s["iv_cycles"][i] / max(1, s["iv_steps"][i]))
                    for i in range(n)))
c.close()
