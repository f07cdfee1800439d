"""One perturbation launch of a small fixture (for ncu captures of the lane kernel) + cycles per step of each
approximation interval of a few modes.  usage: python scripts/prof_lane.py [fixture]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from classpp_public_b200 import modules as M

name = sys.argv[1] if len(sys.argv) > 1 else "lcdm_coarse"
inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
c = M.Context(0); b = M.BackgroundModule(inp, c); t = M.ThermodynamicsModule(inp, b)
pt = M.PerturbationsModule(inp, b, t)
print(name, "perturb ms", c.kernel_ms()["perturb"])
tab = pt._kstat_table()
nk = pt.info.k_size
for ik in sorted(set([0, nk // 4, nk // 2, 3 * nk // 4, nk - 2, nk - 1])):
    s = tab[ik]
    n = int(s["intervals"])
    print("k[%d]=%.4g steps %d failed %d: " % (ik, pt.k_[0][ik], s["steps"], s["failed"]) +
          "  ".join("neq %d: %d steps, %.0f cyc/step" % (s["iv_neq"][i], s["iv_steps"][i], s["iv_cycles"][i] / max(1, s["iv_steps"][i]))
                    for i in range(n)))
c.close()
