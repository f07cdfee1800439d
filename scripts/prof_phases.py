"""Per-phase device profile of the perturbation kernel: cycles / steps per approximation interval."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from classpp_public_b200 import modules as M
name = sys.argv[1] if len(sys.argv) > 1 else "planck18"
inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
ctx = M.Context(0)
bg = M.BackgroundModule(inp, ctx); th = M.ThermodynamicsModule(inp, bg)
pt = M.PerturbationsModule(inp, bg, th)
pt = M.PerturbationsModule(inp, bg, th)
print("perturb ms", ctx.kernel_ms()["perturb"])
prof = pt.kprofile_  # [k][3][6]
ks = pt.kstat_
neqs = sorted(set(prof[:, 0, :].ravel().tolist()) - {0})
tot_cyc = prof[:, 2, :].sum()
print("sum of mode cycles %.3e  (= %.3f s of one warp at 1.965 GHz)" % (tot_cyc, tot_cyc / 1.965e9))
for n in neqs:
    m = prof[:, 0, :] == n
    st = prof[:, 1, :][m].sum(); cy = prof[:, 2, :][m].sum()
    print("neq %4d: modes %4d steps %9d cycles %.3e (%.1f%%)  cycles/step %8.0f" % (n, m.sum(), st, cy, 100 * cy / tot_cyc, cy / max(st, 1)))
per_mode = prof[:, 2, :].sum(axis=1)
o = np.argsort(-per_mode)[:12]
k = pt.k_[0]
for i in o:
    print("k[%d]=%.4g cycles %.3e (%.3f s) steps %d fevals %d lu %d solves %d  neq/steps/cyc:" % (i, k[i], per_mode[i], per_mode[i] / 1.965e9, ks[i, 0], ks[i, 2], ks[i, 4], ks[i, 5]),
          [(int(a), int(b), "%.2e" % c) for a, b, c in zip(*prof[i]) if a])
print("totals: steps %d failed %d fevals %d jac %d lu %d solves %d" % tuple(ks[:, :6].sum(axis=0)))
ctx.close()
