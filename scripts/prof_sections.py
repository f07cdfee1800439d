"""Section profile (library built with EXTRA=-DPT_PROF): cycles per step of each code section, per interval."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from classpp_public_b200 import modules as M
name, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
ctx = M.Context(0)
bg = M.BackgroundModule(inp, ctx); th = M.ThermodynamicsModule(inp, bg)
pt = M.PerturbationsModule(inp, bg, th, k_range=(lo, hi))
names = ["env", "rhs", "solve", "predict", "update", "control", "factor", "adjust", "output", "env_bg", "difupd", "env_th"]
for ik in range(lo, hi):
    pr = pt.kprofile_[ik]; sec = pt.ksections_[ik].reshape(6, 12); ks = pt.kstat_[ik]
    print("k[%d] fevals/step(all)=%.2f lu/step=%.2f" % (ik, ks[2] / ks[0], ks[4] / ks[0]))
    for iv in range(int(ks[6])):
        steps = max(pr[1][iv], 1)
        print("   neq=%3d steps=%6d cyc/step=%7.0f :" % (pr[0][iv], steps, pr[2][iv] / steps),
              " ".join("%s=%.0f" % (n, sec[iv][i] / steps) for i, n in enumerate(names)))
ctx.close()
