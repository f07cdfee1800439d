"""Profiling target: integrate a few k modes [lo,hi) of a fixture (perturbation kernel only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from classpp_public_b200 import modules as M
name, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
ctx = M.Context(0)
bg = M.BackgroundModule(inp, ctx); th = M.ThermodynamicsModule(inp, bg)
pt = M.PerturbationsModule(inp, bg, th, k_range=(lo, hi))
print("perturb ms", ctx.kernel_ms()["perturb"], "steps", pt.kstat_[lo:hi, 0].tolist())
ctx.close()
