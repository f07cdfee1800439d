"""One cosmology through every stage (perturbations -> halofit -> transfer -> spectra -> lensing -> P(k)): the launch list and
per-kernel ncu captures of profiles/ are taken on this script.  usage: python scripts/prof_stages.py [fixture] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from classpp_public_b200 import modules as M
from classpp_public_b200.configs import CONFIGS

name = sys.argv[1] if len(sys.argv) > 1 else "planck18"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
par = CONFIGS[name]
prim = M.AnalyticPrimordial(par.get("A_s", 2.215e-9), par.get("n_s", 0.9619))
tabs, pts = [], []
for _ in range(B):
    c = M.Context(0); b = M.BackgroundModule(inp, c); t = M.ThermodynamicsModule(inp, b)
    tabs.append((c, b, t)); pts.append(M.PerturbationsModule(inp, b, t, solve=False))
M.PerturbationsModule.solve_batch(pts)
for (c, b, t), pt in zip(tabs, pts):
    nl = M.NonlinearModule(inp, b, pt, prim) if int(inp.meta["nl.method"]) else None
    tr = M.TransferModule(inp, b, t, pt, nl)
    sp = M.SpectraModule(inp, pt, prim, nl, tr)
    le = M.LensingModule(inp, sp)
    pk = pt.pk_linear(prim.pk_at_k(pt.k_[0]))
print(name, B, tabs[0][0].kernel_ms(), "n_points", tr.info.n_points, "n_integrals", tr.info.n_integrals)
