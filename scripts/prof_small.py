"""Small profiling target: one pass of the hot path on the coarse LambdaCDM fixture (or argv[1])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from classpp_public_b200 import modules as M
name = sys.argv[1] if len(sys.argv) > 1 else "lcdm_coarse"
inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
a = inp.arrays
ctx = M.Context(0)
bg = M.BackgroundModule(inp, ctx); th = M.ThermodynamicsModule(inp, bg)
class NL: nl_corr_density_m = a.get("nl.nl_corr_density_m")
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    pt = M.PerturbationsModule(inp, bg, th)
    tr = M.TransferModule(inp, bg, th, pt, NL if NL.nl_corr_density_m is not None else None)
    sp = M.SpectraModule(inp, pt, M.TabulatedPrimordial(a["pm.pk_at_transfer_k"]), None, tr)
print("ok", ctx.kernel_ms(), "launches", ctx.launch_count)
