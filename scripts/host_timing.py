"""Host-side timing of the per-cosmology C-ABI calls (grids, transfer, spectra)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from classpp_public_b200 import modules as M
name = sys.argv[1] if len(sys.argv) > 1 else "planck18"
inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
a = inp.arrays
class NL: nl_corr_density_m = a.get("nl.nl_corr_density_m")
nl = NL if NL.nl_corr_density_m is not None else None
ctx = M.Context(0)
t = time.perf_counter(); bg = M.BackgroundModule(inp, ctx); th = M.ThermodynamicsModule(inp, bg); print("bg+th upload %.1f ms" % ((time.perf_counter() - t) * 1e3))
for rep in range(3):
    t0 = time.perf_counter(); pt = M.PerturbationsModule(inp, bg, th, solve=False)
    t1 = time.perf_counter(); M.PerturbationsModule.solve_batch([pt])
    t2 = time.perf_counter(); tr = M.TransferModule(inp, bg, th, pt, nl, compute=False)
    t3 = time.perf_counter(); ctx.check(ctx._lib.clpp_transfer_compute(ctx.handle, M.capi.dptr(np.ascontiguousarray(nl.nl_corr_density_m)) if nl else None, 0, tr.info.q_size, ctx.err))
    t4 = time.perf_counter(); sp = M.SpectraModule(inp, pt, M.TabulatedPrimordial(a["pm.pk_at_transfer_k"]), nl, tr)
    t5 = time.perf_counter()
    print("rep %d: perturb grids %.1f ms | solve %.1f ms | transfer grids %.1f ms | transfer compute %.1f ms | spectra %.1f ms | kernels %s" % (
        rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t5 - t4) * 1e3, {k: round(v, 2) for k, v in ctx.kernel_ms().items()}))
