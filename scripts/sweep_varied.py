"""Batched perturbation launch over DIFFERENT cosmologies (Latin hypercube around Planck-18, +-`spread` of each parameter):
how much of the cohort speed-up (modes of neighbouring cosmologies in lockstep) survives when the cosmologies differ.
Upstream tables come from the unmodified reference (oracle/_ref, thermodynamics level only).
usage: python scripts/sweep_varied.py [n_cosmologies=32] [spread=0.1]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from scipy.stats import qmc
from oracle import refprobe
from refutil import inputs_from_reference
from classpp_public_b200 import modules as M
from classpp_public_b200.configs import CONFIGS

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
spread = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
base = CONFIGS["planck18"]
keys = ["omega_b", "omega_cdm", "H0", "A_s", "n_s", "tau_reio"]
u = qmc.LatinHypercube(d=len(keys), seed=1).random(n)
inps = []
t0 = time.time()
for row in u:
    par = dict(base)
    for k_, x in zip(keys, row):
        par[k_] = float(base[k_]) * (1.0 + spread * (2.0 * x - 1.0))
    ref = refprobe.RefCosmology(par, threads=os.cpu_count()).compute("thermodynamics")
    inps.append(inputs_from_reference(ref))
    ref.close()
print("reference tables for %d cosmologies: %.1f s" % (n, time.time() - t0))
for label, these in (("identical", [inps[0]] * n), ("varied +-%g" % spread, inps)):
    for W, WL in ((1, 1), (4, 1), (4, 2), (4, 4)):
        os.environ["CLPP_COHORT"] = str(W)
        os.environ["CLPP_COHORT_LONG"] = str(WL)
        ctxs, pts = [], []
        for inp in these:
            c = M.Context(0); b = M.BackgroundModule(inp, c); t = M.ThermodynamicsModule(inp, b)
            ctxs.append(c); pts.append(M.PerturbationsModule(inp, b, t, solve=False))
        ms = []
        for rep in range(2):
            M.PerturbationsModule.solve_batch(pts)
            ms.append(ctxs[0].kernel_ms()["perturb"])
        steps = sum(int(p.kstat_[:, 0].sum()) for p in pts)
        print("%-14s modes/CTA %d (long group %d): perturb %.0f ms (%.1f ms per cosmology), %d steps, k sizes %s" %
              (label, W, WL, min(ms), min(ms) / n, steps, sorted({p.info.k_size for p in pts})))
        for c in ctxs:
            c.close()
