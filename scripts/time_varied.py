"""Batched perturbation launch over the bench workload (B different cosmologies: seed-0 Latin hypercube, Planck-18 settings,
tables from the drop-in library) under several launch configurations given as ENV=VAL[,ENV=VAL] strings.
usage: python scripts/time_varied.py B "CLPP_LANE=0" "CLPP_LANE=2,CLPP_LANE_KCUT=0.6" ..."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from concurrent.futures import ThreadPoolExecutor
import numpy as np
from classpp_public_b200 import modules as M, upstream
from classpp_public_b200.configs import CONFIGS

B = int(sys.argv[1])
pars = upstream.latin_hypercube_sweep(B, CONFIGS["planck18"], seed=0)
with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
    inps = list(ex.map(upstream.inputs_for, pars))
ctxs, pts = [], []
for inp in inps:
    c = M.Context(0); b = M.BackgroundModule(inp, c); t = M.ThermodynamicsModule(inp, b)
    ctxs.append(c); pts.append(M.PerturbationsModule(inp, b, t, solve=False))
for cfg in sys.argv[2:]:
    keys = []
    for kv in cfg.split(","):
        if kv:
            k_, v_ = kv.split("="); os.environ[k_] = v_; keys.append(k_)
    ms = []
    for rep in range(2):
        t0 = time.time()
        M.PerturbationsModule.solve_batch(pts)
        ms.append(ctxs[0].kernel_ms()["perturb"])
    steps = sum(int(p.kstat_[:, 0].sum()) for p in pts)
    print("B=%d %-50s perturb %.0f / %.0f ms (%.1f ms per cosmology), %d steps" % (B, cfg, ms[0], ms[1], min(ms) / B, steps), flush=True)
    for k_ in keys:
        os.environ.pop(k_, None)
