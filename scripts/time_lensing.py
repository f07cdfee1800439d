"""Device time of the lensing stage alone (one Planck-18 cosmology): python scripts/time_lensing.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from classpp_public_b200 import modules as M
from test_gpu_parity import run_pipeline
inp = M.Inputs.load("tests/golden/planck18.npz")
ctx, pt, tr, sp = run_pipeline(inp)
for acc in (0, 1):
    inp.meta["pr.accurate_lensing"] = acc
    for rep in range(3):
        t0 = time.perf_counter()
        le = M.LensingModule(inp, sp)
        t1 = time.perf_counter()
        print("accurate=%d rep %d: wall %.2f ms, kernels %.3f ms" % (acc, rep, (t1 - t0) * 1e3, ctx.kernel_ms()["lensing"]))
