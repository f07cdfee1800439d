import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from classpp_public_b200 import modules as M
from test_gpu_parity import run_pipeline
inp = M.Inputs.load('/root/repo/tests/golden/lcdm_coarse.npz')
ctx, pt, tr, sp = run_pipeline(inp); a = sp.cl_[0].copy(); ks_a = pt.kstat_[:, :6].sum(axis=0); ctx.close()
os.environ["CLPP_GENERIC_ONLY"] = "1"
ctx, pt, tr, sp = run_pipeline(inp); b = sp.cl_[0].copy(); ks_b = pt.kstat_[:, :6].sum(axis=0); ctx.close()
nz = b != 0
r = np.abs(a[nz] / b[nz] - 1)
print("max rel diff", r.max(), "median", np.median(r), "steps", ks_a, ks_b)
ref = inp.arrays["ref.cl"]
nzr = ref != 0
print("vs reference: specialised", np.abs(a[nzr] / ref[nzr] - 1).max(), "generic", np.abs(b[nzr] / ref[nzr] - 1).max())
