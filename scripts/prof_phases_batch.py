"""Per-phase device cycles under load: B cosmologies in one batched launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from classpp_public_b200 import modules as M
name = sys.argv[1] if len(sys.argv) > 1 else "planck18"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
ctxs, pts = [], []
for b in range(B):
    c = M.Context(0); bg = M.BackgroundModule(inp, c); th = M.ThermodynamicsModule(inp, bg)
    ctxs.append(c); pts.append(M.PerturbationsModule(inp, bg, th, solve=False))
M.PerturbationsModule.solve_batch(pts)
M.PerturbationsModule.solve_batch(pts)
print("B=%d perturb ms %.1f" % (B, ctxs[0].kernel_ms()["perturb"]))
prof = np.concatenate([p.kprofile_ for p in pts])  # [B*k][3][6]
neqs = sorted(set(prof[:, 0, :].ravel().tolist()) - {0})
tot = prof[:, 2, :].sum()
print("sum of mode cycles %.3e = %.1f s of one warp; per cosmology %.1f s" % (tot, tot / 1.965e9, tot / 1.965e9 / B))
for n in neqs:
    m = prof[:, 0, :] == n
    st = prof[:, 1, :][m].sum(); cy = prof[:, 2, :][m].sum()
    print("neq %4d: steps %9d cycles %.3e (%.1f%%)  cycles/step %8.0f" % (n, st, cy, 100 * cy / tot, cy / max(st, 1)))
k = pts[0].k_[0]
for ik in (len(k) - 1, len(k) - 2, len(k) - 8, len(k) - 14, len(k) - 30, 300):
    cyc = np.array([p.kprofile_[ik][2] for p in pts])  # [B][6]
    tot_m = cyc.sum(axis=1)
    print("k[%d]=%.3g: chain cycles min %.3e mean %.3e max %.3e (%.2f s) | per interval (mean) %s" % (
        ik, k[ik], tot_m.min(), tot_m.mean(), tot_m.max(), tot_m.max() / 1.965e9, ["%.2e" % v for v in cyc.mean(axis=0) if v > 0]))
