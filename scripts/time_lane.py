"""Times the batched perturbation launch (lane kernel: one thread per mode) for batches of identical cosmologies and
compares it with the warp-per-mode kernels (CLPP_WARP_PATH=1).  usage: python scripts/time_lane.py [fixture] [batches...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from classpp_public_b200 import modules as M

name = sys.argv[1] if len(sys.argv) > 1 else "planck18"
batches = [int(x) for x in sys.argv[2:]] or [1, 32, 128]
inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def run(B, warp):
    if warp:
        os.environ["CLPP_WARP_PATH"] = "1"
    else:
        os.environ.pop("CLPP_WARP_PATH", None)
    ctxs, pts = [], []
    for _ in range(B):
        c = M.Context(0); b = M.BackgroundModule(inp, c); t = M.ThermodynamicsModule(inp, b)
        ctxs.append(c); pts.append(M.PerturbationsModule(inp, b, t, solve=False))
    ms = []
    for rep in range(2):
        t0 = time.time()
        M.PerturbationsModule.solve_batch(pts)
        wall = time.time() - t0
        ms.append(ctxs[0].kernel_ms()["perturb"])
    ks = pts[0].kstat_
    src = np.stack(pts[0].sources_[0])
    out = (min(ms), wall, int(ks[:, 0].sum()), int(ks[:, 1].sum()), int(ks[:, 2].sum()), int(ks[:, 4].sum()), src)
    for c in ctxs:
        c.close()
    return out


ref = None
for B in batches:
    for warp in ((False, True) if B <= 32 else (False,)):
        ms, wall, steps, failed, fevals, lus, src = run(B, warp)
        print("%s B=%d %s: perturb %.1f ms (%.2f ms per cosmology; wall of the call %.2f s); per cosmology steps %d failed %d fevals %d lu %d"
              % (name, B, "warp-per-mode" if warp else "lane (thread-per-mode)", ms, ms / B, wall, steps, failed, fevals, lus), flush=True)
        if ref is None:
            ref = src
        else:
            sc = np.max(np.abs(ref), axis=1, keepdims=True)
            print("   max |S - S_first| / max_tau |S| per type:", ["%.1e" % v for v in np.max(np.abs(src - ref) / np.where(sc > 0, sc, 1), axis=1)], flush=True)
