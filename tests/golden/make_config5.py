"""BASELINE config 5 through the UNMODIFIED reference (test infrastructure; needs /root/reference and oracle/_ref):
the n-point seed-0 Latin hypercube (Planck-18 settings, full default grids), one single-threaded reference process per core.
Per cosmology: the C_l table at the l nodes, lensed TT/EE/TE/BB at every 7th l, linear and halofit P_m(k, z=0), the k grid.
Output: tests/golden/_big/config5_lhs<n>.npz (git-ignored: ~35 KB per cosmology; it travels to the GPU box with the tree,
where tests/config5_sweep.py compares the device path with it and writes the summary that IS committed under profiles/).
usage: python tests/golden/make_config5.py [n=1024] [processes=cpu_count]"""
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
L_STRIDE = 7


def one(par):
    from oracle.refprobe import RefCosmology
    t0 = time.time()
    try:
        ref = RefCosmology(par, threads=1).compute("lensing")
    except Exception as e:  # a point the reference itself rejects is recorded, not dropped
        return {"error": str(e)}
    lt = ref.iscalar("le.lt_size")
    lens = ref.get("le.cl_lensed").reshape(-1, lt)
    out = {"cl": ref.get("sp.cl"), "cl_lensed": lens[2::L_STRIDE].copy(), "k": ref.get("pt.k"),
           "pk_lin_m": ref.get("nl.pk_lin_m_at_pt_k"), "pk_nl_m": ref.get("nl.pk_nl_m_at_pt_k"),
           "seconds": time.time() - t0}
    ref.close()
    return out


def main():
    from classpp_public_b200.configs import CONFIGS
    from classpp_public_b200.upstream import latin_hypercube_sweep
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else os.cpu_count()
    pars = latin_hypercube_sweep(n, CONFIGS["planck18"], seed=0)
    out = {"params": np.array(json.dumps(pars)), "l_stride": np.array(L_STRIDE)}
    t0 = time.time()
    failures = []
    with ProcessPoolExecutor(max_workers=procs) as ex:
        for i, r in enumerate(ex.map(one, pars)):
            if "error" in r:
                failures.append((i, r["error"]))
                continue
            for k, v in r.items():
                out["%d__%s" % (i, k)] = np.asarray(v)
            if i % 32 == 0:
                print("config5 reference %d / %d  (%.0f s)" % (i, n, time.time() - t0), flush=True)
    out["failures"] = np.array(json.dumps(failures))
    out["wall_s"] = np.array(time.time() - t0)
    out["processes"] = np.array(procs)
    os.makedirs(os.path.join(HERE, "_big"), exist_ok=True)
    path = os.path.join(HERE, "_big", "config5_lhs%d.npz" % n)
    np.savez(path, **out)
    print("config5 reference -> %s  %.1f MB, %d failures, %.0f s on %d processes"
          % (path, os.path.getsize(path) / 1e6, len(failures), time.time() - t0, procs))


if __name__ == "__main__":
    main()
