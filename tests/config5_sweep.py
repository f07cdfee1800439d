"""BASELINE config 5 on the device, every spectrum against the unmodified reference (a one-off parity run, not collected by
pytest: `python tests/config5_sweep.py [n=1024] [batch=128] [first=n]` on a GPU box; `first`: only the first so many points of
the n-point fixture).

The n-point seed-0 Latin hypercube over (omega_b, omega_cdm, h, ln10^10A_s, n_s, tau_reio) with the Planck-18 settings goes
through the sweep scheduler (classpp_public_b200/sweep.py: batches of `batch` cosmologies, one batched perturbation launch
each, the per-cosmology stages of a batch under the launch of the next one).  Upstream tables (background, thermodynamics)
come from the drop-in library on the host threads and are inside the clock.  Every cosmology is compared with
tests/golden/_big/config5_lhs<n>.npz (tests/golden/make_config5.py: the reference run as single-threaded processes):
k grid bit-exact; unlensed C_l^{TT,EE,TE,pp,Tp,Ep} at the l nodes, lensed TT/EE/TE/BB at every 7th l, linear and halofit
P_m(k, z=0) as maximum relative errors (cross spectra relative to the geometric mean of the auto spectra).
Writes gpurun_out/config5_lhs<n>.json: worst and median error per spectrum, the number of cosmologies above 1e-4, failures,
spectra/s including the upstream host time.  This script never touches oracle/: it only reads the reference's stored output."""
import json
import os
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from classpp_public_b200 import modules as M  # noqa: E402
from classpp_public_b200 import upstream  # noqa: E402
from classpp_public_b200.sweep import SweepPipeline  # noqa: E402

SPECTRA = ("tt", "ee", "te", "pp", "tp", "ep", "lensed_tt", "lensed_ee", "lensed_te", "lensed_bb", "pk_lin", "pk_nl")
TOL = 1e-4


def cl_errors(sp, ref_cl):
    ct = sp.ct_size_
    cl, ref, i = sp.cl_[0].reshape(-1, ct), ref_cl.reshape(-1, ct), sp.info
    e = {}
    for name in ("tt", "ee", "pp"):
        c = getattr(i, "index_ct_" + name)
        e[name] = float(np.max(np.abs(cl[:, c] / ref[:, c] - 1.0)))
    for name, (x, y) in {"te": ("tt", "ee"), "tp": ("tt", "pp"), "ep": ("ee", "pp")}.items():
        c = getattr(i, "index_ct_" + name)
        norm = np.sqrt(ref[:, getattr(i, "index_ct_" + x)] * ref[:, getattr(i, "index_ct_" + y)])
        e[name] = float(np.max(np.abs(cl[:, c] - ref[:, c]) / norm))
    return e


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    with np.load(os.path.join(HERE, "golden", "_big", "config5_lhs%d.npz" % n)) as f:
        z = {k_: f[k_] for k_ in f.files}  # read up front: an NpzFile is a zip archive and must not be read from several threads
    pars = json.loads(str(z["params"]))
    ref_failures = dict((int(i), m) for i, m in json.loads(str(z["failures"])))
    stride = int(z["l_stride"])
    assert len(pars) == n and upstream.available(), "needs shim/_build/libclass_b200.so (python -c 'import __graft_entry__ as g; g.build()')"
    n_fixture = n
    if len(sys.argv) > 3:
        n = min(n, int(sys.argv[3]))
        pars = pars[:n]
    B = min(B, n)
    n_chunks = (n + B - 1) // B
    NSET = 2
    sets = []
    for _ in range(NSET):
        cs = []
        for _ in range(B):
            c = M.Context(0)
            c.set_option("lean_scratch", 1)
            cs.append(c)
        sets.append(cs)
    held = [[None] * B for _ in range(NSET)]
    err = np.full((n, len(SPECTRA)), np.nan)
    failures = {}
    host_s = np.zeros(n)
    lock = threading.Lock()

    def index(chunk, b):
        i = chunk * B + b
        return i if i < n else None

    def front(s, b, chunk):
        i = index(chunk, b)
        j = i if i is not None else n - 1  # a ragged last batch is padded with the last point (not counted twice)
        t0 = time.perf_counter()
        inp = upstream.inputs_for(pars[j])
        host_s[j] = time.perf_counter() - t0
        bg = M.BackgroundModule(inp, sets[s][b])
        th = M.ThermodynamicsModule(inp, bg)
        held[s][b] = (inp, bg, th, i)
        return M.PerturbationsModule(inp, bg, th, solve=False)

    def back(s, b, pt, chunk):
        inp, bg, th, i = held[s][b]
        if i is None:
            return None
        try:
            par = pars[i]
            prim = M.AnalyticPrimordial(par["A_s"], par["n_s"])
            nl = M.NonlinearModule(inp, bg, pt, prim, fetch=True)
            tr = M.TransferModule(inp, bg, th, pt, nl)
            sp = M.SpectraModule(inp, pt, prim, nl, tr)
            le = M.LensingModule(inp, sp)
            if i in ref_failures:
                raise RuntimeError("the reference rejected this point: " + ref_failures[i])
            if not np.array_equal(pt.k_[0], z["%d__k" % i]):
                raise RuntimeError("k grid differs from the reference's")
            e = cl_errors(sp, z["%d__cl" % i])
            lref = z["%d__cl_lensed" % i]
            ll = np.arange(2, le.l_lensed_max_ + 1)[::stride][: len(lref)]
            mine = np.array([le.lensing_cl_at_l(int(l)) for l in ll])
            lref = lref[: len(ll)]
            tt, ee, te, bb = le.index_lt_tt_, le.index_lt_ee_, le.index_lt_te_, le.index_lt_bb_
            e["lensed_tt"] = float(np.max(np.abs(mine[:, tt] / lref[:, tt] - 1)))
            e["lensed_ee"] = float(np.max(np.abs(mine[:, ee] / lref[:, ee] - 1)))
            e["lensed_te"] = float(np.max(np.abs(mine[:, te] - lref[:, te]) / np.sqrt(lref[:, tt] * lref[:, ee])))
            e["lensed_bb"] = float(np.max(np.abs(mine[:, bb] / lref[:, bb] - 1)))
            pk = pt.pk_linear(prim.pk_at_k(pt.k_[0]))
            e["pk_lin"] = float(np.max(np.abs(pk / z["%d__pk_lin_m" % i] - 1)))
            r_nl = nl.nl_corr_density_[0].reshape(pt.info.tau_size, pt.info.k_size)[-1]
            e["pk_nl"] = float(np.max(np.abs(pk * r_nl ** 2 / z["%d__pk_nl_m" % i] - 1)))
            with lock:
                err[i] = [e[k] for k in SPECTRA]
        except Exception as ex:  # recorded per cosmology; the sweep goes on
            with lock:
                failures[i] = repr(ex)[:300]
        return i

    pipe = SweepPipeline(B, front, back, n_sets=NSET)
    t0 = time.perf_counter()
    for chunk in range(n_chunks):
        try:
            pipe.submit(chunk)
        except Exception as ex:  # a failed batched launch surfaces at the next submit of its context set
            failures["batch before %d" % chunk] = repr(ex)[:300]
        print("config5: batch %d / %d submitted at %.0f s" % (chunk + 1, n_chunks, time.perf_counter() - t0), flush=True)
    try:
        pipe.close()
    except Exception as ex:
        failures["drain"] = repr(ex)[:300]
    wall = time.perf_counter() - t0

    done = ~np.isnan(err[:, 0])
    out = {
        "what": "BASELINE config 5: %d-point seed-0 Latin hypercube (omega_b, omega_cdm, h, ln10^10A_s, n_s, tau_reio), Planck-18 "
                "settings at full resolution, one B200, batches of %d; every cosmology against the unmodified reference%s"
                % (n_fixture, B, "" if n == n_fixture else " (this run: the first %d points only)" % n),
        "n": n, "compared": int(done.sum()), "failures": failures, "tolerance": TOL,
        "wall_s": wall, "spectra_per_s_including_upstream": n / wall,
        "perturb_launch_s": [round(x, 2) for x in pipe.solve_seconds],
        "upstream_host_s_per_cosmology_single_thread_mean": float(host_s.mean()),
        "reference": {"wall_s": float(z["wall_s"]), "processes": int(z["processes"]),
                      "seconds_per_cosmology_mean": float(np.mean([z["%d__seconds" % i] for i in range(n_fixture) if i not in ref_failures])),
                      "where": "the build container (not the GPU box's host): a cross-check of the fixture, not a baseline"},
        "max_rel_err": {k: float(np.nanmax(err[:, j])) for j, k in enumerate(SPECTRA)},
        "median_rel_err": {k: float(np.nanmedian(err[:, j])) for j, k in enumerate(SPECTRA)},
        "n_above_tolerance": {k: int(np.sum(err[done, j] > TOL)) for j, k in enumerate(SPECTRA)},
        "worst_point": {k: pars[int(np.nanargmax(err[:, j]))] for j, k in enumerate(SPECTRA)},
    }
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "config5_lhs%d%s.json" % (n_fixture, "" if n == n_fixture else "_first%d" % n))
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: out[k] for k in ("n", "compared", "failures", "wall_s", "spectra_per_s_including_upstream",
                                          "max_rel_err", "n_above_tolerance")}, indent=1))
    for cs in sets:
        for c in cs:
            c.close()


if __name__ == "__main__":
    main()
