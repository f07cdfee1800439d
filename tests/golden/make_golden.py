"""Generate the golden fixtures of tests/golden/ from the UNMODIFIED reference (oracle/_ref).

Run here (container with /root/reference, after `make -C oracle`):
    python tests/golden/make_golden.py
Each <config>.npz holds
  * the upstream inputs of the hot path (scalars in `meta`, background/thermodynamics tables,
    ncdm momentum grids, primordial spectrum on the transfer k grid, halofit correction), and
  * reference outputs: the k/tau/q/l grids (bit-exact targets), the full C_l table, the lensed
    C_l's, and sub-sampled columns of S(k,tau) and rows of Delta_l(q).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.refprobe import RefCosmology  # noqa: E402
from classpp_public_b200.configs import CONFIGS  # noqa: E402
from refutil import inputs_from_reference, reference_sources  # noqa: E402


def make(name, n_k_cols=9, n_l_rows=8):
    ref = RefCosmology(CONFIGS[name], threads=os.cpu_count()).compute("lensing")
    inp = inputs_from_reference(ref)
    nk, nt, ntp = ref.iscalar("pt.k_size"), ref.iscalar("pt.tau_size"), ref.iscalar("pt.tp_size")
    ntt, nl, nq = ref.iscalar("tr.tt_size"), ref.iscalar("tr.l_size"), ref.iscalar("tr.q_size")
    src = reference_sources(ref).reshape(ntp, nt, nk)
    k_cols = np.unique(np.linspace(0, nk - 1, n_k_cols).astype(int))
    tr = ref.get("tr.transfer").reshape(ntt, nl, nq)
    l_rows = np.unique(np.linspace(0, nl - 1, n_l_rows).astype(int))
    extra = {
        "ref.k": ref.get("pt.k"), "ref.tau": ref.get("pt.tau_sampling"),
        "ref.q": ref.get("tr.q"), "ref.kq": ref.get("tr.k"), "ref.l": ref.get("tr.l"),
        "ref.l_size_tt": ref.get("tr.l_size_tt"),
        "ref.sizes": np.array([nk, ref.iscalar("pt.k_size_cl"), ref.iscalar("pt.k_size_cmb"), nt, ntp, ntt, nl, nq,
                               ref.iscalar("sp.ct_size")], dtype=np.float64),
        "ref.tp_index": np.array([ref.iscalar("pt.index_tp_" + n) if ref.iscalar("pt.has_source_" + f) else -1
                                  for n, f in (("t0", "t"), ("t1", "t"), ("t2", "t"), ("p", "p"),
                                               ("delta_m", "delta_m"), ("delta_cb", "delta_cb"),
                                               ("phi_plus_psi", "phi_plus_psi"))], dtype=np.float64),
        "ref.k_cols": k_cols.astype(np.float64), "ref.sources_cols": src[:, :, k_cols],
        "ref.l_rows": l_rows.astype(np.float64), "ref.transfer_rows": tr[:, l_rows, :],
        "ref.cl": ref.get("sp.cl"),
        "ref.cl_lensed": ref.get("le.cl_lensed"),
        "pm.pk_at_transfer_k": ref.get("pm.pk_at_transfer_k"),
    }
    if int(inp.meta["nl.method"]) != 0:
        extra["nl.nl_corr_density_m"] = ref.get("nl.nl_corr_density_m").astype(np.float64)
    out = os.path.join(HERE, name + ".npz")
    inp.save(out, extra)
    print(name, "->", out, "%.2f MB" % (os.path.getsize(out) / 1e6))


def make_pk(names):
    """Matter power spectra of the reference on the perturbation k grid at z = 0 (small, one file for all configs):
    primordial P_R(k_i), linear P_m / P_cb, non-linear (halofit) P_m, sigma8."""
    out = {}
    for name in names:
        ref = RefCosmology(CONFIGS[name], threads=os.cpu_count()).compute("nonlinear")
        out[name + "__pm.pk_at_pt_k"] = ref.get("pm.pk_at_pt_k")
        for key in ("nl.pk_lin_m_at_pt_k", "nl.pk_lin_cb_at_pt_k", "nl.pk_nl_m_at_pt_k", "nl.sigma8_m"):
            v = ref.get(key)
            if v is not None and len(v):
                out[name + "__" + key] = v
        ref.close()
    path = os.path.join(HERE, "pk_z0.npz")
    np.savez_compressed(path, **out)
    print("pk ->", path, "%.2f MB" % (os.path.getsize(path) / 1e6))


def make_pk_z(z_list=(0.0, 0.5, 1.5)):
    """P_m(k, z) at z > 0 (Planck-18 settings with z_max_pk = 2: the late-time source table and its spline in ln tau)."""
    par = dict(CONFIGS["planck18"], z_max_pk=2.0)
    ref = RefCosmology(par, threads=os.cpu_count()).compute("nonlinear")
    out = {"k": ref.get("pt.k"), "z": np.array(z_list), "sigma8": ref.get("nl.sigma8_m")}
    for z in z_list:
        out["pk_lin_%g" % z] = ref.get("nl.pk_lin_m_at_z.%g" % z)
        out["pk_nl_%g" % z] = ref.get("nl.pk_nl_m_at_z.%g" % z)
    ref.close()
    path = os.path.join(HERE, "pk_z.npz")
    np.savez_compressed(path, **out)
    print("pk(z) ->", path, "%.2f MB" % (os.path.getsize(path) / 1e6))


def make_lhs(n):
    """BASELINE config 5 in miniature: the first n points of the seed-0 Latin hypercube (Planck-18 settings, full default
    grids) through the unmodified reference: C_l table, lensed C_l, linear and non-linear P_m(k, z=0) per cosmology."""
    import json
    from classpp_public_b200.upstream import latin_hypercube_sweep
    pars = latin_hypercube_sweep(n, CONFIGS["planck18"], seed=0)
    out = {"params": np.array(json.dumps(pars))}
    for i, par in enumerate(pars):
        ref = RefCosmology(par, threads=os.cpu_count()).compute("lensing")
        out["%d__l" % i] = ref.get("tr.l")
        out["%d__cl" % i] = ref.get("sp.cl")
        out["%d__cl_lensed" % i] = ref.get("le.cl_lensed")
        out["%d__k" % i] = ref.get("pt.k")
        out["%d__pk_lin_m" % i] = ref.get("nl.pk_lin_m_at_pt_k")
        out["%d__pk_nl_m" % i] = ref.get("nl.pk_nl_m_at_pt_k")
        out["%d__sizes" % i] = np.array([ref.iscalar("pt.k_size"), ref.iscalar("pt.tau_size"), ref.iscalar("tr.q_size"),
                                         ref.iscalar("sp.ct_size"), ref.iscalar("le.lt_size")], dtype=np.float64)
        ref.close()
        print("lhs", i, par["h"], par["omega_cdm"], flush=True)
    path = os.path.join(HERE, "lhs%d.npz" % n)
    np.savez_compressed(path, **out)
    print("lhs ->", path, "%.2f MB" % (os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "pkz":
        make_pk_z()
    elif args and args[0] == "lhs":
        make_lhs(int(args[1]) if len(args) > 1 else 16)
    elif args and args[0] == "pk":
        make_pk(args[1:] or ["lcdm_coarse", "lcdm", "planck18", "ncdm3_deg", "lcdm_dense"])
    else:
        for name in (args or ["lcdm_coarse", "lcdm", "planck18"]):
            make(name)
