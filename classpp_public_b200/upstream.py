"""Inputs of the hot path for ARBITRARY parameter sets, from the compiled drop-in library (shim/_build/libclass_b200.so):
the reference's own InputModule / BackgroundModule / ThermodynamicsModule (out of scope, used unchanged) are run for an
.ini-style dict and their public results become an `Inputs` object -- what the batched sweep entry points take.
(The golden fixtures of tests/golden hold the same quantities for a few fixed cosmologies.)"""
import ctypes
import os

import numpy as np

from .modules import Inputs

_HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(os.path.dirname(_HERE), "shim", "_build")
LIB_PATH = os.environ.get("CLPP_DROPIN_LIB") or os.path.join(BUILD, "libclass_b200.so")
_LIB = None

# scalars the three stages read from the reference's input structs and upstream modules
PR_KEYS = """k_min_tau0 k_max_tau0_over_l_max k_step_sub k_step_super k_step_transition k_step_super_reduction
k_per_decade_for_pk k_per_decade_for_bao k_bao_center k_bao_width start_small_k_at_tau_c_over_tau_h
start_large_k_at_tau_h_over_tau_k tight_coupling_trigger_tau_c_over_tau_h tight_coupling_trigger_tau_c_over_tau_k
start_sources_at_tau_c_over_tau_h tight_coupling_approximation l_max_g l_max_pol_g l_max_ur l_max_ncdm
tol_ncdm_initial_w tol_tau_approx tol_perturb_integration perturb_sampling_stepsize smallest_allowed_variation
radiation_streaming_approximation radiation_streaming_trigger_tau_over_tau_k ur_fluid_approximation
ur_fluid_trigger_tau_over_tau_k ncdm_fluid_approximation ncdm_fluid_trigger_tau_over_tau_k evolver curvature_ini
perturb_integration_stepsize
l_logstep l_linstep hyper_x_min hyper_sampling_flat hyper_phi_min_abs q_linstep q_logstep_spline q_logstep_open
transfer_neglect_delta_k_S_t0 transfer_neglect_delta_k_S_t1 transfer_neglect_delta_k_S_t2
transfer_neglect_delta_k_S_e transfer_neglect_late_source l_switch_limber
accurate_lensing delta_l_max num_mu_minus_lmax tol_gauss_legendre
halofit_min_k_nonlinear halofit_k_per_decade halofit_sigma_precision halofit_tol_sigma""".split()
BA_KEYS = """h H0 K sgnK a_today T_cmb Omega0_b has_cdm has_ur has_ncdm has_fld has_curvature has_dcdm has_dr has_scf
has_idr has_idm_dr N_ncdm""".split()
TH_IN_KEYS = "reio_parametrization compute_cb2_derivatives compute_damping_scale".split()
PT_KEYS = """gauge l_scalar_max k_max_for_pk z_max_pk has_cl_cmb_temperature has_cl_cmb_polarization
has_cl_cmb_lensing_potential has_pk_matter has_nl_corrections_based_on_delta_m switch_sw switch_eisw switch_lisw
switch_dop switch_pol eisw_lisw_split_z three_ceff2_ur three_cvis2_ur G_eff_ur""".split()
TR_KEYS = "lcmb_rescale lcmb_tilt lcmb_pivot".split()
BG_KEYS = """bt_size bg_size bg_size_short bg_size_normal conformal_age index_a index_H index_H_prime index_rho_g
index_rho_b index_rho_cdm index_rho_ur index_rho_ncdm1 index_p_ncdm1 index_pseudo_p_ncdm1""".split()
TH_KEYS = """tt_size th_size tau_ini YHe tau_rec rs_rec angular_rescaling tau_free_streaming tau_cut n_e
z_reionization index_xe index_rate index_tau_d index_dkappa index_ddkappa index_dddkappa index_exp_m_kappa index_g
index_dg index_ddg index_Tb index_wb index_cb2 index_dcb2 index_ddcb2 index_r_d""".split()
PM_KEYS = "A_s n_s alpha_s k_pivot primordial_spec_type".split()


def available():
    return os.path.exists(LIB_PATH)


def _lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise ImportError("drop-in library not found at %s: build it with `make -C shim` (needs the reference sources)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.clpp_upstream_create.restype = ctypes.c_void_p
        lib.clpp_upstream_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        lib.clpp_upstream_destroy.argtypes = [ctypes.c_void_p]
        lib.clpp_upstream_get.restype = ctypes.c_long
        lib.clpp_upstream_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_long]
        _LIB = lib
    return _LIB


def collect_inputs(get, scalar, extra_scalar_groups=()):
    """Build an `Inputs` from getter callables (shared with tests/refutil.py, which reads the same keys from the oracle)."""
    meta = {}
    groups = (("pr.", PR_KEYS), ("ba.", BA_KEYS), ("th.", TH_IN_KEYS), ("pt.", PT_KEYS), ("tr.", TR_KEYS), ("bg.", BG_KEYS),
              ("th.", TH_KEYS)) + tuple(extra_scalar_groups)
    for pre, keys in groups:
        for k in keys:
            meta[pre + k] = scalar(pre + k)
    meta["nl.method"] = scalar("nl.method")
    arrays = {"bg.tau_table": get("bg.tau_table"), "bg.background_table": get("bg.background_table"),
              "th.z_table": get("th.z_table"), "th.thermodynamics_table": get("th.thermodynamics_table")}
    if int(meta["ba.has_ncdm"]):
        for k in ("ncdm.q_size", "ncdm.q", "ncdm.w", "ncdm.dlnf0_dlnq", "ncdm.M", "ncdm.factor"):
            arrays[k] = get(k)
    return Inputs(meta, arrays)


def inputs_for(params, threads=1):
    """`Inputs` of the hot path for an .ini-style parameter dict (runs the reference's background + thermodynamics)."""
    lib = _lib()
    p = dict(params)
    p.setdefault("class_dir", BUILD)
    p.setdefault("sBBN file", os.path.join(BUILD, "bbn", "sBBN_2017.dat"))
    p.setdefault("threads", int(threads))
    text = "".join("%s = %s\n" % (k, v) for k, v in p.items())
    err = ctypes.create_string_buffer(2048)
    h = lib.clpp_upstream_create(text.encode(), err)
    if not h:
        raise ValueError("input error: " + err.value.decode(errors="replace"))

    def get(name):
        n = lib.clpp_upstream_get(h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        if n:
            lib.clpp_upstream_get(h, name.encode(), out.ctypes.data_as(ctypes.c_void_p), n)
        return out

    try:
        return collect_inputs(get, lambda k: float(get(k)[0]), extra_scalar_groups=(("pm.", PM_KEYS),))
    finally:
        lib.clpp_upstream_destroy(h)


def latin_hypercube_sweep(n, base, seed=0):
    """BASELINE config 5: n points of scipy's Latin hypercube (d = 6, given seed) over
    omega_b [0.020, 0.024], omega_cdm [0.10, 0.14], h [0.60, 0.75], ln(10^10 A_s) [2.9, 3.2], n_s [0.92, 1.00],
    tau_reio [0.03, 0.09], every other setting taken from `base`."""
    from scipy.stats import qmc
    lo = np.array([0.020, 0.10, 0.60, 2.9, 0.92, 0.03])
    hi = np.array([0.024, 0.14, 0.75, 3.2, 1.00, 0.09])
    u = qmc.LatinHypercube(d=6, seed=seed).random(n)
    out = []
    for row in lo + u * (hi - lo):
        p = dict(base)
        for k in ("H0", "h", "100*theta_s", "A_s", "ln10^{10}A_s", "sigma8", "omega_b", "omega_cdm", "n_s", "tau_reio", "z_reio"):
            p.pop(k, None)
        p.update({"omega_b": float(row[0]), "omega_cdm": float(row[1]), "h": float(row[2]),
                  "A_s": float(np.exp(row[3]) * 1e-10), "n_s": float(row[4]), "tau_reio": float(row[5])})
        out.append(p)
    return out
