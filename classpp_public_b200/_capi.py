"""ctypes binding of include/clpp.h (libclpp.so).  Plumbing only: the numerics are the CUDA
kernels behind the C ABI.  If the library is missing the import fails loudly -- there is no
Python/CPU fallback of the hot path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLPP_LIB") or os.path.join(_HERE, "libclpp.so")  # CLPP_LIB: developer override (build variants)
ERRLEN = 2048


def _fields(spec):
    out = []
    for line in spec.strip().splitlines():
        typ, names = line.split(":")
        ctype = {"int": C.c_int, "double": C.c_double, "long": C.c_long}[typ.strip()]
        for n in names.split(","):
            n = n.strip()
            if n:
                out.append((n, ctype))
    return out


class BackgroundDesc(C.Structure):
    _fields_ = _fields("""
    int: bt_size, bg_size, bg_size_short, bg_size_normal
    int: index_bg_a, index_bg_H, index_bg_H_prime
    int: index_bg_rho_g, index_bg_rho_b, index_bg_rho_cdm, index_bg_rho_ur
    int: index_bg_rho_ncdm1, index_bg_p_ncdm1, index_bg_pseudo_p_ncdm1
    int: has_cdm, has_ur, has_ncdm, N_ncdm, sgnK
    int: has_fld, has_scf, has_dcdm, has_dr, has_idr, has_idm_dr, has_curvature
    double: conformal_age, a_today, H0, K, h, Omega0_b, T_cmb
    """)


class ThermoDesc(C.Structure):
    _fields_ = _fields("""
    int: tt_size, th_size
    int: index_th_xe, index_th_rate, index_th_tau_d, index_th_dkappa, index_th_ddkappa, index_th_dddkappa
    int: index_th_exp_m_kappa, index_th_g, index_th_dg, index_th_ddg
    int: index_th_Tb, index_th_wb, index_th_cb2, index_th_dcb2, index_th_ddcb2, index_th_r_d
    int: compute_cb2_derivatives, compute_damping_scale
    int: reio_parametrization
    double: z_reionization, YHe, n_e
    double: tau_ini, tau_rec, rs_rec, angular_rescaling, tau_free_streaming, tau_cut
    """)


class PerturbDesc(C.Structure):
    _fields_ = _fields("""
    int: has_cl_cmb_temperature, has_cl_cmb_polarization, has_cl_cmb_lensing_potential
    int: has_pk_matter, has_nl_corrections_based_on_delta_m
    int: gauge
    int: l_scalar_max
    double: k_max_for_pk, z_max_pk
    int: switch_sw, switch_eisw, switch_lisw, switch_dop, switch_pol
    double: eisw_lisw_split_z
    double: three_ceff2_ur, three_cvis2_ur
    double: k_min_tau0, k_max_tau0_over_l_max, k_step_sub, k_step_super, k_step_transition
    double: k_step_super_reduction, k_per_decade_for_pk, k_per_decade_for_bao, k_bao_center, k_bao_width
    double: start_small_k_at_tau_c_over_tau_h, start_large_k_at_tau_h_over_tau_k
    double: tight_coupling_trigger_tau_c_over_tau_h, tight_coupling_trigger_tau_c_over_tau_k
    double: start_sources_at_tau_c_over_tau_h
    int: tight_coupling_approximation
    int: l_max_g, l_max_pol_g, l_max_ur, l_max_ncdm
    double: tol_ncdm_initial_w, tol_tau_approx, tol_perturb_integration, perturb_sampling_stepsize
    double: smallest_allowed_variation
    int: radiation_streaming_approximation
    double: radiation_streaming_trigger_tau_over_tau_k
    int: ur_fluid_approximation
    double: ur_fluid_trigger_tau_over_tau_k
    int: ncdm_fluid_approximation
    double: ncdm_fluid_trigger_tau_over_tau_k
    int: evolver
    double: curvature_ini
    double: perturb_integration_stepsize
    """)


class PerturbInfo(C.Structure):
    _fields_ = _fields("""
    int: k_size, k_size_cl, k_size_cmb, tau_size, tp_size, ln_tau_size
    int: index_tp_t0, index_tp_t1, index_tp_t2, index_tp_p, index_tp_delta_m, index_tp_delta_cb, index_tp_phi_plus_psi
    double: k_min, k_max
    """)


class KStat(C.Structure):
    _fields_ = _fields("""
    int: steps, failed, fevals, jacobians, factorizations, solves
    int: intervals, status
    double: tau_ini
    """) + [("iv_neq", C.c_int * 6), ("iv_steps", C.c_int * 6), ("iv_cycles", C.c_longlong * 6), ("prof", C.c_longlong * 72)]


class TransferDesc(C.Structure):
    _fields_ = _fields("""
    int: has_cl_cmb_temperature, has_cl_cmb_polarization, has_cl_cmb_lensing_potential
    int: l_scalar_max
    double: l_logstep, l_linstep
    double: hyper_x_min, hyper_sampling_flat, hyper_phi_min_abs
    double: q_linstep, q_logstep_spline, q_logstep_open
    double: transfer_neglect_delta_k_S_t0, transfer_neglect_delta_k_S_t1, transfer_neglect_delta_k_S_t2, transfer_neglect_delta_k_S_e
    double: transfer_neglect_late_source
    double: l_switch_limber
    double: lcmb_rescale, lcmb_tilt, lcmb_pivot
    """)


class TransferInfo(C.Structure):
    _fields_ = _fields("""
    int: tt_size, l_size, l_size_max, q_size
    int: index_tt_t0, index_tt_t1, index_tt_t2, index_tt_e, index_tt_lcmb
    int: x_size
    long: n_integrals, n_points
    """)


class HalofitDesc(C.Structure):
    _fields_ = _fields("""
    double: halofit_min_k_nonlinear, halofit_k_per_decade, halofit_sigma_precision, halofit_tol_sigma
    """)


class LensingDesc(C.Structure):
    _fields_ = _fields("""
    int: accurate_lensing, delta_l_max, num_mu_minus_lmax
    double: tol_gauss_legendre
    """)


class LensingInfo(C.Structure):
    _fields_ = _fields("""
    int: lt_size, l_size, l_unlensed_max, l_lensed_max
    int: index_lt_tt, index_lt_ee, index_lt_te, index_lt_bb, index_lt_pp, index_lt_tp, index_lt_ep
    """)


class SpectraInfo(C.Structure):
    _fields_ = _fields("""
    int: ct_size, l_size
    int: index_ct_tt, index_ct_ee, index_ct_te, index_ct_bb, index_ct_pp, index_ct_tp, index_ct_ep
    """)


# every symbol declared in include/clpp.h (tests check that the library exports all of them)
SYMBOLS = [
    "clpp_ctx_create", "clpp_ctx_destroy", "clpp_ctx_launch_count", "clpp_version",
    "clpp_ctx_get_stream", "clpp_ctx_get_kernel_ms", "clpp_measure_fp64_peak", "clpp_ctx_set_option",
    "clpp_set_background", "clpp_set_thermo", "clpp_set_ncdm",
    "clpp_perturb_grids", "clpp_perturb_solve", "clpp_perturb_solve_batch", "clpp_perturb_solve_list", "clpp_perturb_get_k", "clpp_perturb_get_tau",
    "clpp_perturb_get_sources", "clpp_perturb_get_kstat", "clpp_perturb_set_sources",
    "clpp_perturb_device_sources",
    "clpp_transfer_grids", "clpp_transfer_compute", "clpp_transfer_get_l", "clpp_transfer_get_q",
    "clpp_transfer_get_transfer", "clpp_transfer_set_transfer", "clpp_transfer_device_transfer",
    "clpp_transfer_get_bessel",
    "clpp_spectra_compute", "clpp_spectra_compute_range", "clpp_spectra_cl_at_l", "clpp_spectra_cl_output", "clpp_pk_linear", "clpp_nonlinear_halofit",
    "clpp_lensing_compute", "clpp_lensing_cl_at_l", "clpp_perturb_sources_at_tau",
]

_lib = None


def lib():
    """Load libclpp.so (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libclpp.so not found at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the B200 hot path has no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        P = C.POINTER
        dp, ip, cp, vp = P(C.c_double), P(C.c_int), C.c_char_p, C.c_void_p
        L.clpp_version.restype = C.c_char_p
        L.clpp_ctx_create.argtypes = [C.c_int, P(vp), cp]
        L.clpp_ctx_set_option.argtypes = [vp, cp, C.c_double, cp]
        L.clpp_ctx_destroy.argtypes = [vp]
        L.clpp_ctx_destroy.restype = None
        L.clpp_ctx_launch_count.argtypes = [vp]
        L.clpp_ctx_launch_count.restype = C.c_long
        L.clpp_ctx_get_stream.argtypes = [vp, P(vp)]
        L.clpp_ctx_get_kernel_ms.argtypes = [vp, dp]
        L.clpp_measure_fp64_peak.argtypes = [vp, dp, cp]
        L.clpp_set_background.argtypes = [vp, P(BackgroundDesc), dp, dp, cp]
        L.clpp_set_thermo.argtypes = [vp, P(ThermoDesc), dp, dp, cp]
        L.clpp_set_ncdm.argtypes = [vp, C.c_int, ip, dp, dp, dp, dp, dp, cp]
        L.clpp_perturb_grids.argtypes = [vp, P(PerturbDesc), P(PerturbInfo), cp]
        L.clpp_perturb_solve.argtypes = [vp, C.c_int, C.c_int, cp]
        L.clpp_perturb_solve_batch.argtypes = [P(vp), C.c_int, cp]
        L.clpp_perturb_solve_list.argtypes = [vp, ip, C.c_int, cp]
        L.clpp_perturb_get_k.argtypes = [vp, dp]
        L.clpp_perturb_get_tau.argtypes = [vp, dp]
        L.clpp_perturb_get_sources.argtypes = [vp, dp, cp]
        L.clpp_perturb_get_kstat.argtypes = [vp, P(KStat)]
        L.clpp_perturb_set_sources.argtypes = [vp, P(PerturbInfo), dp, dp, dp, cp]
        L.clpp_perturb_device_sources.argtypes = [vp, P(vp), P(C.c_long), cp]
        L.clpp_transfer_grids.argtypes = [vp, P(TransferDesc), P(TransferInfo), cp]
        L.clpp_transfer_compute.argtypes = [vp, dp, C.c_int, C.c_int, cp]
        L.clpp_transfer_get_l.argtypes = [vp, ip, ip]
        L.clpp_transfer_get_q.argtypes = [vp, dp, dp]
        L.clpp_transfer_get_transfer.argtypes = [vp, dp, cp]
        L.clpp_transfer_set_transfer.argtypes = [vp, dp, cp]
        L.clpp_transfer_device_transfer.argtypes = [vp, P(vp), P(C.c_long), cp]
        L.clpp_transfer_get_bessel.argtypes = [vp, dp, dp, dp, dp, cp]
        L.clpp_spectra_compute.argtypes = [vp, dp, P(SpectraInfo), dp, cp]
        L.clpp_spectra_compute_range.argtypes = [vp, dp, C.c_int, C.c_int, P(SpectraInfo), dp, cp]
        L.clpp_pk_linear.argtypes = [vp, dp, C.c_int, C.c_int, dp, cp]
        L.clpp_nonlinear_halofit.argtypes = [vp, P(HalofitDesc), dp, dp, ip, cp]
        L.clpp_perturb_sources_at_tau.argtypes = [vp, C.c_int, C.c_double, dp, cp]
        L.clpp_lensing_compute.argtypes = [vp, P(LensingDesc), P(LensingInfo), dp, dp, cp]
        L.clpp_lensing_cl_at_l.argtypes = [vp, C.c_int, dp, cp]
        L.clpp_spectra_cl_at_l.argtypes = [vp, C.c_double, dp, cp]
        L.clpp_spectra_cl_output.argtypes = [vp, C.c_int, dp, cp]
        _lib = L
    return _lib


def dptr(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int))
