"""Test helpers that talk to the oracle (oracle/_ref, the unmodified reference).
Only tests/, smoke() and bench.py's cpu_baseline/reference arm may import this."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from classpp_public_b200.modules import Inputs  # noqa: E402

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


def inputs_from_reference(ref, with_nonlinear=True):
    """Collect the upstream quantities (what the C++ drop-in reads from the reference's
    Input/Background/Thermodynamics/Primordial/Nonlinear modules) into an `Inputs`."""
    meta = {}
    for pre, keys in (("pr.", PR_KEYS), ("ba.", BA_KEYS), ("th.", TH_IN_KEYS), ("pt.", PT_KEYS), ("tr.", TR_KEYS),
                      ("bg.", BG_KEYS), ("th.", TH_KEYS)):
        for k in keys:
            meta[pre + k] = ref.scalar(pre + k)
    meta["nl.method"] = ref.scalar("nl.method")
    arrays = {
        "bg.tau_table": ref.get("bg.tau_table"),
        "bg.background_table": ref.get("bg.background_table"),
        "th.z_table": ref.get("th.z_table"),
        "th.thermodynamics_table": ref.get("th.thermodynamics_table"),
    }
    if int(meta["ba.has_ncdm"]):
        for k in ("ncdm.q_size", "ncdm.q", "ncdm.w", "ncdm.dlnf0_dlnq", "ncdm.M", "ncdm.factor"):
            arrays[k] = ref.get(k)
    return Inputs(meta, arrays)


def perturb_info_from_reference(ref):
    from classpp_public_b200 import _capi as capi
    info = capi.PerturbInfo()
    info.k_size = ref.iscalar("pt.k_size")
    info.k_size_cl = ref.iscalar("pt.k_size_cl")
    info.k_size_cmb = ref.iscalar("pt.k_size_cmb")
    info.tau_size = ref.iscalar("pt.tau_size")
    info.tp_size = ref.iscalar("pt.tp_size")
    info.ln_tau_size = ref.iscalar("pt.ln_tau_size")
    for n, flag in (("t0", "t"), ("t1", "t"), ("t2", "t"), ("p", "p"), ("delta_m", "delta_m"),
                    ("delta_cb", "delta_cb"), ("phi_plus_psi", "phi_plus_psi")):
        has = ref.iscalar("pt.has_source_" + flag)
        setattr(info, "index_tp_" + n, ref.iscalar("pt.index_tp_" + n) if has else -1)
    info.k_min = ref.scalar("pt.k_min")
    info.k_max = ref.scalar("pt.k_max")
    return info


def reference_sources(ref):
    ntp = ref.iscalar("pt.tp_size")
    return np.stack([ref.get("pt.sources.%d" % tp) for tp in range(ntp)])
