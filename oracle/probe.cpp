// oracle/probe.cpp -- TEST INFRASTRUCTURE, not product code.
//
// C API over the UNMODIFIED reference (CLASS++ v2.9.0), linked against
// oracle/_ref/libclass_ref.so (built by oracle/Makefile from /root/reference).
// It constructs the reference's Cosmology object (source/cosmology.h:5-33) stage by
// stage and hands out the stage-level arrays that the parity tests compare with:
//   background/thermodynamics tables (inputs of our hot path),
//   PerturbationsModule::k_, tau_sampling_, sources_   (perturbations_module.h:139-165)
//   TransferModule::q_, l_, transfer_                  (transfer_module.h:38-54)
//   SpectraModule::cl_                                 (spectra_module.h, private -> -fno-access-control)
//   LensingModule::lensing_cl_at_l                     (lensing_module.h:17)
// and the wall time of each module constructor (the CPU baseline of bench.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may use it.
// Compiled with -fno-access-control so that private tables can be read for validation.
#include <chrono>
#include <cstring>
#include <string>
#include <vector>
#include <sstream>
#include <stdexcept>

#include "cosmology.h"
#include "background_module.h"
#include "thermodynamics_module.h"
#include "perturbations_module.h"
#include "primordial_module.h"
#include "nonlinear_module.h"
#include "transfer_module.h"
#include "spectra_module.h"
#include "lensing_module.h"
#include "non_cold_dark_matter.h"

namespace {

struct Probe {
  FileContent fc;
  std::unique_ptr<Cosmology> cosmo;
  double t_background = 0, t_thermo = 0, t_perturb = 0, t_primordial = 0, t_nonlinear = 0,
         t_transfer = 0, t_spectra = 0, t_lensing = 0;
};

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void set_err(char* err, const std::string& s) {
  if (err) { std::strncpy(err, s.c_str(), 2047); err[2047] = 0; }
}

long put(const double* src, long n, double* out, long cap) {
  if (out) { long m = n < cap ? n : cap; for (long i = 0; i < m; i++) out[i] = src[i]; }
  return n;
}
long puti(const int* src, long n, double* out, long cap) {
  if (out) { long m = n < cap ? n : cap; for (long i = 0; i < m; i++) out[i] = (double)src[i]; }
  return n;
}
long put1(double v, double* out, long cap) {
  if (out && cap > 0) out[0] = v;
  return 1;
}

}  // namespace

extern "C" {

// params: "name = value\n" lines (the .ini surface, input_module.cpp:56-181)
void* rp_create(const char* params, char* err) {
  try {
    std::vector<std::pair<std::string, std::string>> kv;
    std::istringstream ss(params);
    std::string line;
    while (std::getline(ss, line)) {
      auto p = line.find('=');
      if (p == std::string::npos) continue;
      auto trim = [](std::string s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
      };
      std::string n = trim(line.substr(0, p)), v = trim(line.substr(p + 1));
      if (n.empty() || n[0] == '#') continue;
      kv.emplace_back(n, v);
    }
    Probe* pr = new Probe();
    ErrorMsg errmsg;
    if (parser_init(&pr->fc, (int)kv.size(), "probe", errmsg) == _FAILURE_) {
      set_err(err, errmsg); delete pr; return nullptr;
    }
    for (size_t i = 0; i < kv.size(); i++) {
      snprintf(pr->fc.name[i], _ARGUMENT_LENGTH_MAX_, "%s", kv[i].first.c_str());
      snprintf(pr->fc.value[i], _ARGUMENT_LENGTH_MAX_, "%s", kv[i].second.c_str());
      pr->fc.read[i] = 0;
    }
    pr->cosmo.reset(new Cosmology(pr->fc));
    return pr;
  } catch (std::exception& e) {
    set_err(err, e.what());
    return nullptr;
  }
}

void rp_destroy(void* h) { delete static_cast<Probe*>(h); }

// level: 0 background, 1 thermo, 2 perturb, 3 primordial, 4 nonlinear, 5 transfer, 6 spectra, 7 lensing
int rp_compute(void* h, int level, char* err) {
  Probe* p = static_cast<Probe*>(h);
  try {
    double t0;
    t0 = now(); p->cosmo->GetBackgroundModule(); p->t_background += now() - t0; if (level < 1) return 0;
    t0 = now(); p->cosmo->GetThermodynamicsModule(); p->t_thermo += now() - t0; if (level < 2) return 0;
    t0 = now(); p->cosmo->GetPerturbationsModule(); p->t_perturb += now() - t0; if (level < 3) return 0;
    t0 = now(); p->cosmo->GetPrimordialModule(); p->t_primordial += now() - t0; if (level < 4) return 0;
    t0 = now(); p->cosmo->GetNonlinearModule(); p->t_nonlinear += now() - t0; if (level < 5) return 0;
    t0 = now(); p->cosmo->GetTransferModule(); p->t_transfer += now() - t0; if (level < 6) return 0;
    t0 = now(); p->cosmo->GetSpectraModule(); p->t_spectra += now() - t0; if (level < 7) return 0;
    t0 = now(); p->cosmo->GetLensingModule(); p->t_lensing += now() - t0;
    return 0;
  } catch (std::exception& e) {
    set_err(err, e.what());
    return 1;
  }
}

// Generic getter: copies the named quantity (as doubles) into out[0..cap) and returns its
// length; out == NULL just returns the length. Returns -1 for an unknown name.
long rp_get(void* h, const char* cname, double* out, long cap) {
  Probe* p = static_cast<Probe*>(h);
  std::string name(cname);
  Cosmology& c = *p->cosmo;
  const InputModule& in = *c.GetInputModule();
  const precision& pr = in.precision_;
  const background& ba = in.background_;
  const thermo& th = in.thermodynamics_;
  const perturbs& pt = in.perturbations_;
  const transfers& tr = in.transfers_;
  const nonlinear& nl = in.nonlinear_;

#define S(n, v) if (name == n) return put1((double)(v), out, cap)
  // ---- timings ----
  S("time.background", p->t_background); S("time.thermo", p->t_thermo); S("time.perturb", p->t_perturb);
  S("time.primordial", p->t_primordial); S("time.nonlinear", p->t_nonlinear);
  S("time.transfer", p->t_transfer); S("time.spectra", p->t_spectra); S("time.lensing", p->t_lensing);

  // ---- precision parameters used by the hot path (include/precisions.h) ----
#define PR(x) S("pr." #x, pr.x)
  PR(k_min_tau0); PR(k_max_tau0_over_l_max); PR(k_step_sub); PR(k_step_super); PR(k_step_transition);
  PR(k_step_super_reduction); PR(k_per_decade_for_pk); PR(k_per_decade_for_bao); PR(k_bao_center); PR(k_bao_width);
  PR(start_small_k_at_tau_c_over_tau_h); PR(start_large_k_at_tau_h_over_tau_k);
  PR(tight_coupling_trigger_tau_c_over_tau_h); PR(tight_coupling_trigger_tau_c_over_tau_k);
  PR(start_sources_at_tau_c_over_tau_h); PR(tight_coupling_approximation);
  PR(l_max_g); PR(l_max_pol_g); PR(l_max_ur); PR(l_max_ncdm);
  PR(tol_ncdm_initial_w); PR(tol_tau_approx); PR(tol_perturb_integration); PR(perturb_sampling_stepsize);
  PR(perturb_integration_stepsize); PR(smallest_allowed_variation);
  PR(radiation_streaming_approximation); PR(radiation_streaming_trigger_tau_over_tau_k);
  PR(radiation_streaming_trigger_tau_c_over_tau);
  PR(ur_fluid_approximation); PR(ur_fluid_trigger_tau_over_tau_k);
  PR(ncdm_fluid_approximation); PR(ncdm_fluid_trigger_tau_over_tau_k);
  PR(evolver); PR(curvature_ini); PR(entropy_ini);
  PR(l_logstep); PR(l_linstep); PR(hyper_x_min); PR(hyper_sampling_flat); PR(hyper_phi_min_abs);
  PR(hyper_flat_approximation_nu);
  PR(q_linstep); PR(q_logstep_spline); PR(q_logstep_open); PR(q_logstep_trapzd); PR(q_numstep_transition);
  PR(transfer_neglect_delta_k_S_t0); PR(transfer_neglect_delta_k_S_t1); PR(transfer_neglect_delta_k_S_t2);
  PR(transfer_neglect_delta_k_S_e); PR(transfer_neglect_late_source); PR(l_switch_limber);
  PR(accurate_lensing); PR(delta_l_max); PR(num_mu_minus_lmax); PR(tol_gauss_legendre);
  PR(halofit_min_k_nonlinear); PR(halofit_k_per_decade); PR(halofit_sigma_precision); PR(halofit_tol_sigma);
#undef PR
  // ---- background / thermo / perturbs input structs ----
  S("ba.h", ba.h); S("ba.H0", ba.H0); S("ba.K", ba.K); S("ba.sgnK", ba.sgnK); S("ba.a_today", ba.a_today);
  S("ba.T_cmb", ba.T_cmb); S("ba.Omega0_b", ba.Omega0_b);
  S("ba.has_cdm", ba.has_cdm); S("ba.has_ur", ba.has_ur); S("ba.has_ncdm", ba.has_ncdm);
  S("ba.has_lambda", ba.has_lambda); S("ba.has_fld", ba.has_fld); S("ba.has_curvature", ba.has_curvature);
  S("ba.has_dcdm", ba.has_dcdm); S("ba.has_dr", ba.has_dr); S("ba.has_scf", ba.has_scf);
  S("ba.has_idr", ba.has_idr); S("ba.has_idm_dr", ba.has_idm_dr);
  S("ba.N_ncdm", ba.N_ncdm); S("ba.number_of_threads", ba.number_of_threads);
  S("th.reio_parametrization", th.reio_parametrization); S("th.compute_cb2_derivatives", th.compute_cb2_derivatives);
  S("th.compute_damping_scale", th.compute_damping_scale);
  S("pt.gauge", pt.gauge); S("pt.l_scalar_max", pt.l_scalar_max); S("pt.l_lss_max", pt.l_lss_max);
  S("pt.k_max_for_pk", pt.k_max_for_pk); S("pt.z_max_pk", pt.z_max_pk);
  S("pt.has_cl_cmb_temperature", pt.has_cl_cmb_temperature); S("pt.has_cl_cmb_polarization", pt.has_cl_cmb_polarization);
  S("pt.has_cl_cmb_lensing_potential", pt.has_cl_cmb_lensing_potential); S("pt.has_pk_matter", pt.has_pk_matter);
  S("pt.has_nl_corrections_based_on_delta_m", pt.has_nl_corrections_based_on_delta_m);
  S("pt.has_density_transfers", pt.has_density_transfers); S("pt.has_velocity_transfers", pt.has_velocity_transfers);
  S("pt.has_scalars", pt.has_scalars); S("pt.has_tensors", pt.has_tensors); S("pt.has_ad", pt.has_ad);
  S("pt.has_cls", pt.has_cls); S("pt.has_perturbed_recombination", pt.has_perturbed_recombination);
  S("pt.switch_sw", pt.switch_sw); S("pt.switch_eisw", pt.switch_eisw); S("pt.switch_lisw", pt.switch_lisw);
  S("pt.switch_dop", pt.switch_dop); S("pt.switch_pol", pt.switch_pol); S("pt.eisw_lisw_split_z", pt.eisw_lisw_split_z);
  S("pt.three_ceff2_ur", pt.three_ceff2_ur); S("pt.three_cvis2_ur", pt.three_cvis2_ur); S("pt.G_eff_ur", pt.G_eff_ur);
  S("tr.lcmb_rescale", tr.lcmb_rescale); S("tr.lcmb_tilt", tr.lcmb_tilt); S("tr.lcmb_pivot", tr.lcmb_pivot);
  S("nl.method", nl.method);

  // ---- background module ----
  if (name.rfind("bg.", 0) == 0) {
    const BackgroundModule& bg = *c.GetBackgroundModule();
    S("bg.bt_size", bg.bt_size_); S("bg.bg_size", bg.bg_size_); S("bg.bg_size_short", bg.bg_size_short_);
    S("bg.bg_size_normal", bg.bg_size_normal_); S("bg.conformal_age", bg.conformal_age_);
#define BI(x) S("bg.index_" #x, bg.index_bg_##x##_)
    BI(a); BI(H); BI(H_prime); BI(rho_g); BI(rho_b); BI(rho_cdm); BI(rho_lambda); BI(rho_ur);
    BI(rho_ncdm1); BI(p_ncdm1); BI(pseudo_p_ncdm1); BI(rho_tot); BI(p_tot); BI(p_tot_prime); BI(Omega_r);
    BI(rho_crit); BI(Omega_m); BI(conf_distance); BI(D); BI(f);
#undef BI
    if (name == "bg.tau_table") return put(bg.tau_table_, bg.bt_size_, out, cap);
    if (name == "bg.background_table") return put(bg.background_table_, (long)bg.bt_size_ * bg.bg_size_, out, cap);
    if (name == "bg.d2background_dtau2_table")
      return put(bg.d2background_dtau2_table_, (long)bg.bt_size_ * bg.bg_size_, out, cap);
    return -1;
  }
  // ---- ncdm helper object (tools/non_cold_dark_matter.h:65-76) ----
  if (name.rfind("ncdm.", 0) == 0) {
    const NonColdDarkMatter* nc = in.ncdm_.get();
    if (!nc || ba.N_ncdm == 0) return 0;
    if (name == "ncdm.q_size") return puti(nc->q_size_ncdm_, ba.N_ncdm, out, cap);
    if (name == "ncdm.M") return put(nc->M_ncdm_, ba.N_ncdm, out, cap);
    if (name == "ncdm.factor") return put(nc->factor_ncdm_, ba.N_ncdm, out, cap);
    std::vector<double> flat;
    for (int n = 0; n < ba.N_ncdm; n++)
      for (int i = 0; i < nc->q_size_ncdm_[n]; i++) {
        if (name == "ncdm.q") flat.push_back(nc->q_ncdm_[n][i]);
        else if (name == "ncdm.w") flat.push_back(nc->w_ncdm_[n][i]);
        else if (name == "ncdm.dlnf0_dlnq") flat.push_back(nc->dlnf0_dlnq_ncdm_[n][i]);
        else return -1;
      }
    return put(flat.data(), (long)flat.size(), out, cap);
  }
  // ---- thermodynamics module ----
  if (name.rfind("th.", 0) == 0) {
    const ThermodynamicsModule& t = *c.GetThermodynamicsModule();
    S("th.tt_size", t.tt_size_); S("th.th_size", t.th_size_); S("th.tau_ini", t.tau_ini_); S("th.YHe", t.YHe_);
    S("th.z_rec", t.z_rec_); S("th.tau_rec", t.tau_rec_); S("th.rs_rec", t.rs_rec_);
    S("th.angular_rescaling", t.angular_rescaling_); S("th.tau_free_streaming", t.tau_free_streaming_);
    S("th.tau_cut", t.tau_cut_); S("th.n_e", t.n_e_); S("th.z_reionization", t.z_reionization_);
#define TI(x) S("th.index_" #x, t.index_th_##x##_)
    TI(xe); TI(rate); TI(tau_d); TI(dkappa); TI(ddkappa); TI(dddkappa); TI(exp_m_kappa); TI(g); TI(dg); TI(ddg);
    TI(Tb); TI(wb); TI(cb2); TI(dcb2); TI(ddcb2); TI(r_d);
#undef TI
    if (name == "th.z_table") return put(t.z_table_, t.tt_size_, out, cap);
    if (name == "th.thermodynamics_table") return put(t.thermodynamics_table_, (long)t.tt_size_ * t.th_size_, out, cap);
    if (name == "th.d2thermodynamics_dz2_table")
      return put(t.d2thermodynamics_dz2_table_, (long)t.tt_size_ * t.th_size_, out, cap);
    return -1;
  }
  // ---- perturbations module ----
  if (name.rfind("pt.", 0) == 0) {
    const PerturbationsModule& m = *c.GetPerturbationsModule();
    const int md = m.index_md_scalars_;
    S("pt.md_size", m.md_size_); S("pt.ic_size", m.ic_size_[md]); S("pt.tp_size", m.tp_size_[md]);
    S("pt.k_size", m.k_size_[md]); S("pt.k_size_cl", m.k_size_cl_[md]); S("pt.k_size_cmb", m.k_size_cmb_[md]);
    S("pt.tau_size", m.tau_size_); S("pt.ln_tau_size", m.ln_tau_size_); S("pt.k_min", m.k_min_); S("pt.k_max", m.k_max_);
#define PI(x) S("pt.index_tp_" #x, m.index_tp_##x##_)
    PI(t0); PI(t1); PI(t2); PI(p); PI(delta_m); PI(delta_cb); PI(phi_plus_psi);
#undef PI
    S("pt.has_source_t", m.has_source_t_); S("pt.has_source_p", m.has_source_p_);
    S("pt.has_source_delta_m", m.has_source_delta_m_); S("pt.has_source_delta_cb", m.has_source_delta_cb_);
    S("pt.has_source_phi_plus_psi", m.has_source_phi_plus_psi_);
    S("pt.has_source_theta_m", m.has_source_theta_m_); S("pt.has_source_delta_g", m.has_source_delta_g_);
    S("pt.has_source_phi", m.has_source_phi_); S("pt.has_source_psi", m.has_source_psi_);
    if (name == "pt.k") return put(m.k_[md], m.k_size_[md], out, cap);
    if (name == "pt.tau_sampling") return put(m.tau_sampling_, m.tau_size_, out, cap);
    if (name.rfind("pt.sources.", 0) == 0) {  // pt.sources.<index_tp> (adiabatic ic)
      int tp = std::stoi(name.substr(11));
      if (tp < 0 || tp >= m.tp_size_[md]) return -1;
      return put(m.sources_[md][tp], (long)m.tau_size_ * m.k_size_[md], out, cap);
    }
    return -1;
  }
  // ---- primordial spectrum on the transfer k grid (what SpectraModule asks for, spectra_module.cpp:996) ----
  if (name == "pm.pk_at_transfer_k") {
    const PrimordialModule& pm = *c.GetPrimordialModule();
    const TransferModule& t = *c.GetTransferModule();
    const int md = c.GetPerturbationsModule()->index_md_scalars_;
    std::vector<double> pk(t.q_size_);
    for (int i = 0; i < t.q_size_; i++) {
      double v[8];
      pm.primordial_spectrum_at_k(md, linear, t.k_[md][i], v);
      pk[i] = v[0];
    }
    return put(pk.data(), t.q_size_, out, cap);
  }
  // ---- primordial spectrum on the perturbation k grid, and the reference's linear P(k, z=0) on that grid ----
  if (name == "pm.pk_at_pt_k") {
    const PrimordialModule& pm = *c.GetPrimordialModule();
    const PerturbationsModule& m = *c.GetPerturbationsModule();
    const int md = m.index_md_scalars_;
    std::vector<double> pk(m.k_size_[md]);
    for (int i = 0; i < m.k_size_[md]; i++) {
      double v[8];
      pm.primordial_spectrum_at_k(md, linear, m.k_[md][i], v);
      pk[i] = v[0];
    }
    return put(pk.data(), m.k_size_[md], out, cap);
  }
  if (name == "nl.pk_nl_m_at_pt_k") {  // non-linear P_m(k, z=0) on the perturbation k grid (nonlinear_pk_at_z, pk_nonlinear)
    const NonlinearModule& n = *c.GetNonlinearModule();
    const PerturbationsModule& m = *c.GetPerturbationsModule();
    const int nk = m.k_size_[m.index_md_scalars_];
    if (!n.has_pk_m_ || nl.method == nl_none) return 0;
    std::vector<double> pk(n.k_size_);
    if (n.nonlinear_pk_at_z(linear, pk_nonlinear, 0., n.index_pk_m_, pk.data(), nullptr) != 0) return -1;
    return put(pk.data(), nk, out, cap);
  }
  if (name.rfind("nl.pk_lin_m_at_z.", 0) == 0 || name.rfind("nl.pk_nl_m_at_z.", 0) == 0) {  // P_m(k, z) on the perturbation k grid
    const bool nonlin = name.rfind("nl.pk_nl_m_at_z.", 0) == 0;
    const double z = std::stod(name.substr(nonlin ? 16 : 17));
    const NonlinearModule& n = *c.GetNonlinearModule();
    const PerturbationsModule& m = *c.GetPerturbationsModule();
    const int nk = m.k_size_[m.index_md_scalars_];
    if (!n.has_pk_m_ || (nonlin && nl.method == nl_none)) return 0;
    std::vector<double> pk(n.k_size_);
    if (n.nonlinear_pk_at_z(linear, nonlin ? pk_nonlinear : pk_linear, z, n.index_pk_m_, pk.data(), nullptr) != 0) return -1;
    return put(pk.data(), nk, out, cap);
  }
  if (name == "nl.sigma8_m") {
    const NonlinearModule& n = *c.GetNonlinearModule();
    if (!n.has_pk_m_) return 0;
    return put1(n.sigma8_[n.index_pk_m_], out, cap);
  }
  if (name == "nl.pk_lin_m_at_pt_k" || name == "nl.pk_lin_cb_at_pt_k") {
    const NonlinearModule& n = *c.GetNonlinearModule();
    const PerturbationsModule& m = *c.GetPerturbationsModule();
    const int md = m.index_md_scalars_;
    const int nk = m.k_size_[md];
    // (nonlinear_pks_at_kvec_and_zvec dereferences the cb table even without cb: use the per-spectrum accessor;
    //  the module's k grid starts with the perturbation grid, nonlinear_get_k_list)
    const bool cb = (name == "nl.pk_lin_cb_at_pt_k");
    if (cb && !n.has_pk_cb_) return 0;
    if (!cb && !n.has_pk_m_) return 0;
    std::vector<double> pk(n.k_size_);
    if (n.nonlinear_pk_at_z(linear, pk_linear, 0., cb ? n.index_pk_cb_ : n.index_pk_m_, pk.data(), nullptr) != 0) return -1;
    for (int i = 0; i < nk; i++)
      if (fabs(exp(n.ln_k_[i]) / m.k_[md][i] - 1.) > 1e-12) return -1;
    return put(pk.data(), nk, out, cap);
  }
  // ---- nonlinear corrections seen by the transfer stage (transfer_module.cpp:559-590) ----
  if (name.rfind("nl.", 0) == 0) {
    const NonlinearModule& n = *c.GetNonlinearModule();
    const PerturbationsModule& m = *c.GetPerturbationsModule();
    const int md = m.index_md_scalars_;
    S("nl.has_pk_m", n.has_pk_m_); S("nl.has_pk_cb", n.has_pk_cb_);
    S("nl.index_pk_m", n.index_pk_m_); S("nl.index_pk_cb", n.index_pk_cb_);
    if (name == "nl.nl_corr_density_m") {
      if (nl.method == nl_none) return 0;
      return put(n.nl_corr_density_[n.index_pk_m_], (long)m.tau_size_ * m.k_size_[md], out, cap);
    }
    return -1;
  }
  // ---- transfer module ----
  if (name.rfind("tr.", 0) == 0) {
    const TransferModule& t = *c.GetTransferModule();
    const int md = c.GetPerturbationsModule()->index_md_scalars_;
    S("tr.tt_size", t.tt_size_[md]); S("tr.l_size", t.l_size_[md]); S("tr.l_size_max", t.l_size_max_);
    S("tr.q_size", t.q_size_); S("tr.index_tt_t0", t.index_tt_t0_); S("tr.index_tt_t1", t.index_tt_t1_);
    S("tr.index_tt_t2", t.index_tt_t2_); S("tr.index_tt_e", t.index_tt_e_); S("tr.index_tt_lcmb", t.index_tt_lcmb_);
    if (name == "tr.l") return puti(t.l_, t.l_size_max_, out, cap);
    if (name == "tr.l_size_tt") return puti(t.l_size_tt_[md], t.tt_size_[md], out, cap);
    if (name == "tr.q") return put(t.q_, t.q_size_, out, cap);
    if (name == "tr.k") return put(t.k_[md], t.q_size_, out, cap);
    if (name == "tr.transfer")
      return put(t.transfer_[md], (long)t.tt_size_[md] * t.l_size_[md] * t.q_size_, out, cap);
    return -1;
  }
  // ---- spectra module ----
  if (name.rfind("sp.", 0) == 0) {
    const SpectraModule& s = *c.GetSpectraModule();
    const int md = c.GetPerturbationsModule()->index_md_scalars_;
    S("sp.ct_size", s.ct_size_); S("sp.l_size", s.l_size_[md]); S("sp.l_max_tot", s.l_max_tot_);
    S("sp.index_ct_tt", s.index_ct_tt_); S("sp.index_ct_ee", s.index_ct_ee_); S("sp.index_ct_te", s.index_ct_te_);
    S("sp.index_ct_bb", s.index_ct_bb_); S("sp.index_ct_pp", s.index_ct_pp_); S("sp.index_ct_tp", s.index_ct_tp_);
    S("sp.index_ct_ep", s.index_ct_ep_);
    S("sp.has_tt", s.has_tt_); S("sp.has_ee", s.has_ee_); S("sp.has_te", s.has_te_); S("sp.has_bb", s.has_bb_);
    S("sp.has_pp", s.has_pp_); S("sp.has_tp", s.has_tp_); S("sp.has_ep", s.has_ep_);
    if (name == "sp.l") return put(s.l_, s.l_size_[md], out, cap);
    if (name == "sp.l_max_ct") return puti(s.l_max_ct_[md], s.ct_size_, out, cap);
    if (name == "sp.cl") return put(s.cl_[md], (long)s.l_size_[md] * s.ic_ic_size_[md] * s.ct_size_, out, cap);
    if (name == "sp.ddcl") return put(s.ddcl_[md], (long)s.l_size_[md] * s.ic_ic_size_[md] * s.ct_size_, out, cap);
    if (name.rfind("sp.cl_at_l.", 0) == 0) {  // sp.cl_at_l.<lmax>: [l=0..lmax][ct]  (spectra_cl_at_l, :220)
      int lmax = std::stoi(name.substr(11));
      std::vector<double> res((long)(lmax + 1) * s.ct_size_, 0.);
      std::vector<double> tmp(s.ct_size_);
      std::vector<double*> cl_md(s.md_size_, nullptr), cl_md_ic(s.md_size_, nullptr);
      std::vector<std::vector<double>> b1(s.md_size_, std::vector<double>(s.ct_size_)),
          b2(s.md_size_, std::vector<double>(s.ct_size_ * 8));
      for (int i = 0; i < s.md_size_; i++) { cl_md[i] = b1[i].data(); cl_md_ic[i] = b2[i].data(); }
      for (int l = 2; l <= lmax; l++) {
        s.spectra_cl_at_l((double)l, tmp.data(), cl_md.data(), cl_md_ic.data());
        for (int ct = 0; ct < s.ct_size_; ct++) res[(long)l * s.ct_size_ + ct] = tmp[ct];
      }
      return put(res.data(), (long)res.size(), out, cap);
    }
    return -1;
  }
  // ---- lensing module ----
  if (name.rfind("le.", 0) == 0) {
    const LensingModule& le = *c.GetLensingModule();
    S("le.l_lensed_max", le.l_lensed_max_); S("le.lt_size", le.lt_size_);
    S("le.index_lt_tt", le.index_lt_tt_); S("le.index_lt_ee", le.index_lt_ee_); S("le.index_lt_te", le.index_lt_te_);
    S("le.index_lt_bb", le.index_lt_bb_); S("le.index_lt_pp", le.index_lt_pp_); S("le.index_lt_tp", le.index_lt_tp_);
    if (name == "le.cl_lensed") {  // [l=0..l_lensed_max][lt]
      long n = (long)(le.l_lensed_max_ + 1) * le.lt_size_;
      std::vector<double> res(n, 0.);
      for (int l = 2; l <= le.l_lensed_max_; l++) le.lensing_cl_at_l(l, res.data() + (long)l * le.lt_size_);
      return put(res.data(), n, out, cap);
    }
    return -1;
  }
#undef S
  return -1;
}

}  // extern "C"
