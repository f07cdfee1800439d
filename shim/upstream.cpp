// Upstream inputs of the hot path, from the drop-in library itself (libclass_b200.so): the reference's InputModule /
// BackgroundModule / ThermodynamicsModule / NonColdDarkMatter -- out of scope of this repository and used unchanged
//  -- run for a parameter set given as "name = value" lines (the .ini surface), and their public results
// are handed out over a small C API.  classpp_public_b200/upstream.py turns them into the `Inputs` the batched sweep
// entry points (clpp_perturb_solve_batch ...) take: this is how a parameter sweep feeds many DIFFERENT cosmologies to one
// launch, which the reference's one-Cosmology-at-a-time constructors cannot express.
#include <cstring>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "background_module.h"
#include "cosmology.h"
#include "non_cold_dark_matter.h"
#include "thermodynamics_module.h"

namespace {
struct Upstream {
  FileContent fc;
  std::unique_ptr<Cosmology> cosmo;
};
void set_err(char* err, const std::string& s) {
  if (err) { std::strncpy(err, s.c_str(), 2047); err[2047] = 0; }
}
long put(const double* src, long n, double* out, long cap) {
  if (out) for (long i = 0; i < (n < cap ? n : cap); i++) out[i] = src[i];
  return n;
}
long put1(double v, double* out, long cap) {
  if (out && cap > 0) out[0] = v;
  return 1;
}
}  // namespace

extern "C" {

void* clpp_upstream_create(const char* params, char* err) {
  try {
    std::vector<std::pair<std::string, std::string>> kv;
    std::istringstream ss(params);
    std::string line;
    auto trim = [](const std::string& s) {
      const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
      return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    };
    while (std::getline(ss, line)) {
      const size_t p = line.find('=');
      if (p == std::string::npos) continue;
      const std::string n = trim(line.substr(0, p)), v = trim(line.substr(p + 1));
      if (!n.empty() && n[0] != '#') kv.emplace_back(n, v);
    }
    std::unique_ptr<Upstream> u(new Upstream());
    ErrorMsg msg;
    if (parser_init(&u->fc, (int)kv.size(), "clpp_upstream", msg) == _FAILURE_) { set_err(err, msg); return nullptr; }
    for (size_t i = 0; i < kv.size(); i++) {
      snprintf(u->fc.name[i], _ARGUMENT_LENGTH_MAX_, "%s", kv[i].first.c_str());
      snprintf(u->fc.value[i], _ARGUMENT_LENGTH_MAX_, "%s", kv[i].second.c_str());
      u->fc.read[i] = 0;
    }
    u->cosmo.reset(new Cosmology(u->fc));
    u->cosmo->GetThermodynamicsModule();  // background + thermodynamics: everything the hot path reads
    return u.release();
  } catch (std::exception& e) {
    set_err(err, e.what());
    return nullptr;
  }
}

void clpp_upstream_destroy(void* h) { delete static_cast<Upstream*>(h); }

// named scalars / arrays (as doubles): returns the length, copies at most cap values; -1: unknown name
long clpp_upstream_get(void* h, const char* cname, double* out, long cap) {
  Upstream* u = static_cast<Upstream*>(h);
  const std::string name(cname);
  Cosmology& c = *u->cosmo;
  const InputModule& in = *c.GetInputModule();
  const precision& pr = in.precision_;
  const background& ba = in.background_;
  const thermo& th = in.thermodynamics_;
  const perturbs& pt = in.perturbations_;
  const primordial& pm = in.primordial_;
#define S(n, v) if (name == n) return put1((double)(v), out, cap)
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
  PR(ur_fluid_approximation); PR(ur_fluid_trigger_tau_over_tau_k);
  PR(ncdm_fluid_approximation); PR(ncdm_fluid_trigger_tau_over_tau_k);
  PR(evolver); PR(curvature_ini);
  PR(l_logstep); PR(l_linstep); PR(hyper_x_min); PR(hyper_sampling_flat); PR(hyper_phi_min_abs);
  PR(q_linstep); PR(q_logstep_spline); PR(q_logstep_open);
  PR(transfer_neglect_delta_k_S_t0); PR(transfer_neglect_delta_k_S_t1); PR(transfer_neglect_delta_k_S_t2);
  PR(transfer_neglect_delta_k_S_e); PR(transfer_neglect_late_source); PR(l_switch_limber);
  PR(accurate_lensing); PR(delta_l_max); PR(num_mu_minus_lmax); PR(tol_gauss_legendre);
  PR(halofit_min_k_nonlinear); PR(halofit_k_per_decade); PR(halofit_sigma_precision); PR(halofit_tol_sigma);
#undef PR
  S("ba.h", ba.h); S("ba.H0", ba.H0); S("ba.K", ba.K); S("ba.sgnK", ba.sgnK); S("ba.a_today", ba.a_today);
  S("ba.T_cmb", ba.T_cmb); S("ba.Omega0_b", ba.Omega0_b);
  S("ba.has_cdm", ba.has_cdm); S("ba.has_ur", ba.has_ur); S("ba.has_ncdm", ba.has_ncdm); S("ba.has_fld", ba.has_fld);
  S("ba.has_curvature", ba.has_curvature); S("ba.has_dcdm", ba.has_dcdm); S("ba.has_dr", ba.has_dr); S("ba.has_scf", ba.has_scf);
  S("ba.has_idr", ba.has_idr); S("ba.has_idm_dr", ba.has_idm_dr); S("ba.N_ncdm", ba.N_ncdm);
  S("th.reio_parametrization", th.reio_parametrization); S("th.compute_cb2_derivatives", th.compute_cb2_derivatives);
  S("th.compute_damping_scale", th.compute_damping_scale);
  S("pt.gauge", pt.gauge); S("pt.l_scalar_max", pt.l_scalar_max); S("pt.k_max_for_pk", pt.k_max_for_pk); S("pt.z_max_pk", pt.z_max_pk);
  S("pt.has_cl_cmb_temperature", pt.has_cl_cmb_temperature); S("pt.has_cl_cmb_polarization", pt.has_cl_cmb_polarization);
  S("pt.has_cl_cmb_lensing_potential", pt.has_cl_cmb_lensing_potential); S("pt.has_pk_matter", pt.has_pk_matter);
  S("pt.has_nl_corrections_based_on_delta_m", pt.has_nl_corrections_based_on_delta_m);
  S("pt.switch_sw", pt.switch_sw); S("pt.switch_eisw", pt.switch_eisw); S("pt.switch_lisw", pt.switch_lisw);
  S("pt.switch_dop", pt.switch_dop); S("pt.switch_pol", pt.switch_pol); S("pt.eisw_lisw_split_z", pt.eisw_lisw_split_z);
  S("pt.three_ceff2_ur", pt.three_ceff2_ur); S("pt.three_cvis2_ur", pt.three_cvis2_ur); S("pt.G_eff_ur", pt.G_eff_ur);
  S("tr.lcmb_rescale", in.transfers_.lcmb_rescale); S("tr.lcmb_tilt", in.transfers_.lcmb_tilt); S("tr.lcmb_pivot", in.transfers_.lcmb_pivot);
  S("nl.method", in.nonlinear_.method);
  // analytic primordial spectrum P_R(k) = A_s (k/k_pivot)^(n_s - 1 + ...) (primordial.h): enough for AnalyticPrimordial
  S("pm.A_s", pm.A_s); S("pm.n_s", pm.n_s); S("pm.alpha_s", pm.alpha_s); S("pm.k_pivot", pm.k_pivot);
  S("pm.primordial_spec_type", pm.primordial_spec_type);
  if (name.rfind("bg.", 0) == 0) {
    const BackgroundModule& bg = *c.GetBackgroundModule();
    S("bg.bt_size", bg.bt_size_); S("bg.bg_size", bg.bg_size_); S("bg.bg_size_short", bg.bg_size_short_);
    S("bg.bg_size_normal", bg.bg_size_normal_); S("bg.conformal_age", bg.conformal_age_);
#define BI(x) S("bg.index_" #x, bg.index_bg_##x##_)
    BI(a); BI(H); BI(H_prime); BI(rho_g); BI(rho_b); BI(rho_cdm); BI(rho_ur); BI(rho_ncdm1); BI(p_ncdm1); BI(pseudo_p_ncdm1);
#undef BI
    if (name == "bg.tau_table") return put(bg.tau_table_, bg.bt_size_, out, cap);
    if (name == "bg.background_table") return put(bg.background_table_, (long)bg.bt_size_ * bg.bg_size_, out, cap);
    return -1;
  }
  if (name.rfind("ncdm.", 0) == 0) {
    const NonColdDarkMatter* nc = in.ncdm_.get();
    if (!nc || ba.N_ncdm == 0) return 0;
    std::vector<double> flat;
    for (int n = 0; n < ba.N_ncdm; n++) {
      if (name == "ncdm.q_size") flat.push_back(nc->q_size_ncdm_[n]);
      else if (name == "ncdm.M") flat.push_back(nc->M_ncdm_[n]);
      else if (name == "ncdm.factor") flat.push_back(nc->factor_ncdm_[n]);
      else
        for (int i = 0; i < nc->q_size_ncdm_[n]; i++) {
          if (name == "ncdm.q") flat.push_back(nc->q_ncdm_[n][i]);
          else if (name == "ncdm.w") flat.push_back(nc->w_ncdm_[n][i]);
          else if (name == "ncdm.dlnf0_dlnq") flat.push_back(nc->dlnf0_dlnq_ncdm_[n][i]);
          else return -1;
        }
    }
    return put(flat.data(), (long)flat.size(), out, cap);
  }
  if (name.rfind("th.", 0) == 0) {
    const ThermodynamicsModule& t = *c.GetThermodynamicsModule();
    S("th.tt_size", t.tt_size_); S("th.th_size", t.th_size_); S("th.tau_ini", t.tau_ini_); S("th.YHe", t.YHe_);
    S("th.tau_rec", t.tau_rec_); S("th.rs_rec", t.rs_rec_); S("th.angular_rescaling", t.angular_rescaling_);
    S("th.tau_free_streaming", t.tau_free_streaming_); S("th.tau_cut", t.tau_cut_); S("th.n_e", t.n_e_);
    S("th.z_reionization", t.z_reionization_);
#define TI(x) S("th.index_" #x, t.index_th_##x##_)
    TI(xe); TI(rate); TI(tau_d); TI(dkappa); TI(ddkappa); TI(dddkappa); TI(exp_m_kappa); TI(g); TI(dg); TI(ddg);
    TI(Tb); TI(wb); TI(cb2); TI(dcb2); TI(ddcb2); TI(r_d);
#undef TI
    if (name == "th.z_table") return put(t.z_table_, t.tt_size_, out, cap);
    if (name == "th.thermodynamics_table") return put(t.thermodynamics_table_, (long)t.tt_size_ * t.th_size_, out, cap);
    return -1;
  }
#undef S
  return -1;
}

}  // extern "C"
