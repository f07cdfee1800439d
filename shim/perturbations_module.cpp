// Drop-in PerturbationsModule: the reference's class (source/perturbations_module.h:7-178, unchanged header) with its
// constructor re-implemented over the C ABI of libclpp.so.  Replaces source/perturbations_module.cpp of the reference:
// grids on the host (bit-exact), every k mode integrated on the GPU, sources handed back in the reference layout
// sources_[index_md][index_ic*tp_size+index_tp][index_tau*k_size+index_k].
//
// Scope of the device path (anything else throws std::runtime_error with the reason, like the reference does for
// inconsistent input): scalar modes, adiabatic initial conditions, synchronous gauge, flat space, species
// photons/baryons/cdm/ur/ncdm/Lambda, sources T0 T1 T2 P delta_m delta_cb phi+psi.
#include <map>
#include <mutex>
#include <stdexcept>
#include <vector>

#include "background_module.h"
#include "non_cold_dark_matter.h"
#include "perturbations_module.h"
#include "thermodynamics_module.h"

#include "clpp_shim.h"

namespace clpp_shim {
static std::mutex g_mutex;
static std::map<const PerturbationsModule*, clpp_ctx*> g_ctx;
void register_ctx(const PerturbationsModule* p, clpp_ctx* c) {
  std::lock_guard<std::mutex> lock(g_mutex);
  g_ctx[p] = c;
}
clpp_ctx* ctx_of(const PerturbationsModule* p) {
  std::lock_guard<std::mutex> lock(g_mutex);
  auto it = g_ctx.find(p);
  return it == g_ctx.end() ? nullptr : it->second;
}
void release_ctx(const PerturbationsModule* p) {
  clpp_ctx* c = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_ctx.find(p);
    if (it != g_ctx.end()) { c = it->second; g_ctx.erase(it); }
  }
  if (c) clpp_ctx_destroy(c);
}
}  // namespace clpp_shim

PerturbationsModule::PerturbationsModule(InputModulePtr input_module, BackgroundModulePtr background_module,
                                         ThermodynamicsModulePtr thermodynamics_module)
    : BaseModule(std::move(input_module)),
      background_module_(std::move(background_module)),
      thermodynamics_module_(std::move(thermodynamics_module)) {
  if (perturb_init() != _SUCCESS_) {
    clpp_shim::release_ctx(this);
    throw std::runtime_error(error_message_);
  }
}

PerturbationsModule::~PerturbationsModule() {
  perturb_free();
  clpp_shim::release_ctx(this);
}

int PerturbationsModule::perturb_init() {
  md_size_ = 0;
  ln_tau_size_ = 0;
  ln_tau_ = nullptr;
  tau_sampling_ = nullptr;
  index_k_output_values_ = nullptr;
  for (int f = 0; f < _MAX_NUMBER_OF_K_FILES_; f++) {
    scalar_perturbations_data_[f] = vector_perturbations_data_[f] = tensor_perturbations_data_[f] = nullptr;
    size_scalar_perturbation_data_[f] = size_vector_perturbation_data_[f] = size_tensor_perturbation_data_[f] = 0;
  }
  number_of_scalar_titles_ = number_of_vector_titles_ = number_of_tensor_titles_ = 0;
  scalar_titles_[0] = vector_titles_[0] = tensor_titles_[0] = '\0';
  if (ppt->has_perturbations == _FALSE_) {
    if (ppt->perturbations_verbose > 0) printf("No sources requested. Perturbation module skipped.\n");
    return _SUCCESS_;
  }
  if (ppt->perturbations_verbose > 0) printf("Computing sources (B200 path, libclpp %s)\n", clpp_version());

  // ---- what this path integrates
  CLPP_SHIM_TEST(ppt->has_scalars == _FALSE_ || ppt->has_vectors == _TRUE_ || ppt->has_tensors == _TRUE_, error_message_,
                 "the B200 path integrates scalar modes only (modes = s)");
  CLPP_SHIM_TEST(ppt->has_ad == _FALSE_ || ppt->has_bi == _TRUE_ || ppt->has_cdi == _TRUE_ || ppt->has_nid == _TRUE_ ||
                     ppt->has_niv == _TRUE_,
                 error_message_, "the B200 path integrates adiabatic initial conditions only (ic = ad)");
  CLPP_SHIM_TEST(ppt->has_perturbed_recombination == _TRUE_, error_message_,
                 "perturbed recombination is not supported by the B200 path");
  CLPP_SHIM_TEST(ppt->has_density_transfers == _TRUE_ || ppt->has_velocity_transfers == _TRUE_ ||
                     ppt->has_metricpotential_transfers == _TRUE_ || ppt->has_Nbody_gauge_transfers == _TRUE_,
                 error_message_, "outputs dTk / vTk (density, velocity, metric transfer functions) are not supported by the B200 path");
  CLPP_SHIM_TEST(ppt->has_cl_number_count == _TRUE_ || ppt->has_cl_lensing_potential == _TRUE_, error_message_,
                 "outputs nCl / sCl (number count, galaxy lensing C_l) are not supported by the B200 path");
  CLPP_SHIM_TEST(ppt->k_output_values_num > 0, error_message_, "k_output_values is not supported by the B200 path");

  // ---- upstream tables -> device context
  clpp_ctx* ctx = nullptr;
  CLPP_SHIM_CALL(clpp_ctx_create(clpp_shim::device(), &ctx, clpp_err_), error_message_);
  clpp_shim::register_ctx(this, ctx);
  const BackgroundModule& bg = *background_module_;
  const ThermodynamicsModule& th = *thermodynamics_module_;
  clpp_background_desc bd;
  memset(&bd, 0, sizeof(bd));
  bd.bt_size = bg.bt_size_; bd.bg_size = bg.bg_size_; bd.bg_size_short = bg.bg_size_short_; bd.bg_size_normal = bg.bg_size_normal_;
  bd.index_bg_a = bg.index_bg_a_; bd.index_bg_H = bg.index_bg_H_; bd.index_bg_H_prime = bg.index_bg_H_prime_;
  bd.index_bg_rho_g = bg.index_bg_rho_g_; bd.index_bg_rho_b = bg.index_bg_rho_b_;
  bd.index_bg_rho_cdm = pba->has_cdm ? bg.index_bg_rho_cdm_ : 0;
  bd.index_bg_rho_ur = pba->has_ur ? bg.index_bg_rho_ur_ : 0;
  bd.index_bg_rho_ncdm1 = pba->has_ncdm ? bg.index_bg_rho_ncdm1_ : 0;
  bd.index_bg_p_ncdm1 = pba->has_ncdm ? bg.index_bg_p_ncdm1_ : 0;
  bd.index_bg_pseudo_p_ncdm1 = pba->has_ncdm ? bg.index_bg_pseudo_p_ncdm1_ : 0;
  bd.has_cdm = pba->has_cdm; bd.has_ur = pba->has_ur; bd.has_ncdm = pba->has_ncdm; bd.N_ncdm = pba->N_ncdm; bd.sgnK = pba->sgnK;
  bd.has_fld = pba->has_fld; bd.has_scf = pba->has_scf; bd.has_dcdm = pba->has_dcdm; bd.has_dr = pba->has_dr;
  bd.has_idr = pba->has_idr; bd.has_idm_dr = pba->has_idm_dr; bd.has_curvature = pba->has_curvature;
  bd.conformal_age = bg.conformal_age_; bd.a_today = pba->a_today; bd.H0 = pba->H0; bd.K = pba->K; bd.h = pba->h;
  bd.Omega0_b = pba->Omega0_b; bd.T_cmb = pba->T_cmb;
  CLPP_SHIM_CALL(clpp_set_background(ctx, &bd, bg.tau_table_, bg.background_table_, clpp_err_), error_message_);
  if (pba->has_ncdm == _TRUE_) {
    const NonColdDarkMatter& nc = *ncdm_;
    std::vector<int> qs(pba->N_ncdm);
    std::vector<double> q, w, dl, M(pba->N_ncdm), fac(pba->N_ncdm);
    for (int n = 0; n < pba->N_ncdm; n++) {
      qs[n] = nc.q_size_ncdm_[n];
      M[n] = nc.M_ncdm_[n];
      fac[n] = nc.factor_ncdm_[n];
      for (int i = 0; i < qs[n]; i++) {
        q.push_back(nc.q_ncdm_[n][i]);
        w.push_back(nc.w_ncdm_[n][i]);
        dl.push_back(nc.dlnf0_dlnq_ncdm_[n][i]);
      }
    }
    CLPP_SHIM_CALL(clpp_set_ncdm(ctx, pba->N_ncdm, qs.data(), q.data(), w.data(), dl.data(), M.data(), fac.data(), clpp_err_),
                   error_message_);
  }
  clpp_thermo_desc td;
  memset(&td, 0, sizeof(td));
  td.tt_size = th.tt_size_; td.th_size = th.th_size_;
  td.index_th_xe = th.index_th_xe_; td.index_th_rate = th.index_th_rate_; td.index_th_tau_d = th.index_th_tau_d_;
  td.index_th_dkappa = th.index_th_dkappa_; td.index_th_ddkappa = th.index_th_ddkappa_; td.index_th_dddkappa = th.index_th_dddkappa_;
  td.index_th_exp_m_kappa = th.index_th_exp_m_kappa_; td.index_th_g = th.index_th_g_; td.index_th_dg = th.index_th_dg_;
  td.index_th_ddg = th.index_th_ddg_; td.index_th_Tb = th.index_th_Tb_; td.index_th_wb = th.index_th_wb_;
  td.index_th_cb2 = th.index_th_cb2_;
  td.compute_cb2_derivatives = pth->compute_cb2_derivatives; td.compute_damping_scale = pth->compute_damping_scale;
  td.index_th_dcb2 = pth->compute_cb2_derivatives ? th.index_th_dcb2_ : 0;
  td.index_th_ddcb2 = pth->compute_cb2_derivatives ? th.index_th_ddcb2_ : 0;
  td.index_th_r_d = pth->compute_damping_scale ? th.index_th_r_d_ : 0;
  td.reio_parametrization = pth->reio_parametrization;
  td.z_reionization = th.z_reionization_; td.YHe = th.YHe_; td.n_e = th.n_e_;
  td.tau_ini = th.tau_ini_; td.tau_rec = th.tau_rec_; td.rs_rec = th.rs_rec_; td.angular_rescaling = th.angular_rescaling_;
  td.tau_free_streaming = th.tau_free_streaming_; td.tau_cut = th.tau_cut_;
  CLPP_SHIM_CALL(clpp_set_thermo(ctx, &td, th.z_table_, th.thermodynamics_table_, clpp_err_), error_message_);

  // ---- grids (host, bit-exact) and indices
  clpp_perturb_desc pd;
  memset(&pd, 0, sizeof(pd));
  pd.has_cl_cmb_temperature = ppt->has_cl_cmb_temperature; pd.has_cl_cmb_polarization = ppt->has_cl_cmb_polarization;
  pd.has_cl_cmb_lensing_potential = ppt->has_cl_cmb_lensing_potential; pd.has_pk_matter = ppt->has_pk_matter;
  pd.has_nl_corrections_based_on_delta_m = ppt->has_nl_corrections_based_on_delta_m;
  pd.gauge = ppt->gauge; pd.l_scalar_max = ppt->l_scalar_max; pd.k_max_for_pk = ppt->k_max_for_pk; pd.z_max_pk = ppt->z_max_pk;
  pd.switch_sw = ppt->switch_sw; pd.switch_eisw = ppt->switch_eisw; pd.switch_lisw = ppt->switch_lisw;
  pd.switch_dop = ppt->switch_dop; pd.switch_pol = ppt->switch_pol; pd.eisw_lisw_split_z = ppt->eisw_lisw_split_z;
  pd.three_ceff2_ur = ppt->three_ceff2_ur; pd.three_cvis2_ur = ppt->three_cvis2_ur;
#define PR(x) pd.x = ppr->x
  PR(k_min_tau0); PR(k_max_tau0_over_l_max); PR(k_step_sub); PR(k_step_super); PR(k_step_transition);
  PR(k_step_super_reduction); PR(k_per_decade_for_pk); PR(k_per_decade_for_bao); PR(k_bao_center); PR(k_bao_width);
  PR(start_small_k_at_tau_c_over_tau_h); PR(start_large_k_at_tau_h_over_tau_k);
  PR(tight_coupling_trigger_tau_c_over_tau_h); PR(tight_coupling_trigger_tau_c_over_tau_k);
  PR(start_sources_at_tau_c_over_tau_h); PR(tight_coupling_approximation);
  PR(l_max_g); PR(l_max_pol_g); PR(l_max_ur); PR(l_max_ncdm);
  PR(tol_ncdm_initial_w); PR(tol_tau_approx); PR(tol_perturb_integration); PR(perturb_sampling_stepsize);
  PR(smallest_allowed_variation); PR(radiation_streaming_approximation); PR(radiation_streaming_trigger_tau_over_tau_k);
  PR(ur_fluid_approximation); PR(ur_fluid_trigger_tau_over_tau_k);
  PR(ncdm_fluid_approximation); PR(ncdm_fluid_trigger_tau_over_tau_k);
  PR(evolver); PR(curvature_ini); PR(perturb_integration_stepsize);
#undef PR
  clpp_perturb_info pi;
  CLPP_SHIM_CALL(clpp_perturb_grids(ctx, &pd, &pi, clpp_err_), error_message_);

  // modes, initial conditions, source types (perturb_indices_of_perturbs, perturbations_module.cpp:843-1235, scalar branch)
  index_md_scalars_ = 0;
  md_size_ = 1;
  index_ic_ad_ = 0;
  tp_size_ = (int*)malloc(sizeof(int));
  ic_size_ = (int*)malloc(sizeof(int));
  ic_size_[0] = 1;
  tp_size_[0] = pi.tp_size;
  has_cmb_ = (ppt->has_cl_cmb_temperature == _TRUE_ || ppt->has_cl_cmb_polarization == _TRUE_) ? _TRUE_ : _FALSE_;
  has_source_t_ = pi.index_tp_t0 >= 0 ? _TRUE_ : _FALSE_;
  has_source_p_ = pi.index_tp_p >= 0 ? _TRUE_ : _FALSE_;
  has_source_delta_m_ = pi.index_tp_delta_m >= 0 ? _TRUE_ : _FALSE_;
  has_source_delta_cb_ = pi.index_tp_delta_cb >= 0 ? _TRUE_ : _FALSE_;
  has_source_phi_plus_psi_ = pi.index_tp_phi_plus_psi >= 0 ? _TRUE_ : _FALSE_;
  has_lss_ = (has_source_delta_m_ || has_source_phi_plus_psi_) ? _TRUE_ : _FALSE_;
  has_source_delta_tot_ = has_source_delta_g_ = has_source_delta_b_ = has_source_delta_cdm_ = has_source_delta_dcdm_ = _FALSE_;
  has_source_delta_fld_ = has_source_delta_scf_ = has_source_delta_dr_ = has_source_delta_ur_ = has_source_delta_idr_ = _FALSE_;
  has_source_delta_idm_dr_ = has_source_delta_ncdm_ = has_source_theta_m_ = has_source_theta_cb_ = has_source_theta_tot_ = _FALSE_;
  has_source_theta_g_ = has_source_theta_b_ = has_source_theta_cdm_ = has_source_theta_dcdm_ = has_source_theta_fld_ = _FALSE_;
  has_source_theta_scf_ = has_source_theta_dr_ = has_source_theta_ur_ = has_source_theta_idr_ = has_source_theta_idm_dr_ = _FALSE_;
  has_source_theta_ncdm_ = has_source_phi_ = has_source_phi_prime_ = has_source_psi_ = has_source_h_ = has_source_h_prime_ = _FALSE_;
  has_source_eta_ = has_source_eta_prime_ = has_source_H_T_Nb_prime_ = has_source_k2gamma_Nb_ = _FALSE_;
  index_tp_t0_ = pi.index_tp_t0; index_tp_t1_ = pi.index_tp_t1; index_tp_t2_ = pi.index_tp_t2; index_tp_p_ = pi.index_tp_p;
  index_tp_delta_m_ = pi.index_tp_delta_m; index_tp_delta_cb_ = pi.index_tp_delta_cb;
  index_tp_phi_plus_psi_ = pi.index_tp_phi_plus_psi;

  const int nk = pi.k_size, nt = pi.tau_size, ntp = pi.tp_size;
  k_size_ = (int*)malloc(sizeof(int)); k_size_cl_ = (int*)malloc(sizeof(int)); k_size_cmb_ = (int*)malloc(sizeof(int));
  k_size_[0] = nk; k_size_cl_[0] = pi.k_size_cl; k_size_cmb_[0] = pi.k_size_cmb;
  k_min_ = pi.k_min; k_max_ = pi.k_max;
  k_ = (double**)malloc(sizeof(double*));
  k_[0] = (double*)malloc(nk * sizeof(double));
  clpp_perturb_get_k(ctx, k_[0]);
  tau_size_ = nt;
  tau_sampling_ = (double*)malloc(nt * sizeof(double));
  clpp_perturb_get_tau(ctx, tau_sampling_);

  // late-time table for the interpolation of sources in 0 < z < z_max_pk (perturb_timesampling_for_sources :1541-1592)
  if (ppt->z_max_pk == 0.) {
    ln_tau_size_ = 1;
  } else {
    double tau_lower;
    class_call(background_module_->background_tau_of_z(ppt->z_max_pk, &tau_lower), background_module_->error_message_,
               error_message_);
    CLPP_SHIM_TEST(tau_lower <= tau_sampling_[0], error_message_,
                   "you asked for zmax=%e, i.e. taumin=%e, smaller than or equal to the first possible value =%e; it should be "
                   "strictly bigger for a successfull interpolation", ppt->z_max_pk, tau_lower, tau_sampling_[0]);
    int first = 0;
    while (tau_sampling_[first] < tau_lower) first++;
    first = first - 1 - 4 > 0 ? first - 1 - 4 : 0;  // the sample before tau(z_max) and four more against edge effects
    ln_tau_size_ = nt - first;
    ln_tau_ = (double*)malloc(ln_tau_size_ * sizeof(double));
    for (int i = 0; i < ln_tau_size_; i++) ln_tau_[i] = log(tau_sampling_[first + i]);
  }

  // ---- the k loop of perturb_init (:668-717), on the GPU
  CLPP_SHIM_CALL(clpp_perturb_solve(ctx, 0, nk, clpp_err_), error_message_);
  sources_ = (double***)malloc(sizeof(double**));
  late_sources_ = (double***)malloc(sizeof(double**));
  ddlate_sources_ = (double***)malloc(sizeof(double**));
  sources_[0] = (double**)malloc(ntp * sizeof(double*));
  late_sources_[0] = (double**)malloc(ntp * sizeof(double*));
  ddlate_sources_[0] = (double**)malloc(ntp * sizeof(double*));
  {
    const size_t per = (size_t)nt * nk;
    std::vector<double> all(per * ntp);
    CLPP_SHIM_CALL(clpp_perturb_get_sources(ctx, all.data(), clpp_err_), error_message_);
    for (int tp = 0; tp < ntp; tp++) {
      sources_[0][tp] = (double*)malloc(per * sizeof(double));
      memcpy(sources_[0][tp], all.data() + per * tp, per * sizeof(double));
      late_sources_[0][tp] = nullptr;
      ddlate_sources_[0][tp] = nullptr;
      if (ln_tau_size_ > 1) {
        late_sources_[0][tp] = sources_[0][tp] + (size_t)(nt - ln_tau_size_) * nk;
        ddlate_sources_[0][tp] = (double*)malloc((size_t)nk * ln_tau_size_ * sizeof(double));
        class_call(array_spline_table_lines(ln_tau_, ln_tau_size_, late_sources_[0][tp], nk, ddlate_sources_[0][tp],
                                            _SPLINE_EST_DERIV_, error_message_),
                   error_message_, error_message_);
      }
    }
  }
  return _SUCCESS_;
}

int PerturbationsModule::perturb_free() {
  if (ppt->has_perturbations == _FALSE_ || md_size_ == 0) return _SUCCESS_;
  for (int tp = 0; tp < tp_size_[0]; tp++) {
    free(sources_[0][tp]);
    if (ln_tau_size_ > 1) free(ddlate_sources_[0][tp]);
  }
  free(sources_[0]); free(late_sources_[0]); free(ddlate_sources_[0]);
  free(sources_); free(late_sources_); free(ddlate_sources_);
  free(k_[0]); free(k_);
  free(tau_sampling_);
  if (ln_tau_size_ > 1) free(ln_tau_);
  free(tp_size_); free(ic_size_); free(k_size_); free(k_size_cl_); free(k_size_cmb_);
  md_size_ = 0;
  return _SUCCESS_;
}

// S^X(k, tau) for every k at an arbitrary time (perturbations_module.cpp:79-132): linear in tau on the full table when no
// late-time table exists, cubic spline in ln(tau) on the late-time table otherwise (same branches as the reference).
int PerturbationsModule::perturb_sources_at_tau(int index_md, int index_ic, int index_tp, double tau, double* psource) const {
  const int nk = k_size_[index_md];
  const int slot = index_ic * tp_size_[index_md] + index_tp;
  if (ln_tau_size_ > 1 && tau >= exp(ln_tau_[0]) ) {
    int last_index;
    class_call(array_interpolate_spline(ln_tau_, ln_tau_size_, late_sources_[index_md][slot], ddlate_sources_[index_md][slot], nk,
                                        log(tau), &last_index, psource, nk, error_message_),
               error_message_, error_message_);
  } else {
    class_call(array_interpolate_two_bis(tau_sampling_, 1, 0, sources_[index_md][slot], nk, tau_size_, tau, psource, nk,
                                         error_message_),
               error_message_, error_message_);
  }
  return _SUCCESS_;
}

// File output of the perturbations at k_output_values (perturbations_module.cpp:146-435): not on the device path; the
// constructor already refuses k_output_values, so there is never any data to return.
int PerturbationsModule::perturb_output_data(enum file_format, double, int, double*) const {
  snprintf(error_message_, sizeof(ErrorMsg), "perturb_output_data: k_output_values is not supported by the B200 path");
  return _FAILURE_;
}
int PerturbationsModule::perturb_output_titles(enum file_format, char titles[_MAXTITLESTRINGLENGTH_]) const {
  titles[0] = '\0';
  return _SUCCESS_;
}
int PerturbationsModule::perturb_output_firstline_and_ic_suffix(int index_ic, char first_line[_LINE_LENGTH_MAX_], FileName ic_suffix) const {
  first_line[0] = '\0';
  ic_suffix[0] = '\0';
  if (index_ic == index_ic_ad_) {
    strcpy(ic_suffix, "ad");
    strcpy(first_line, "for adiabatic (AD) mode (normalized to initial curvature=1) ");
  }
  return _SUCCESS_;
}
