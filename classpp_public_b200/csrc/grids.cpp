// Host-side sampling grids of the path. These must be BIT-EXACT with the reference, so they
// are evaluated on the host with the same libm and the same floating-point expressions.
//
//   clpp_host_perturb_grids:
//     source-type indices   <-> perturb_indices_of_perturbs   (perturbations_module.cpp:843-1235)
//     k grid                <-> perturb_get_k_list            (perturbations_module.cpp:1628-2238)
//     tau grid              <-> perturb_timesampling_for_sources (perturbations_module.cpp:1247-1619)
//   clpp_host_transfer_grids:
//     l grid                <-> transfer_get_l_list           (transfer_module.cpp:694-871)
//     q grid                <-> transfer_get_q_list           (transfer_module.cpp:884-1096)
//     k(q)                  <-> transfer_get_k_list           (transfer_module.cpp:1106-1167)
//
// Scope of the B200 path: scalar modes, adiabatic initial conditions, synchronous gauge,
// flat space (K = 0), CMB temperature/polarisation/lensing-potential + matter sources.
// Anything else is rejected loudly (there is no fallback).
#include <algorithm>
#include <cmath>

#include "clpp_internal.h"

static int k_list(clpp_ctx* c, char* err) {
  const clpp_perturb_desc& p = c->pd;
  const clpp_background_desc& bg = c->bg;
  const clpp_thermo_desc& th = c->th;

  CLPP_CHECK(p.k_step_transition != 0., err, "stop to avoid division by zero");
  CLPP_CHECK(th.rs_rec != 0., err, "stop to avoid division by zero");

  const double tau0 = bg.conformal_age;
  const double k_min = p.k_min_tau0 / tau0;
  const double k_rec = 2. * CLPP_PI / th.rs_rec;  // wavenumber of the sound horizon at recombination
  double k_max_cmb = k_min, k_max_cl = k_min, k_max = k_min;
  const bool has_cls = p.has_cl_cmb_temperature || p.has_cl_cmb_polarization || p.has_cl_cmb_lensing_potential;
  if (has_cls) {
    k_max_cmb = p.k_max_tau0_over_l_max * p.l_scalar_max / tau0 / th.angular_rescaling;
    k_max_cl = k_max_cmb;
    k_max = k_max_cmb;
  }
  if (p.has_pk_matter || p.has_nl_corrections_based_on_delta_m) k_max = std::max(k_max, p.k_max_for_pk);
  CLPP_CHECK(k_min >= 0., err, "buggy definition of k_min");
  CLPP_CHECK(k_max >= k_min, err, "buggy definition of k_min and/or k_max");

  std::vector<double>& K = c->k;
  K.clear();
  double k = k_min;
  K.push_back(k);
  // (1) CMB range: linear steps, refined above the sound-horizon scale and towards k -> 0
  while (k < k_max_cmb) {
    double step = (p.k_step_super +
                   0.5 * (tanh((k - k_rec) / k_rec / p.k_step_transition) + 1.) * (p.k_step_sub - p.k_step_super)) *
                  k_rec;
    const double scale2 = pow(bg.a_today * bg.H0, 2) + fabs(bg.K);
    step *= (k * k / scale2 + 1.) / (k * k / scale2 + 1. / p.k_step_super_reduction);
    CLPP_CHECK(step / k >= p.smallest_allowed_variation, err,
               "k step =%e < machine precision : leads either to numerical error or infinite loop", step * k_rec);
    k += step;
    CLPP_CHECK(k > K.back(), err, "consecutive values of k should differ and should be in growing order");
    K.push_back(k);
  }
  c->pinfo.k_size_cmb = (int)K.size();
  // (2)+(3) logarithmic steps with a denser band around the BAO scale
  auto log_step = [&](double kk) {
    return kk * pow(10., 1. / (p.k_per_decade_for_pk +
                               (p.k_per_decade_for_bao - p.k_per_decade_for_pk) *
                                   (1. - tanh(pow((log(kk) - log(p.k_bao_center * k_rec)) / log(p.k_bao_width), 4)))));
  };
  while (k < k_max_cl) {
    k = log_step(k);
    K.push_back(k);
  }
  c->pinfo.k_size_cl = (int)K.size();
  while (k < k_max) {
    k = log_step(k);
    K.push_back(k);
  }
  c->pinfo.k_size = (int)K.size();
  c->pinfo.k_min = K.front();
  c->pinfo.k_max = K.back();
  return CLPP_SUCCESS;
}

static int tau_sampling(clpp_ctx* c, char* err) {
  const clpp_perturb_desc& p = c->pd;
  const clpp_background_desc& bg = c->bg;
  const clpp_thermo_desc& th = c->th;
  std::vector<double> pvb(bg.bg_size), pvt(th.th_size);
  int first_b = 0, first_t = 0;

  const bool has_cmb = p.has_cl_cmb_temperature || p.has_cl_cmb_polarization;
  CLPP_CHECK(has_cmb, err, "the B200 path needs at least one CMB source (tCl/pCl) in the output");

  auto ratio_at = [&](double tau, double* r) -> int {
    if (clpp_background_at_tau(c, tau, bg.bg_size_short, CLPP_INTER_NORMAL, &first_b, pvb.data(), err)) return 1;
    if (clpp_thermodynamics_at_z(c, 1. / pvb[bg.index_bg_a] - 1., CLPP_INTER_NORMAL, &first_t, pvb.data(), pvt.data(),
                                 err))
      return 1;
    *r = pvb[bg.index_bg_a] * pvb[bg.index_bg_H] / pvt[th.index_th_dkappa];
    return 0;
  };

  // start of the sampling: bisection on tau_c/tau_h = aH/kappa' = start_sources_at_tau_c_over_tau_h
  double tau_lower = th.tau_ini, tau_upper = th.tau_rec, r;
  if (ratio_at(tau_lower, &r)) return CLPP_FAILURE;
  CLPP_CHECK(!(r > p.start_sources_at_tau_c_over_tau_h), err,
             "your choice of initial time for computing sources is inappropriate: it corresponds to an earlier time "
             "than the one at which the integration of thermodynamical variables started (tau=%g). You should "
             "increase either 'start_sources_at_tau_c_over_tau_h' or 'recfast_z_initial'\n", tau_lower);
  if (ratio_at(tau_upper, &r)) return CLPP_FAILURE;
  CLPP_CHECK(!(r < p.start_sources_at_tau_c_over_tau_h), err,
             "your choice of initial time for computing sources is inappropriate: it corresponds to a time after "
             "recombination. You should decrease 'start_sources_at_tau_c_over_tau_h'\n");
  double tau_mid = 0.5 * (tau_lower + tau_upper);
  while (tau_upper - tau_lower > p.tol_tau_approx) {
    if (ratio_at(tau_mid, &r)) return CLPP_FAILURE;
    if (r > p.start_sources_at_tau_c_over_tau_h) tau_upper = tau_mid; else tau_lower = tau_mid;
    tau_mid = 0.5 * (tau_lower + tau_upper);
  }
  const double tau_ini = tau_mid;

  // march to today with steps = perturb_sampling_stepsize / sqrt(rate_thermo^2 + |2a''/a - (a'/a)^2|)
  std::vector<double>& T = c->tau;
  T.clear();
  int last_b = first_b, last_t = first_t;
  double tau = tau_ini;
  T.push_back(tau);
  while (tau < bg.conformal_age) {
    if (clpp_background_at_tau(c, tau, bg.bg_size_short, CLPP_INTER_CLOSEBY, &last_b, pvb.data(), err)) return 1;
    if (clpp_thermodynamics_at_z(c, 1. / pvb[bg.index_bg_a] - 1., CLPP_INTER_CLOSEBY, &last_t, pvb.data(), pvt.data(),
                                 err))
      return 1;
    const double rate_thermo = pvt[th.index_th_rate];
    const double a_prime_over_a = pvb[bg.index_bg_H] * pvb[bg.index_bg_a];
    const double a_primeprime_over_a = pvb[bg.index_bg_H_prime] * pvb[bg.index_bg_a] + 2. * a_prime_over_a * a_prime_over_a;
    const double rate_isw_squared = fabs(2. * a_primeprime_over_a - a_prime_over_a * a_prime_over_a);
    double timescale_source = sqrt(rate_thermo * rate_thermo + rate_isw_squared);
    CLPP_CHECK(timescale_source != 0., err, "null evolution rate, integration is diverging");
    timescale_source = 1. / timescale_source;
    CLPP_CHECK(!(fabs(p.perturb_sampling_stepsize * timescale_source / tau) < p.smallest_allowed_variation), err,
               "integration step =%e < machine precision : leads either to numerical error or infinite loop",
               p.perturb_sampling_stepsize * timescale_source);
    tau = tau + p.perturb_sampling_stepsize * timescale_source;
    T.push_back(tau);
  }
  T.back() = bg.conformal_age;  // the last sample sits exactly on today
  c->pinfo.tau_size = (int)T.size();

  CLPP_CHECK(p.z_max_pk >= 0, err, "asked for negative redshift z=%e", p.z_max_pk);
  // z_max_pk > 0: the last samples (from the one before tau(z_max_pk), plus four more) form the late-time table that
  // perturb_sources_at_tau splines in ln(tau) (:1541-1592); the sampling itself does not change
  c->pinfo.ln_tau_size = 1;
  if (p.z_max_pk > 0.) {
    const double a_target = bg.a_today / (1. + p.z_max_pk);
    const HostTable& B = c->bgt;
    int i = 0;
    while (i < B.n_lines - 2 && B.y[(size_t)(i + 1) * B.n_cols + bg.index_bg_a] < a_target) i++;
    const double a0 = B.y[(size_t)i * B.n_cols + bg.index_bg_a], a1 = B.y[(size_t)(i + 1) * B.n_cols + bg.index_bg_a];
    const double tau_lower = B.x[i] + (B.x[i + 1] - B.x[i]) * (a_target - a0) / (a1 - a0);
    CLPP_CHECK(tau_lower > T[0], err,
               "you asked for zmax=%e, i.e. taumin=%e, smaller than or equal to the first possible value =%e; it should be "
               "strictly bigger for a successfull interpolation", p.z_max_pk, tau_lower, T[0]);
    int first = 0;
    while (T[first] < tau_lower) first++;
    first = std::max(first - 5, 0);
    c->pinfo.ln_tau_size = (int)T.size() - first;
  }
  return CLPP_SUCCESS;
}

int clpp_host_perturb_grids(clpp_ctx* c, char* err) {
  const clpp_perturb_desc& p = c->pd;
  const clpp_background_desc& bg = c->bg;
  CLPP_CHECK(c->has_bg && c->has_th, err, "background and thermodynamics tables must be set before the perturbation grids");
  CLPP_CHECK(p.gauge == 1, err, "the B200 path integrates in the synchronous gauge only");
  CLPP_CHECK(bg.has_cdm, err,
             "In the synchronous gauge, it is not self-consistent to assume no CDM: the later is used to define the "
             "initial timelike hypersurface. You can either add a negligible amount of CDM or switch to newtonian gauge");
  CLPP_CHECK(bg.sgnK == 0 && !bg.has_curvature, err, "the B200 path supports flat space only (K=0)");
  CLPP_CHECK(!bg.has_fld && !bg.has_scf && !bg.has_dcdm && !bg.has_dr && !bg.has_idr && !bg.has_idm_dr, err,
             "species fld/scf/dcdm/dr/idr/idm_dr are not supported by the B200 path");
  CLPP_CHECK(p.evolver == 0 || p.evolver == 1, err, "evolver = %d: must be 0 (rk) or 1 (ndf15)", p.evolver);
  CLPP_CHECK(p.tight_coupling_approximation >= CLPP_TCA_FIRST_ORDER_MB &&
                 p.tight_coupling_approximation <= CLPP_TCA_COMPROMISE_CLASS,
             err, "your tight_coupling_approximation is set to %d, out of range defined in perturbations.h",
             p.tight_coupling_approximation);
  CLPP_CHECK(p.tight_coupling_approximation == CLPP_TCA_COMPROMISE_CLASS ||
                 p.tight_coupling_approximation == CLPP_TCA_FIRST_ORDER_CAMB ||
                 p.tight_coupling_approximation == CLPP_TCA_FIRST_ORDER_MB,
             err, "tight_coupling_approximation=%d not implemented on the B200 path (0,1,5 are)",
             p.tight_coupling_approximation);
  CLPP_CHECK(p.radiation_streaming_approximation >= CLPP_RSA_NULL && p.radiation_streaming_approximation <= CLPP_RSA_NONE,
             err, "your radiation_streaming_approximation is set to %d, out of range defined in perturbations.h",
             p.radiation_streaming_approximation);
  if (bg.has_ur)
    CLPP_CHECK(p.ur_fluid_approximation >= CLPP_UFA_MB && p.ur_fluid_approximation <= CLPP_UFA_NONE, err,
               "your ur_fluid_approximation is set to %d, out of range defined in perturbations.h",
               p.ur_fluid_approximation);
  if (bg.has_ncdm) {
    CLPP_CHECK(p.ncdm_fluid_approximation >= CLPP_NCDMFA_MB && p.ncdm_fluid_approximation <= CLPP_NCDMFA_NONE, err,
               "your ncdm_fluid_approximation is set to %d, out of range defined in perturbations.h",
               p.ncdm_fluid_approximation);
    CLPP_CHECK(c->N_ncdm == bg.N_ncdm, err, "ncdm momentum grids not set (clpp_set_ncdm) for %d species", bg.N_ncdm);
  }
  CLPP_CHECK(p.l_max_g >= 4, err,
             "ppr->l_max_g should be at least 4, i.e. we must integrate at least over photon density, velocity, shear, "
             "third and fourth momentum");
  CLPP_CHECK(p.l_max_pol_g >= 4, err, "ppr->l_max_pol_g should be at least 4");
  if (bg.has_ur) CLPP_CHECK(p.l_max_ur >= 4, err, "ppr->l_max_ur should be at least 4");
  if (bg.has_ncdm) CLPP_CHECK(p.l_max_ncdm >= 4, err, "ppr->l_max_ncdm should be at least 4");
  CLPP_CHECK(bg.h <= 1.5 && bg.h >= 0.3, err,
             "Your value of pba->h=%e is out of the bounds [%e , %e] and could cause a crash of the perturbation ODE "
             "integration.", bg.h, 0.3, 1.5);
  CLPP_CHECK(!(bg.Omega0_b * bg.h * bg.h < 0.005) && !(bg.Omega0_b * bg.h * bg.h > 0.039), err,
             "Your value of omega_b=%e is out of the bounds [%e , %e] and could cause a crash of the perturbation ODE "
             "integration.", bg.Omega0_b * bg.h * bg.h, 0.005, 0.039);

  // source types, in the reference's order: t2, p | t0, t1, delta_m, delta_cb, ..., phi_plus_psi
  clpp_perturb_info& I = c->pinfo;
  int tp = 0;
  const bool has_t = p.has_cl_cmb_temperature, has_p = p.has_cl_cmb_polarization;
  const bool has_pp = p.has_cl_cmb_lensing_potential;
  const bool has_dm = p.has_pk_matter || p.has_nl_corrections_based_on_delta_m;
  const bool has_dcb = has_dm && bg.has_ncdm;
  I.index_tp_t2 = has_t ? tp++ : -1;
  I.index_tp_p = has_p ? tp++ : -1;
  I.index_tp_t0 = has_t ? tp++ : -1;
  I.index_tp_t1 = has_t ? tp++ : -1;
  I.index_tp_delta_m = has_dm ? tp++ : -1;
  I.index_tp_delta_cb = has_dcb ? tp++ : -1;
  I.index_tp_phi_plus_psi = has_pp ? tp++ : -1;
  I.tp_size = tp;
  CLPP_CHECK(tp > 0, err,
             "inconsistent input: you asked for scalars, so you should have at least one non-zero scalar source type");

  if (k_list(c, err)) return CLPP_FAILURE;
  if (tau_sampling(c, err)) return CLPP_FAILURE;
  c->has_pgrids = true;
  return CLPP_SUCCESS;
}

// ------------------------------------------------------------------------------------------
int clpp_host_transfer_grids(clpp_ctx* c, char* err) {
  const clpp_transfer_desc& t = c->td;
  const clpp_background_desc& bg = c->bg;
  const clpp_thermo_desc& th = c->th;
  clpp_transfer_info& I = c->tinfo;
  CLPP_CHECK(c->has_pgrids, err, "perturbation grids (or injected sources) must exist before the transfer grids");
  CLPP_CHECK(bg.sgnK == 0, err, "the B200 path supports flat space only (K=0)");

  // transfer types, reference order: t2, e | t0, t1, lcmb
  int tt = 0;
  I.index_tt_t2 = t.has_cl_cmb_temperature ? tt++ : -1;
  I.index_tt_e = t.has_cl_cmb_polarization ? tt++ : -1;
  I.index_tt_t0 = t.has_cl_cmb_temperature ? tt++ : -1;
  I.index_tt_t1 = t.has_cl_cmb_temperature ? tt++ : -1;
  I.index_tt_lcmb = t.has_cl_cmb_lensing_potential ? tt++ : -1;
  I.tt_size = tt;
  CLPP_CHECK(tt > 0, err, "no harmonic-space transfer function requested");

  const double tau0 = bg.conformal_age;
  const double q_period = 2. * CLPP_PI / (tau0 - th.tau_rec) * th.angular_rescaling;

  // ---- q list (flat): logarithmic at small q turning linear (step q_period*q_linstep) at large q
  {
    const double q_min = c->pinfo.k_min;
    const double q_max = c->k[c->pinfo.k_size_cl - 1];
    const double q_logstep_spline = t.q_logstep_spline / pow(th.angular_rescaling, t.q_logstep_open);
    double q_step = 1. + q_period * t.q_logstep_spline;
    int q_size_max = 5 * (int)(log(q_max / q_min) / log(q_step));
    q_step = q_period * t.q_linstep;
    q_size_max += 5 * (int)((q_max - q_min) / q_step);
    std::vector<double>& Q = c->q;
    Q.clear();
    Q.push_back(q_min);
    while (Q.back() < q_max) {
      CLPP_CHECK((int)Q.size() < q_size_max, err, "buggy q-list definition");
      const double qp = Q.back();
      Q.push_back(qp + q_period * t.q_linstep * qp / (qp + t.q_linstep / q_logstep_spline));
    }
    if (Q.back() > q_max) Q.pop_back();
    CLPP_CHECK(Q.size() >= 2, err, "buggy q-list definition");
    I.q_size = (int)Q.size();
  }
  // ---- k(q) for scalars in flat space: k = sqrt(q^2 - K) = q, first value snapped onto k_min
  {
    c->kq.resize(I.q_size);
    for (int i = 0; i < I.q_size; i++) c->kq[i] = sqrt(c->q[i] * c->q[i] - bg.K * (0. + 1.));
    if (c->kq[0] < c->k[0]) {
      CLPP_CHECK((c->k[0] - c->kq[0]) < 10. * 2.2204460492503131e-16, err,
                 "bug in k_list calculation: in perturbation module k_min=%e, in transfer module k_min=%e, "
                 "interpolation impossible", c->k[0], c->kq[0]);
      c->kq[0] = c->k[0];
    }
    CLPP_CHECK(!(c->kq[I.q_size - 1] > c->k[c->pinfo.k_size_cl - 1]), err,
               "bug in k_list calculation: in perturbation module k_max=%e, in transfer module k_max=%e, "
               "interpolation impossible", c->k[c->pinfo.k_size_cl - 1], c->kq[I.q_size - 1]);
  }
  // ---- l list: logarithmic steps until the step reaches l_linstep, then linear; last = l_max
  {
    const int l_max = t.l_scalar_max;
    const double rho = th.angular_rescaling;
    std::vector<int>& L = c->l;
    L.clear();
    L.push_back(2);
    int increment = std::max((int)(L.back() * (pow(t.l_logstep, rho) - 1.)), 1);
    while (((L.back() + increment) < l_max) && (increment < t.l_linstep * rho)) {
      L.push_back(L.back() + increment);
      increment = std::max((int)(L.back() * (pow(t.l_logstep, rho) - 1.)), 1);
    }
    increment = t.l_linstep * rho;
    while ((L.back() + increment) <= l_max) L.push_back(L.back() + increment);
    if (L.back() != l_max) L.push_back(l_max);
    I.l_size_max = (int)L.size();
    // every CMB type runs to l_scalar_max; keep two guard multipoles when available
    c->l_size_tt.assign(I.tt_size, 0);
    I.l_size = 0;
    for (int itt = 0; itt < I.tt_size; itt++) {
      int il = 0;
      while (L[il] < l_max) il++;
      int n = il + 1;
      if (n < I.l_size_max) n++;
      if (n < I.l_size_max) n++;
      c->l_size_tt[itt] = n;
      I.l_size = std::max(I.l_size, n);
    }
  }
  c->has_tgrids = true;
  return CLPP_SUCCESS;
}
