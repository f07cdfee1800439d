// Stage 1: per-wavenumber Einstein-Boltzmann integration on the device.
//
// One WARP integrates one k-mode from its initial time to today (one CTA = one warp, so the
// hardware block scheduler is the work queue; modes are issued in decreasing-k order like the
// reference's task loop, perturbations_module.cpp:685).  Everything of a mode lives in shared
// memory: state vector, the NDF backward-difference array, the LU factors of (I - h/(G(1-alpha)) J).
// Error norms / step control use warp-shuffle reductions; the approximation state machine
// (tight coupling -> full hierarchy -> ur fluid / ncdm fluid -> radiation streaming) is per-mode
// device state: switch times are bisected on the device and the state vector is re-laid-out in
// shared memory at each switch.
//
// Reference functions restated here (all in source/perturbations_module.cpp unless noted):
//   perturb_solve :2463-2787, perturb_find_approximation_number/_switches :2940-3231,
//   perturb_approximations :5443-5670, perturb_vector_init :3271-4688,
//   perturb_initial_conditions :4723-5408, perturb_einstein :5840-6045,
//   perturb_total_stress_energy :6047-6703, perturb_sources_member :6731-7285,
//   perturb_derivs_member :7861-9218, perturb_tca_slip_and_shear :9229-9516,
//   perturb_rsa_delta_and_theta :9530-9636,
//   evolver_ndf15 (tools/evolver_ndf15.cpp:62-705), adjust_stepsize :907-943,
//   interp_from_dif :860-905, new_linearisation :945-998, ludcmp/lubksb :1001-1064.
//
// Differences of method (not of result): the Jacobian is assembled column by column from
// f(t, e_j) -- the system is linear and homogeneous in y, so this equals the reference's
// finite-difference numjac (:1213-1539) up to its O(1e-8) truncation noise -- and the linear
// algebra is a dense warp-parallel LU in shared memory instead of the CPU's sparse LU.
#include <cmath>

#include "device.h"

#define PT_MAX_NCDM 3
#define PT_MAX_INTERVALS 6
#define PT_FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------
struct PtParams {
  // tables
  const double *bg_tau, *bg_y, *bg_dd;
  int bt_size, bg_size, bg_size_normal;
  const double *th_z, *th_y, *th_dd;
  int tt_size, th_size;
  // background column indices
  int ia, iH, iHp, irho_g, irho_b, irho_cdm, irho_ur, irho_ncdm1, ip_ncdm1, ipseudo_p_ncdm1;
  // thermo column indices
  int ixe, idkappa, iddkappa, idddkappa, iexp_m_kappa, ig, idg, iddg, icb2, iwb, iTb, itau_d, irate, ir_d, idcb2, iddcb2;
  int compute_cb2_derivatives, compute_damping_scale;
  double th_linear_below_z;  // < 0: never use linear interpolation
  double n_e, YHe, T_cmb, tau_free_streaming;
  int has_ur, has_ncdm, N_ncdm;
  int ncdm_q_size[PT_MAX_NCDM], ncdm_q_off[PT_MAX_NCDM];
  double ncdm_M[PT_MAX_NCDM], ncdm_factor[PT_MAX_NCDM];
  const double *ncdm_q, *ncdm_w, *ncdm_dlnf0;
  double a_today;
  // precision
  double start_small_k_at_tau_c_over_tau_h, start_large_k_at_tau_h_over_tau_k;
  double tca_trigger_tau_c_over_tau_h, tca_trigger_tau_c_over_tau_k;
  int tca_method, rsa_method, ufa_method, ncdmfa_method;
  double rsa_trigger, ufa_trigger, ncdmfa_trigger;
  int l_max_g, l_max_pol_g, l_max_ur, l_max_ncdm;
  double tol_ncdm_initial_w, tol_tau_approx, rtol, hmin_allowed;
  double curvature_ini, three_ceff2_ur, three_cvis2_ur;
  int switch_sw, switch_eisw, switch_lisw, switch_dop, switch_pol;
  double eisw_lisw_split_z;
  // grids / output
  const double* k;
  int k_size;
  const double* tau;
  int tau_size;
  double* sources;  // [tp][k][tau]
  int tp_t0, tp_t1, tp_t2, tp_p, tp_delta_m, tp_delta_cb, tp_phi_plus_psi;
  const int* order;
  int n_modes;
  clpp_kstat* kstat;
  // per-slot scratch in global memory (Jacobian), indexed by blockIdx
  double* jac;
  int neq_max, ld;  // ld: odd leading dimension of the LU matrix in shared memory
};

struct Approx {
  int tca_off, rsa_on, ufa_on, ncdmfa_on;  // monotone flags (0 -> 1 in time)
};

struct Layout {
  int neq;
  int delta_g, theta_g, shear_g, l3_g, pol0_g;  // -1 when absent
  int delta_b, theta_b, delta_cdm;
  int delta_ur, theta_ur, shear_ur, l3_ur;
  int psi0_ncdm1;
  int eta;
  int l_max_g, l_max_pol_g, l_max_ur;
  int l_max_ncdm, q_size_ncdm[PT_MAX_NCDM], ncdm_off[PT_MAX_NCDM];
};

// background + thermodynamics + derived metric quantities at the current time; identical in
// all lanes (kept in registers)
struct Env {
  double tau, a, H, Hp, rho_g, rho_b, rho_cdm, rho_ur;
  double dkappa, ddkappa, exp_m_kappa, g, dg, cb2;
};

struct Metric {
  double h_prime, eta_prime, alpha, alpha_prime, h_prime_prime;
  double delta_rho, rho_plus_p_theta, rho_plus_p_shear, delta_p;
  double rsa_delta_g, rsa_theta_g, rsa_delta_ur, rsa_theta_ur;
  double delta_m, theta_m, delta_cb, theta_cb;
  double tca_shear_g, tca_slip;
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(PT_FULL, v, o));
  return v;
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PT_FULL, v, o);
  return v;
}

// bracket x in a growing array: closeby (cursor) or bisection; returns inf with X[inf] <= x <= X[inf+1]
__device__ __forceinline__ int table_locate(const double* __restrict__ X, int n, double x, int cursor, bool closeby) {
  int inf, sup;
  if (closeby) {
    inf = min(max(cursor, 0), n - 2);
    while (inf > 0 && x < X[inf]) inf--;
    sup = inf + 1;
    while (sup < n - 1 && x > X[sup]) sup++;
    inf = sup - 1;
  } else {
    inf = 0;
    sup = n - 1;
    while (sup - inf > 1) {
      const int mid = (int)(0.5 * (inf + sup));
      if (x < X[mid]) sup = mid; else inf = mid;
    }
  }
  return inf;
}

struct Mode {
  // shared-memory views of this warp
  double *pvb, *pvt;
  double *y, *ynew, *f0, *pred, *psi, *difkp1, *del, *invwt, *tmp, *yi, *ypi;
  double* dif;  // [7][neq_pad]
  double* LU;   // [neq][ld] column-major: LU[i + j*ld]
  int* piv;
  double* J;    // global scratch [neq][neq] column-major
  int neq_pad;
  // mode state
  double k, k2;
  int ik, lane;
  int cur_bg, cur_th;
  Approx ap;
  Layout L;
  Env e;
  Metric m;
  double tca_shear_last;  // photon shear of the last Newton-iteration RHS call (used by the sources while TCA is on)
  clpp_kstat st;
  int status;
};

// ---------------------------------------------------------------------------------------------
// background_at_tau (normal_info columns) + thermodynamics_at_z, cooperative over lanes
__device__ void env_at(const PtParams& P, Mode& M, double tau, bool closeby) {
  const int lane = M.lane;
  // background
  {
    const int inf = table_locate(P.bg_tau, P.bt_size, tau, M.cur_bg, closeby);
    M.cur_bg = inf;
    const double x0 = P.bg_tau[inf], x1 = P.bg_tau[inf + 1];
    const double h = x1 - x0, b = (tau - x0) / h, a = 1 - b;
    if (lane < P.bg_size_normal) {
      const size_t r0 = (size_t)inf * P.bg_size + lane, r1 = r0 + P.bg_size;
      M.pvb[lane] = a * P.bg_y[r0] + b * P.bg_y[r1] +
                    ((a * a * a - a) * P.bg_dd[r0] + (b * b * b - b) * P.bg_dd[r1]) * h * h / 6.;
    }
  }
  __syncwarp();
  const double av = M.pvb[P.ia], Hv = M.pvb[P.iH], Hp = M.pvb[P.iHp];
  const double z = 1. / av - 1.;
  // thermodynamics
  const double z_last = P.th_z[P.tt_size - 1];
  if (z >= z_last) {
    if (lane == 0) {
      const double* row = P.th_y + (size_t)(P.tt_size - 1) * P.th_size;
      const double x0 = row[P.ixe];
      double* pv = M.pvt;
      pv[P.ixe] = x0;
      pv[P.idkappa] = (1. + z) * (1. + z) * P.n_e * x0 * CLPP_sigma * CLPP_Mpc_over_m;
      pv[P.itau_d] = row[P.itau_d] * pow((1 + z) / (1. + z_last), 2);
      if (P.compute_damping_scale) pv[P.ir_d] = row[P.ir_d] * pow((1 + z) / (1. + z_last), -1.5);
      pv[P.iddkappa] = -Hv * 2. / (1. + z) * pv[P.idkappa];
      pv[P.idddkappa] = (Hv * Hv / (1. + z) - Hp) * 2. / (1. + z) * pv[P.idkappa];
      pv[P.iexp_m_kappa] = 0.;
      pv[P.ig] = 0.;
      pv[P.idg] = 0.;
      pv[P.iddg] = 0.;
      pv[P.iTb] = P.T_cmb * (1. + z);
      pv[P.iwb] = CLPP_k_B / (CLPP_c * CLPP_c * CLPP_m_H) * (1. + (1. / CLPP_not4 - 1.) * P.YHe + x0 * (1. - P.YHe)) *
                  P.T_cmb * (1. + z);
      pv[P.icb2] = pv[P.iwb] * 4. / 3.;
      if (P.compute_cb2_derivatives) {
        pv[P.idcb2] = -Hv * av * pv[P.icb2];
        pv[P.iddcb2] = -Hp * av * pv[P.icb2];
      }
      pv[P.irate] = pv[P.idkappa];
    }
  } else {
    const bool linear = (z < P.th_linear_below_z);
    const int inf = table_locate(P.th_z, P.tt_size, z, M.cur_th, closeby && !linear);
    M.cur_th = inf;
    const double x0 = P.th_z[inf], x1 = P.th_z[inf + 1];
    const double h = x1 - x0, b = (z - x0) / h, a = 1 - b;
    if (lane < P.th_size) {
      const size_t r0 = (size_t)inf * P.th_size + lane, r1 = r0 + P.th_size;
      double v = a * P.th_y[r0] + b * P.th_y[r1];
      if (!linear) v += ((a * a * a - a) * P.th_dd[r0] + (b * b * b - b) * P.th_dd[r1]) * h * h / 6.;
      M.pvt[lane] = v;
    }
  }
  __syncwarp();
  Env& e = M.e;
  e.tau = tau;
  e.a = av; e.H = Hv; e.Hp = Hp;
  e.rho_g = M.pvb[P.irho_g]; e.rho_b = M.pvb[P.irho_b]; e.rho_cdm = M.pvb[P.irho_cdm];
  e.rho_ur = P.has_ur ? M.pvb[P.irho_ur] : 0.;
  e.dkappa = M.pvt[P.idkappa]; e.ddkappa = M.pvt[P.iddkappa];
  e.exp_m_kappa = M.pvt[P.iexp_m_kappa]; e.g = M.pvt[P.ig]; e.dg = M.pvt[P.idg];
  e.cb2 = M.pvt[P.icb2];
}

// perturb_approximations: flags at time tau (uses bisection lookups, inter_normal)
__device__ Approx approximations_at(const PtParams& P, Mode& M, double tau) {
  env_at(P, M, tau, false);
  const Env& e = M.e;
  Approx a;
  const double tau_k = 1. / M.k, tau_h = 1. / (e.H * e.a);
  if (e.dkappa == 0.) a.tca_off = 1;
  else {
    const double tau_c = 1. / e.dkappa;
    a.tca_off = ((tau_c / tau_h < P.tca_trigger_tau_c_over_tau_h) && (tau_c / tau_k < P.tca_trigger_tau_c_over_tau_k)) ? 0 : 1;
  }
  a.rsa_on = ((tau / tau_k > P.rsa_trigger) && (tau > P.tau_free_streaming) && (P.rsa_method != CLPP_RSA_NONE)) ? 1 : 0;
  a.ufa_on = (P.has_ur && (tau / tau_k > P.ufa_trigger) && (P.ufa_method != CLPP_UFA_NONE)) ? 1 : 0;
  a.ncdmfa_on = (P.has_ncdm && (tau / tau_k > P.ncdmfa_trigger) && (P.ncdmfa_method != CLPP_NCDMFA_NONE)) ? 1 : 0;
  return a;
}

__device__ __forceinline__ int approx_flag(const Approx& a, int which) {
  return which == 0 ? a.tca_off : which == 1 ? a.rsa_on : which == 2 ? a.ufa_on : a.ncdmfa_on;
}

// perturb_vector_init (index part): layout of the state vector for a set of approximations
__device__ Layout make_layout(const PtParams& P, const Approx& ap) {
  Layout L;
  int n = 0;
  L.delta_g = L.theta_g = L.shear_g = L.l3_g = L.pol0_g = -1;
  L.delta_ur = L.theta_ur = L.shear_ur = L.l3_ur = -1;
  L.psi0_ncdm1 = -1;
  L.l_max_g = P.l_max_g; L.l_max_pol_g = P.l_max_pol_g; L.l_max_ur = P.l_max_ur;
  if (!ap.rsa_on) {
    L.delta_g = n++;
    L.theta_g = n++;
    if (ap.tca_off) {
      L.shear_g = n++;
      L.l3_g = n; n += P.l_max_g - 2;
      L.pol0_g = n; n += P.l_max_pol_g + 1;
    }
  }
  L.delta_b = n++;
  L.theta_b = n++;
  L.delta_cdm = n++;
  if (P.has_ur && !ap.rsa_on) {
    L.delta_ur = n++;
    L.theta_ur = n++;
    L.shear_ur = n++;
    if (!ap.ufa_on) { L.l3_ur = n; n += P.l_max_ur - 2; }
  }
  L.l_max_ncdm = 0;
  if (P.has_ncdm) {
    L.psi0_ncdm1 = n;
    L.l_max_ncdm = ap.ncdmfa_on ? 2 : P.l_max_ncdm;
    for (int s = 0; s < P.N_ncdm; s++) {
      L.q_size_ncdm[s] = ap.ncdmfa_on ? 1 : P.ncdm_q_size[s];
      L.ncdm_off[s] = n;
      n += (L.l_max_ncdm + 1) * L.q_size_ncdm[s];
    }
  }
  L.eta = n++;
  L.neq = n;
  return L;
}

// ---------------------------------------------------------------------------------------------
// Right-hand side f(tau, y) for the environment currently stored in M.e/M.pvb/M.pvt.
// Also fills M.m (metric and derived quantities) for the caller.
template <bool WANT_MATTER>
__device__ void rhs_apply(const PtParams& P, Mode& M, const double* __restrict__ y, double* __restrict__ dy) {
  const Layout& L = M.L;
  const Approx& ap = M.ap;
  const Env& e = M.e;
  const int lane = M.lane;
  const double k = M.k, k2 = M.k2;
  const double a = e.a, a2 = a * a, aH = e.H * a;
  const double R = 4. / 3. * e.rho_g / e.rho_b;
  Metric& m = M.m;

  // ---- perturb_total_stress_energy
  double delta_g = 0., theta_g = 0., shear_g = 0.;
  if (ap.tca_off) {
    if (!ap.rsa_on) { delta_g = y[L.delta_g]; theta_g = y[L.theta_g]; shear_g = y[L.shear_g]; }
  } else {
    delta_g = y[L.delta_g]; theta_g = y[L.theta_g]; shear_g = 0.;
  }
  double delta_ur = 0., theta_ur = 0., shear_ur = 0.;
  if (P.has_ur && !ap.rsa_on) { delta_ur = y[L.delta_ur]; theta_ur = y[L.theta_ur]; shear_ur = y[L.shear_ur]; }
  const double delta_b = y[L.delta_b], theta_b = y[L.theta_b], delta_cdm = y[L.delta_cdm], eta = y[L.eta];
  const double delta_p_b_over_rho_b = e.cb2 * delta_b;

  double delta_rho = e.rho_g * delta_g + e.rho_b * delta_b;
  double rpt = 4. / 3. * e.rho_g * theta_g + e.rho_b * theta_b;
  double rps = 4. / 3. * e.rho_g * shear_g;
  double delta_p = 1. / 3. * e.rho_g * delta_g + e.rho_b * delta_p_b_over_rho_b;
  double delta_rho_m = 0., rho_m = 0., rpt_m = 0., rpm = 0.;
  if (WANT_MATTER) {
    delta_rho_m = e.rho_b * delta_b; rho_m = e.rho_b;
    rpt_m = e.rho_b * theta_b; rpm = e.rho_b;
  }
  delta_rho += e.rho_cdm * delta_cdm;
  if (WANT_MATTER) { delta_rho_m += e.rho_cdm * delta_cdm; rho_m += e.rho_cdm; rpm += e.rho_cdm; }
  if (P.has_ur) {
    delta_rho = delta_rho + e.rho_ur * delta_ur;
    rpt = rpt + 4. / 3. * e.rho_ur * theta_ur;
    rps = rps + 4. / 3. * e.rho_ur * shear_ur;
    delta_p += 1. / 3. * e.rho_ur * delta_ur;
  }
  if (WANT_MATTER) {
    m.delta_cb = delta_rho_m / rho_m;
    m.theta_cb = rpt_m / rpm;
  }
  if (P.has_ncdm) {
    for (int s = 0; s < P.N_ncdm; s++) {
      const double rho_n = M.pvb[P.irho_ncdm1 + s], p_n = M.pvb[P.ip_ncdm1 + s];
      double d_n, t_n;
      if (ap.ncdmfa_on) {
        const double pseudo_p = M.pvb[P.ipseudo_p_ncdm1 + s];
        const double w_n = p_n / rho_n;
        const double cg2 = w_n * (1.0 - 1.0 / (3.0 + 3.0 * w_n) * (3.0 * w_n - 2.0 + pseudo_p / p_n));
        const int idx = L.ncdm_off[s];
        d_n = y[idx]; t_n = y[idx + 1];
        delta_rho += rho_n * y[idx];
        rpt += (rho_n + p_n) * y[idx + 1];
        rps += (rho_n + p_n) * y[idx + 2];
        delta_p += cg2 * rho_n * y[idx];
      } else {
        const double factor = P.ncdm_factor[s] * pow(P.a_today / a, 4);
        double s_rho = 0., s_theta = 0., s_shear = 0., s_p = 0.;
        const int nq = L.q_size_ncdm[s], stride = L.l_max_ncdm + 1;
        const double Ms = P.ncdm_M[s];
        for (int iq = lane; iq < nq; iq += 32) {
          const int idx = L.ncdm_off[s] + iq * stride;
          const double q = P.ncdm_q[P.ncdm_q_off[s] + iq], w0 = P.ncdm_w[P.ncdm_q_off[s] + iq];
          const double q2 = q * q, eps = sqrt(q2 + Ms * Ms * a2);
          s_rho += q2 * eps * w0 * y[idx];
          s_theta += q2 * q * w0 * y[idx + 1];
          s_shear += q2 * q2 / eps * w0 * y[idx + 2];
          s_p += q2 * q2 / eps * w0 * y[idx];
        }
        s_rho = wsum(s_rho) * factor;
        s_theta = wsum(s_theta) * k * factor;
        s_shear = wsum(s_shear) * 2.0 / 3.0 * factor;
        s_p = wsum(s_p) * factor / 3.;
        d_n = s_rho / rho_n;
        t_n = s_theta / (rho_n + p_n);
        delta_rho += s_rho; rpt += s_theta; rps += s_shear; delta_p += s_p;
      }
      if (WANT_MATTER) {
        delta_rho_m += rho_n * d_n; rho_m += rho_n;
        rpt_m += (rho_n + p_n) * t_n; rpm += (rho_n + p_n);
      }
    }
  }
  if (WANT_MATTER) {
    m.delta_m = delta_rho_m / rho_m;
    m.theta_m = rpt_m / rpm;
  }

  // ---- perturb_einstein (synchronous gauge, K = 0)
  const double h_prime = (k2 * eta + 1.5 * a2 * delta_rho) / (0.5 * aH);
  double rsa_delta_g = 0., rsa_theta_g = 0., rsa_delta_ur = 0., rsa_theta_ur = 0.;
  if (ap.rsa_on) {
    if (P.rsa_method != CLPP_RSA_NULL) {
      rsa_delta_g = 4. / k2 * (aH * h_prime - k2 * eta);
      rsa_theta_g = -0.5 * h_prime;
    }
    if (P.rsa_method == CLPP_RSA_MD_WITH_REIO) {
      rsa_delta_g += -4. / k2 * e.dkappa * (theta_b + 0.5 * h_prime);
      rsa_theta_g += 3. / k2 * (e.ddkappa * (theta_b + 0.5 * h_prime) +
                                e.dkappa * (-aH * theta_b + e.cb2 * k2 * delta_b - aH * h_prime + k2 * eta));
    }
    if (P.has_ur && P.rsa_method != CLPP_RSA_NULL) {
      rsa_delta_ur = 4. / k2 * (aH * h_prime - k2 * eta);
      rsa_theta_ur = -0.5 * h_prime;
    }
    delta_rho += e.rho_g * rsa_delta_g;
    rpt += 4. / 3. * e.rho_g * rsa_theta_g;
    if (P.has_ur) {
      delta_rho += e.rho_ur * rsa_delta_ur;
      rpt += 4. / 3. * e.rho_ur * rsa_theta_ur;
    }
  }
  const double eta_prime = (1.5 * a2 * rpt) / k2;
  const double h_prime_prime = -2. * aH * h_prime + 2. * k2 * eta - 9. * a2 * delta_p;
  const double alpha = (h_prime + 6. * eta_prime) / 2. / k2;
  if (!ap.tca_off) {
    const double sg = 16. / 45. / e.dkappa * (theta_g + k2 * alpha);
    rps += 4. / 3. * e.rho_g * sg;
  }
  const double alpha_prime = -2. * aH * alpha + eta - 4.5 * (a2 / k2) * rps;
  if (WANT_MATTER) {
    m.delta_m += 3. * a * e.H * m.theta_m / k2;
    m.delta_cb += 3. * a * e.H * m.theta_cb / k2;
  }
  m.h_prime = h_prime; m.eta_prime = eta_prime; m.alpha = alpha; m.alpha_prime = alpha_prime;
  m.h_prime_prime = h_prime_prime;
  m.delta_rho = delta_rho; m.rho_plus_p_theta = rpt; m.rho_plus_p_shear = rps; m.delta_p = delta_p;
  m.rsa_delta_g = rsa_delta_g; m.rsa_theta_g = rsa_theta_g; m.rsa_delta_ur = rsa_delta_ur; m.rsa_theta_ur = rsa_theta_ur;
  if (dy == nullptr) return;

  // ---- perturb_derivs
  const double cotKgen = 1.0 / (k * e.tau);
  const double metric_continuity = h_prime / 2.;
  const double metric_shear = k2 * alpha;
  const double metric_ufa_class = h_prime / 2.;
  if (ap.rsa_on) { delta_g = rsa_delta_g; theta_g = rsa_theta_g; }

  double dtheta_b;
  if (ap.tca_off) {
    dtheta_b = -aH * theta_b + k2 * delta_p_b_over_rho_b + R * e.dkappa * (theta_g - theta_b);
  } else {
    // ---- perturb_tca_slip_and_shear
    const double a_primeprime_over_a = e.Hp * a + 2. * aH * aH;
    const double tau_c = 1. / e.dkappa;
    const double dtau_c = -e.ddkappa * tau_c * tau_c;
    const double F = tau_c / (1 + R);
    double F_prime = 0.;
    if (P.tca_method >= CLPP_TCA_SECOND_ORDER_CLASS) F_prime = dtau_c / (1 + R) + tau_c * aH * R / (1 + R) / (1 + R);
    const double metric_shear_prime = k2 * alpha_prime;
    double slip;
    if (P.tca_method == CLPP_TCA_FIRST_ORDER_MB) {
      slip = 2. * R / (1. + R) * aH * (theta_b - theta_g) +
             F * (-a_primeprime_over_a * theta_b +
                  k2 * (-aH * delta_g / 2. + e.cb2 * (-theta_b - metric_continuity) - 4. / 3. * (-theta_g - metric_continuity) / 4.));
    } else {
      slip = (dtau_c / tau_c - 2. * aH / (1. + R)) * (theta_b - theta_g) +
             F * (-a_primeprime_over_a * theta_b +
                  k2 * (-aH * delta_g / 2. + e.cb2 * (-theta_b - metric_continuity) - 4. / 3. * (-theta_g - metric_continuity) / 4.));
    }
    double sg = 16. / 45. * tau_c * (theta_g + metric_shear);
    const double theta_prime = (-aH * theta_b + k2 * (e.cb2 * delta_b + R / 4. * delta_g)) / (1. + R);
    const double shear_g_prime = 16. / 45. * (tau_c * (theta_prime + metric_shear_prime) + dtau_c * (theta_g + metric_shear));
    if (P.tca_method == CLPP_TCA_COMPROMISE_CLASS) {
      slip = (1. - 2. * aH * F) * slip +
             F * k2 * (2. * aH * sg + shear_g_prime - (1. / 3. - e.cb2) * (F * theta_prime + 2. * F_prime * theta_b));
      sg = (1. - 11. / 6. * dtau_c) * sg - 11. / 6. * tau_c * 16. / 45. * tau_c * (theta_prime + metric_shear_prime);
    }
    m.tca_shear_g = sg;
    m.tca_slip = slip;
    dtheta_b = (-aH * theta_b + k2 * (delta_p_b_over_rho_b + R * (delta_g / 4. - sg)) + R * slip) / (1. + R);
  }

  if (lane == 0) {
    if (!ap.rsa_on) dy[L.delta_g] = -4. / 3. * (theta_g + metric_continuity);
    dy[L.delta_b] = -(theta_b + metric_continuity);
    dy[L.theta_b] = dtheta_b;
    dy[L.delta_cdm] = -metric_continuity;
    dy[L.eta] = eta_prime;
  }
  if (!ap.rsa_on) {
    if (ap.tca_off) {
      const int lg = L.l_max_g, lp = L.l_max_pol_g;
      const double* yg = y + L.delta_g;  // yg[l] = F_l (l>=3), yg[2] = shear
      const double* yp = y + L.pol0_g;
      const double P0 = (yp[0] + yp[2] + 2. * yg[2]) / 8.;
      if (lane == 1) {
        dy[L.theta_g] = k2 * (delta_g / 4. - yg[2]) + e.dkappa * (theta_b - theta_g);
        dy[L.shear_g] = 0.5 * (8. / 15. * (theta_g + metric_shear) - 3. / 5. * k * yg[3] - e.dkappa * (2. * yg[2] - 4. / 5. * P0));
        dy[L.l3_g] = k / 7.0 * (3. * 2. * yg[2] - 4. * yg[4]) - e.dkappa * yg[3];
      }
      if (lane == 2) {
        dy[L.pol0_g] = -k * yp[1] - e.dkappa * (yp[0] - 4. * P0);
        dy[L.pol0_g + 1] = k / 3. * (yp[0] - 2. * yp[2]) - e.dkappa * yp[1];
        dy[L.pol0_g + 2] = k / 5. * (2. * yp[1] - 3. * yp[3]) - e.dkappa * (yp[2] - 4. / 5. * P0);
      }
      for (int l = 4 + lane; l <= lg; l += 32) {
        if (l < lg) dy[L.delta_g + l] = k / (2.0 * l + 1.0) * (l * yg[l - 1] - (l + 1) * yg[l + 1]) - e.dkappa * yg[l];
        else dy[L.delta_g + l] = k * (yg[l - 1] - (1. + l) * cotKgen * yg[l]) - e.dkappa * yg[l];
      }
      for (int l = 3 + lane; l <= lp; l += 32) {
        if (l < lp) dy[L.pol0_g + l] = k / (2. * l + 1) * (l * yp[l - 1] - (l + 1.) * yp[l + 1]) - e.dkappa * yp[l];
        else dy[L.pol0_g + l] = k * (yp[l - 1] - (l + 1) * cotKgen * yp[l]) - e.dkappa * yp[l];
      }
    } else if (lane == 1) {
      dy[L.theta_g] = -(dtheta_b + aH * theta_b - k2 * delta_p_b_over_rho_b) / R + k2 * (0.25 * delta_g - m.tca_shear_g);
    }
  }
  if (P.has_ur && !ap.rsa_on) {
    const double* yu = y + L.delta_ur;
    if (lane == 3) {
      dy[L.delta_ur] = -4. / 3. * (yu[1] + metric_continuity) +
                       (1. - P.three_ceff2_ur) * aH * (yu[0] + 4. * aH * yu[1] / k / k);
      dy[L.theta_ur] = k2 * (P.three_ceff2_ur * yu[0] / 4. - yu[2]) - (1. - P.three_ceff2_ur) * aH * yu[1];
      if (!ap.ufa_on) {
        dy[L.shear_ur] = 0.5 * (8. / 15. * (yu[1] + metric_shear) - 3. / 5. * k * yu[3] -
                                (1. - P.three_cvis2_ur) * (8. / 15. * (yu[1] + metric_shear)));
        dy[L.l3_ur] = k / 7. * (3. * 2. * yu[2] - 4. * yu[4]);
      } else {
        if (P.ufa_method == CLPP_UFA_MB) dy[L.shear_ur] = -3. / e.tau * yu[2] + 2. / 3. * (yu[1] + metric_shear);
        else if (P.ufa_method == CLPP_UFA_HU) dy[L.shear_ur] = -3. * aH * yu[2] + 2. / 3. * (yu[1] + metric_shear);
        else dy[L.shear_ur] = -3. / e.tau * yu[2] + 2. / 3. * (yu[1] + metric_ufa_class);
      }
    }
    if (!ap.ufa_on) {
      const int lu = L.l_max_ur;
      for (int l = 4 + lane; l <= lu; l += 32) {
        if (l < lu) dy[L.delta_ur + l] = k / (2. * l + 1) * (l * yu[l - 1] - (l + 1.) * yu[l + 1]);
        else dy[L.delta_ur + l] = k * (yu[l - 1] - (1. + l) * cotKgen * yu[l]);
      }
    }
  }
  if (P.has_ncdm) {
    if (ap.ncdmfa_on) {
      if (lane < P.N_ncdm) {
        const int s = lane;
        const double rho_n = M.pvb[P.irho_ncdm1 + s], p_n = M.pvb[P.ip_ncdm1 + s], pseudo_p = M.pvb[P.ipseudo_p_ncdm1 + s];
        const double pseudo_p_over_p = pseudo_p / p_n;
        const double w_n = p_n / rho_n;
        const double ca2 = w_n / 3.0 / (1.0 + w_n) * (5.0 - pseudo_p / p_n);
        const double ceff2 = ca2;
        const double cvis2 = (P.ncdmfa_method == CLPP_NCDMFA_HU) ? w_n : 3. * w_n * ca2;
        const int idx = L.ncdm_off[s];
        dy[idx] = -(1.0 + w_n) * (y[idx + 1] + metric_continuity) - 3.0 * aH * (ceff2 - w_n) * y[idx];
        dy[idx + 1] = -aH * (1.0 - 3.0 * ca2) * y[idx + 1] + ceff2 / (1.0 + w_n) * k2 * y[idx] - k2 * y[idx + 2];
        if (P.ncdmfa_method == CLPP_NCDMFA_MB)
          dy[idx + 2] = -3.0 * (aH * (2. / 3. - ca2 - pseudo_p_over_p / 3.) + 1. / e.tau) * y[idx + 2] +
                        8.0 / 3.0 * cvis2 / (1.0 + w_n) * (y[idx + 1] + metric_shear);
        else if (P.ncdmfa_method == CLPP_NCDMFA_HU)
          dy[idx + 2] = -3.0 * aH * ca2 / w_n * y[idx + 2] + 8.0 / 3.0 * cvis2 / (1.0 + w_n) * (y[idx + 1] + metric_shear);
        else
          dy[idx + 2] = -3.0 * (aH * (2. / 3. - ca2 - pseudo_p_over_p / 3.) + 1. / e.tau) * y[idx + 2] +
                        8.0 / 3.0 * cvis2 / (1.0 + w_n) * (y[idx + 1] + metric_ufa_class);
      }
    } else {
      const int stride = L.l_max_ncdm + 1, lm = L.l_max_ncdm;
      for (int s = 0; s < P.N_ncdm; s++) {
        const int tot = L.q_size_ncdm[s] * stride;
        const double Ms = P.ncdm_M[s];
        for (int e_i = lane; e_i < tot; e_i += 32) {
          const int iq = e_i / stride, l = e_i - iq * stride;
          const int idx = L.ncdm_off[s] + iq * stride;
          const double q = P.ncdm_q[P.ncdm_q_off[s] + iq];
          const double dlnf0 = P.ncdm_dlnf0[P.ncdm_q_off[s] + iq];
          const double eps = sqrt(q * q + a2 * Ms * Ms);
          const double qk = k * q / eps;
          double v;
          if (l == 0) v = -qk * y[idx + 1] + metric_continuity * dlnf0 / 3.;
          else if (l == 1) v = qk / 3.0 * (y[idx] - 2 * y[idx + 2]);
          else if (l == 2) v = qk / 5.0 * (2 * y[idx + 1] - 3. * y[idx + 3]) - metric_shear * 2. / 15. * dlnf0;
          else if (l < lm) v = qk / (2. * l + 1.0) * (l * y[idx + (l - 1)] - (l + 1.) * y[idx + (l + 1)]);
          else v = qk * y[idx + l - 1] - (1. + l) * k * cotKgen * y[idx + l];
          dy[idx + l] = v;
        }
      }
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// perturb_sources_member: source functions at sample index_tau from (y, dy)
__device__ void write_sources(const PtParams& P, Mode& M, double tau, const double* y, const double* dy, int index_tau) {
  env_at(P, M, tau, true);
  rhs_apply<true>(P, M, y, nullptr);
  if (M.lane != 0) return;
  const Layout& L = M.L;
  const Approx& ap = M.ap;
  const Env& e = M.e;
  const Metric& m = M.m;
  const double k = M.k;
  const double z = P.a_today / e.a - 1.;
  const double aH = e.a * e.H;
  const double aH_prime = e.Hp * e.a + pow(e.H * e.a, 2);
  double delta_g, Pi;
  if (ap.rsa_on) { delta_g = m.rsa_delta_g; Pi = 0.; }
  else {
    delta_g = y[L.delta_g];
    if (!ap.tca_off) Pi = 5. * M.tca_shear_last / 8.;
    else Pi = (y[L.pol0_g] + y[L.pol0_g + 2] + 2. * y[L.shear_g]) / 8.;
  }
  const size_t stride_tp = (size_t)P.k_size * P.tau_size;
  double* out = P.sources + (size_t)M.ik * P.tau_size + index_tau;
  if (P.tp_t0 >= 0) {
    int switch_isw = 1;
    if ((P.switch_eisw == 0) && (z >= P.eisw_lisw_split_z)) switch_isw = 0;
    if ((P.switch_lisw == 0) && (z < P.eisw_lisw_split_z)) switch_isw = 0;
    const double eta = y[L.eta], theta_b = y[L.theta_b], dtheta_b = dy[L.theta_b];
    out[P.tp_t0 * stride_tp] =
        P.switch_sw * e.g * (delta_g / 4. + m.alpha_prime) +
        switch_isw * (e.g * (eta - m.alpha_prime - 2 * aH * m.alpha) +
                      e.exp_m_kappa * 2. * (m.eta_prime - aH_prime * m.alpha - aH * m.alpha_prime)) +
        P.switch_dop * (e.g * (dtheta_b / k / k + m.alpha_prime) + e.dg * (theta_b / k / k + m.alpha));
    out[P.tp_t1 * stride_tp] = switch_isw * e.exp_m_kappa * k * (m.alpha_prime + 2. * aH * m.alpha - eta);
    out[P.tp_t2 * stride_tp] = P.switch_pol * e.g * Pi;
  }
  if (P.tp_p >= 0) out[P.tp_p * stride_tp] = sqrt(6.) * e.g * Pi;
  if (P.tp_phi_plus_psi >= 0) out[P.tp_phi_plus_psi * stride_tp] = y[L.eta] + m.alpha_prime;
  if (P.tp_delta_m >= 0) out[P.tp_delta_m * stride_tp] = m.delta_m;
  if (P.tp_delta_cb >= 0) out[P.tp_delta_cb * stride_tp] = m.delta_cb;
}

// ---------------------------------------------------------------------------------------------
// dense LU of A = I - c J with partial pivoting (Crout-free right-looking form, rows over lanes)
__device__ bool lu_factor(Mode& M, int n, int ld, double c) {
  const int lane = M.lane;
  double* A = M.LU;
  // build A from the global Jacobian (column-major, coalesced)
  for (int j = 0; j < n; j++)
    for (int i = lane; i < n; i += 32) A[i + j * ld] = (i == j ? 1.0 : 0.0) - c * M.J[i + (size_t)j * n];
  __syncwarp();
  for (int j = 0; j < n; j++) {
    // pivot search in column j
    double best = -1.;
    int bi = j;
    for (int i = j + lane; i < n; i += 32) {
      const double v = fabs(A[i + j * ld]);
      if (v > best) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(PT_FULL, best, o);
      const int oi = __shfl_xor_sync(PT_FULL, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) M.piv[j] = bi;
    if (best == 0.) {
      if (lane == 0) A[j + j * ld] = 1e-50;  // TINY, as ludcmp does for a singular pivot
    }
    if (bi != j) {
      for (int cc = lane; cc < n; cc += 32) {
        const double t = A[j + cc * ld];
        A[j + cc * ld] = A[bi + cc * ld];
        A[bi + cc * ld] = t;
      }
    }
    __syncwarp();
    const double pinv = 1.0 / A[j + j * ld];
    for (int i = j + 1 + lane; i < n; i += 32) A[i + j * ld] *= pinv;
    __syncwarp();
    // trailing update, flat over the (n-j-1)^2 block
    const int mrem = n - j - 1;
    if (mrem > 0) {
      if (mrem >= 32) {
        for (int cc = j + 1; cc < n; cc++) {
          const double ajc = A[j + cc * ld];
          if (ajc != 0.)
            for (int i = j + 1 + lane; i < n; i += 32) A[i + cc * ld] -= A[i + j * ld] * ajc;
        }
      } else {
        const int tot = mrem * mrem;
        for (int t = lane; t < tot; t += 32) {
          const int cc = j + 1 + t / mrem, i = j + 1 + t % mrem;
          A[i + cc * ld] -= A[i + j * ld] * A[j + cc * ld];
        }
      }
    }
    __syncwarp();
  }
  return true;
}

// solve A x = b in place (b in shared memory)
__device__ void lu_solve(Mode& M, int n, int ld, double* b) {
  const int lane = M.lane;
  const double* A = M.LU;
  // apply the row permutation
  if (lane == 0) {
    for (int j = 0; j < n; j++) {
      const int p = M.piv[j];
      if (p != j) { const double t = b[j]; b[j] = b[p]; b[p] = t; }
    }
  }
  __syncwarp();
  // forward substitution (unit lower), column oriented
  for (int j = 0; j < n - 1; j++) {
    const double xj = b[j];
    if (xj != 0.)
      for (int i = j + 1 + lane; i < n; i += 32) b[i] -= A[i + j * ld] * xj;
    __syncwarp();
  }
  // back substitution
  for (int j = n - 1; j >= 0; j--) {
    if (lane == 0) b[j] = b[j] / A[j + j * ld];
    __syncwarp();
    const double xj = b[j];
    for (int i = lane; i < j; i += 32) b[i] -= A[i + j * ld] * xj;
    __syncwarp();
  }
}

// Jacobian J = A(tau): column j = f(tau, e_j) (the environment M.e must be set at tau)
__device__ void jacobian(const PtParams& P, Mode& M) {
  const int n = M.L.neq, lane = M.lane;
  double* e_j = M.tmp;
  double* col = M.del;
  for (int i = lane; i < n; i += 32) e_j[i] = 0.;
  __syncwarp();
  for (int j = 0; j < n; j++) {
    if (lane == 0) { e_j[j] = 1.; if (j > 0) e_j[j - 1] = 0.; }
    __syncwarp();
    rhs_apply<false>(P, M, e_j, col);
    for (int i = lane; i < n; i += 32) M.J[i + (size_t)j * n] = col[i];
    __syncwarp();
  }
  M.st.jacobians++;
  M.st.fevals += n;
}

// rescale the backward differences when the step changes by the factor r (k = current order)
__device__ void adjust_stepsize(Mode& M, double r, int k) {
  const double U[5][5] = {{-1, -2, -3, -4, -5}, {0, 1, 3, 6, 10}, {0, 0, -1, -4, -10}, {0, 0, 0, 1, 5}, {0, 0, 0, 0, -1}};
  double RU[5][5], tmpv[5];
  for (int ii = 1; ii <= 5; ii++) RU[0][ii - 1] = -ii * r;
  for (int jj = 2; jj <= 5; jj++)
    for (int ii = 1; ii <= 5; ii++) RU[jj - 1][ii - 1] = RU[jj - 2][ii - 1] * (1.0 - (1.0 + ii * r) / jj);
  for (int ii = 0; ii < 5; ii++) {
    for (int kk = 0; kk < 5; kk++) tmpv[kk] = RU[ii][kk];
    for (int jj = 0; jj < 5; jj++) {
      double s = 0.0;
      for (int kk = 0; kk < 5; kk++) s += tmpv[kk] * U[kk][jj];
      RU[ii][jj] = s;
    }
  }
  const int n = M.L.neq, np = M.neq_pad;
  for (int i = M.lane; i < n; i += 32) {
    double row[5];
    for (int kk = 0; kk < k; kk++) row[kk] = M.dif[kk * np + i];
    for (int jj = 0; jj < k; jj++) {
      double s = 0.0;
      for (int kk = 0; kk < k; kk++) s += row[kk] * RU[kk][jj];
      M.dif[jj * np + i] = s;
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// NDF1-5 over [t0, tfinal] for the current layout; y in M.y (in/out). `next` = index of the next
// source sample time (carried across intervals). Returns false on failure.
__device__ bool ndf15(const PtParams& P, Mode& M, double t0, double tfinal, int* next_io) {
  const double G[5] = {1.0, 3.0 / 2.0, 11.0 / 6.0, 25.0 / 12.0, 137.0 / 60.0};
  const double alpha[5] = {-37.0 / 200, -1.0 / 9.0, -8.23e-2, -4.15e-2, 0};
  double invGa[5], erconst[5];
  for (int i = 0; i < 5; i++) {
    invGa[i] = 1.0 / (G[i] * (1.0 - alpha[i]));
    erconst[i] = alpha[i] * G[i] + 1.0 / (2.0 + i);
  }
  const double abstol = 1e-15, eps = 1e-16, threshold = abstol;
  const int maxit = 4, maxk = 5;
  const double rtol = P.rtol;
  const int n = M.L.neq, np = M.neq_pad, ld = P.ld, lane = M.lane;
  double *y = M.y, *ynew = M.ynew, *f0 = M.f0, *pred = M.pred, *psi = M.psi, *difkp1 = M.difkp1, *del = M.del,
         *invwt = M.invwt, *dif = M.dif;
  const double* t_vec = P.tau;
  const int tres = P.tau_size;
  int next = *next_io;
  while (next < tres && t_vec[next] < t0) next++;

  for (int j = 0; j < 7; j++)
    for (int i = lane; i < n; i += 32) dif[j * np + i] = 0.;
  const double htspan = fabs(tfinal - t0);
  double t = t0, tnew = t0;
  env_at(P, M, t0, true);
  rhs_apply<false>(P, M, y, f0);
  M.st.fevals++;
  const double hmax = (tfinal - t0) / 10.0;
  jacobian(P, M);
  bool Jcurrent = true;
  double hmin = 16.0 * eps * fabs(t);
  // initial step from |f0/wt| and the second derivative estimate
  double rh = 0.0;
  for (int i = lane; i < n; i += 32) {
    const double wt = fmax(fabs(y[i]), threshold);
    M.tmp[i] = wt;
    rh = fmax(rh, 1.25 / sqrt(rtol) * fabs(f0[i] / wt));
  }
  rh = wmax(rh);
  double absh = fmin(hmax, htspan);
  if (absh * rh > 1.0) absh = 1.0 / rh;
  absh = fmax(absh, hmin);
  double h = absh;
  {
    const double tdel = (t + fmin(sqrt(eps) * fmax(fabs(t), fabs(t + h)), absh)) - t;
    env_at(P, M, t + tdel, true);
    rhs_apply<false>(P, M, y, del);  // f(t+tdel, y)
    M.st.fevals++;
    rh = 0.0;
    for (int i = lane; i < n; i += 32) {
      double s = 0.0;
      for (int j = 0; j < n; j++) s += M.J[i + (size_t)j * n] * f0[j];
      s += (del[i] - f0[i]) / tdel;
      rh = fmax(rh, 1.25 * sqrt(0.5 * fabs(s / M.tmp[i]) / rtol));
    }
    rh = wmax(rh);
    absh = fmin(hmax, htspan);
    if (absh * rh > 1.0) absh = 1.0 / rh;
    absh = fmax(absh, hmin);
    h = absh;
  }
  int k = 1, klast = k;
  double abshlast = absh;
  for (int i = lane; i < n; i += 32) dif[0 * np + i] = h * f0[i];
  __syncwarp();
  double hinvGak = h * invGa[k - 1];
  int nconhk = 0;
  lu_factor(M, n, ld, hinvGak);
  M.st.factorizations++;
  bool havrate = false;
  bool done = false, at_hmin = false;
  double rate = 0., oldnrm = 0., err = 0.;

  while (!done) {
    hmin = P.hmin_allowed;
    absh = fmin(hmax, fmax(hmin, absh));
    if (fabs(absh - hmin) < 100 * eps) {
      if (at_hmin) absh = abshlast;
      at_hmin = true;
    } else {
      at_hmin = false;
    }
    h = absh;
    if (1.1 * absh >= fabs(tfinal - t)) {
      h = tfinal - t;
      absh = fabs(h);
      done = true;
    }
    if (((fabs(absh - abshlast) / absh) > 1e-6) || (k != klast)) {
      adjust_stepsize(M, absh / abshlast, k);
      hinvGak = h * invGa[k - 1];
      nconhk = 0;
      lu_factor(M, n, ld, hinvGak);
      M.st.factorizations++;
      havrate = false;
    }
    bool nofailed = true;
    for (;;) {  // loop for advancing one step
      bool gotynew = false;
      while (!gotynew) {
        tnew = t + h;
        if (done) tnew = tfinal;
        h = tnew - t;
        double minnrm = 0.0;
        for (int i = lane; i < n; i += 32) {
          double ps = 0.0, pr = y[i];
          for (int j = 0; j < k; j++) {
            const double d = dif[j * np + i];
            ps += d * G[j] * invGa[k - 1];
            pr += d;
          }
          psi[i] = ps;
          pred[i] = pr;
          ynew[i] = pr;
          difkp1[i] = 0.0;
          const double iw = 1.0 / fmax(fmax(fabs(pr), fabs(y[i])), threshold);
          invwt[i] = iw;
          minnrm = fmax(minnrm, 100 * eps * fabs(pr * iw));
        }
        minnrm = wmax(minnrm);
        __syncwarp();
        env_at(P, M, tnew, true);
        bool tooslow = false;
        for (int iter = 1; iter <= maxit; iter++) {
          rhs_apply<false>(P, M, ynew, f0);
          if (!M.ap.tca_off) M.tca_shear_last = M.m.tca_shear_g;
          M.st.fevals++;
          for (int i = lane; i < n; i += 32) del[i] = hinvGak * f0[i] - (psi[i] + difkp1[i]);
          __syncwarp();
          lu_solve(M, n, ld, del);
          M.st.solves++;
          double newnrm = 0.0;
          for (int i = lane; i < n; i += 32) {
            newnrm = fmax(newnrm, fabs(del[i] * invwt[i]));
            difkp1[i] += del[i];
            ynew[i] = pred[i] + difkp1[i];
          }
          newnrm = wmax(newnrm);
          __syncwarp();
          if (newnrm <= minnrm) { gotynew = true; break; }
          else if (iter == 1) {
            if (havrate) {
              const double errit = newnrm * rate / (1.0 - rate);
              if (errit <= 0.05 * rtol) { gotynew = true; break; }
            } else {
              rate = 0.0;
            }
          } else if (newnrm > 0.9 * oldnrm) {
            tooslow = true;
            break;
          } else {
            rate = fmax(0.9 * rate, newnrm / oldnrm);
            havrate = true;
            const double errit = newnrm * rate / (1.0 - rate);
            if (errit <= 0.5 * rtol) { gotynew = true; break; }
            else if (iter == maxit) { tooslow = true; break; }
            else if (0.5 * rtol < errit * pow(rate, (double)(maxit - iter))) { tooslow = true; break; }
          }
          oldnrm = newnrm;
        }
        if (tooslow) {
          M.st.failed++;
          if (!Jcurrent) {
            env_at(P, M, t, true);
            rhs_apply<false>(P, M, y, f0);
            M.st.fevals++;
            jacobian(P, M);
            Jcurrent = true;
          } else if (absh <= hmin) {
            M.status = 2;  // step size too small
            return false;
          } else {
            abshlast = absh;
            absh = fmax(0.3 * absh, hmin);
            h = absh;
            done = false;
            adjust_stepsize(M, absh / abshlast, k);
            hinvGak = h * invGa[k - 1];
            nconhk = 0;
          }
          lu_factor(M, n, ld, hinvGak);
          M.st.factorizations++;
          havrate = false;
        }
      }
      // error estimate
      err = 0.0;
      for (int i = lane; i < n; i += 32) err = fmax(err, fabs(difkp1[i] * invwt[i]));
      err = wmax(err) * erconst[k - 1];
      if (err > rtol) {
        M.st.failed++;
        if (absh <= hmin) {
          M.status = 2;
          return false;
        }
        abshlast = absh;
        if (nofailed) {
          nofailed = false;
          double hopt = absh * fmax(0.1, 0.833 * pow((rtol / err), (1.0 / (k + 1))));
          if (k > 1) {
            double errkm1 = 0.0;
            for (int i = lane; i < n; i += 32) errkm1 = fmax(errkm1, fabs((dif[(k - 1) * np + i] + difkp1[i]) * invwt[i]));
            errkm1 = wmax(errkm1) * erconst[k - 2];
            const double hkm1 = absh * fmax(0.1, 0.769 * pow((rtol / errkm1), (1.0 / k)));
            if (hkm1 > hopt) {
              hopt = fmin(absh, hkm1);
              k = k - 1;
            }
          }
          absh = fmax(hmin, hopt);
        } else {
          absh = fmax(hmin, 0.5 * absh);
        }
        h = absh;
        if (absh < abshlast) done = false;
        adjust_stepsize(M, absh / abshlast, k);
        hinvGak = h * invGa[k - 1];
        nconhk = 0;
        lu_factor(M, n, ld, hinvGak);
        M.st.factorizations++;
        havrate = false;
      } else {
        break;
      }
    }
    M.st.steps++;
    // update the difference array
    for (int i = lane; i < n; i += 32) {
      dif[(k + 1) * np + i] = difkp1[i] - dif[k * np + i];
      dif[k * np + i] = difkp1[i];
      for (int j = k - 1; j >= 0; j--) dif[j * np + i] += dif[(j + 1) * np + i];
    }
    __syncwarp();
    // ---- output at the sample times passed by this step
    while ((next < tres) && ((tnew - t_vec[next]) >= 0.0)) {
      if (tnew == t_vec[next]) {
        write_sources(P, M, t_vec[next], ynew, f0, next);
      } else {
        const double s = (t_vec[next] - tnew) / h;
        double vecy[5], vecdy[5];
        double prod = 1.0, sumfrac = 0., fact = 1.0;
        for (int j = 0; j < k; j++) {
          prod *= (s + j);
          fact *= (j + 1);
          sumfrac += 1.0 / (s + j);
          vecy[j] = prod / fact;
          vecdy[j] = prod * sumfrac / (h * fact);
        }
        for (int i = lane; i < n; i += 32) {
          double a1 = 0, a2 = 0;
          for (int j = 0; j < k; j++) {
            a1 += vecy[j] * dif[j * np + i];
            a2 += vecdy[j] * dif[j * np + i];
          }
          M.yi[i] = ynew[i] + a1;
          M.ypi[i] = a2;
        }
        __syncwarp();
        write_sources(P, M, t_vec[next], M.yi, M.ypi, next);
      }
      next++;
    }
    if (done) break;
    klast = k;
    abshlast = absh;
    nconhk = min(nconhk + 1, maxk + 2);
    if (nconhk >= k + 2) {
      double temp = 1.2 * pow((err / rtol), (1.0 / (k + 1.0)));
      double hopt = (temp > 0.1) ? absh / temp : 10 * absh;
      int kopt = k;
      if (k > 1) {
        double errkm1 = 0.0;
        for (int i = lane; i < n; i += 32) errkm1 = fmax(errkm1, fabs(dif[(k - 1) * np + i] * invwt[i]));
        errkm1 = wmax(errkm1) * erconst[k - 2];
        temp = 1.3 * pow((errkm1 / rtol), (1.0 / k));
        const double hkm1 = (temp > 0.1) ? absh / temp : 10 * absh;
        if (hkm1 > hopt) { hopt = hkm1; kopt = k - 1; }
      }
      if (k < maxk) {
        double errkp1 = 0.0;
        for (int i = lane; i < n; i += 32) errkp1 = fmax(errkp1, fabs(dif[(k + 1) * np + i] * invwt[i]));
        errkp1 = wmax(errkp1) * erconst[k];
        temp = 1.4 * pow((errkp1 / rtol), (1.0 / (k + 2.0)));
        const double hkp1 = (temp > 0.1) ? absh / temp : 10 * absh;
        if (hkp1 > hopt) { hopt = hkp1; kopt = k + 1; }
      }
      if (hopt > absh) {
        absh = hopt;
        if (k != kopt) k = kopt;
      }
    }
    t = tnew;
    for (int i = lane; i < n; i += 32) y[i] = ynew[i];
    __syncwarp();
    Jcurrent = false;
  }
  // final state: y <- ynew, and a last RHS call so that the environment and the TCA/RSA
  // by-products are current at the end of the interval (evolver_ndf15.cpp:653-662)
  for (int i = lane; i < n; i += 32) y[i] = ynew[i];
  __syncwarp();
  env_at(P, M, tnew, true);
  rhs_apply<false>(P, M, y, f0);
  if (!M.ap.tca_off) M.tca_shear_last = M.m.tca_shear_g;
  M.st.fevals++;
  *next_io = next;
  return true;
}

// ---------------------------------------------------------------------------------------------
// perturb_initial_conditions: adiabatic mode, synchronous gauge, flat space
__device__ void initial_conditions(const PtParams& P, Mode& M, double tau) {
  env_at(P, M, tau, false);
  const Env& e = M.e;
  const Layout& L = M.L;
  const int lane = M.lane;
  const double k = M.k, a = e.a;
  double rho_r = e.rho_g, rho_m = e.rho_b + e.rho_cdm, rho_nu = 0.;
  if (P.has_ur) { rho_r += e.rho_ur; rho_nu += e.rho_ur; }
  for (int s = 0; s < P.N_ncdm; s++) { rho_r += M.pvb[P.irho_ncdm1 + s]; rho_nu += M.pvb[P.irho_ncdm1 + s]; }
  const double fracnu = rho_nu / rho_r;
  const double fracb = e.rho_b / rho_m;
  const double om = a * rho_m / sqrt(rho_r);
  const double ktau_two = k * k * tau * tau, ktau_three = k * tau * ktau_two;
  const double ci = P.curvature_ini;
  const double delta_g = -ktau_two / 3. * (1. - om * tau / 5.) * ci;
  const double theta_g = -k * ktau_three / 36. * (1. - 3. * (1. + 5. * fracb - fracnu) / 20. / (1. - fracnu) * om * tau) * ci;
  const double delta_ur = delta_g;
  const double theta_ur = -k * ktau_three / 36. / (4. * fracnu + 15.) *
                          (4. * fracnu + 11. + 12. - 3. * (8. * fracnu * fracnu + 50. * fracnu + 275.) / 20. / (2. * fracnu + 15.) * tau * om) * ci;
  const double shear_ur = ktau_two / (45. + 12. * fracnu) * (3. - 1.) * (1. + (4. * fracnu - 5.) / 4. / (2. * fracnu + 15.) * tau * om) * ci;
  const double l3_ur = ktau_three * 2. / 7. / (12. * fracnu + 45.) * ci;
  const double eta = ci * (1. - ktau_two / 12. / (15. + 4. * fracnu) *
                                    (5. + 4. * fracnu - (16. * fracnu * fracnu + 280. * fracnu + 325) / 10. / (2. * fracnu + 15.) * tau * om));
  for (int i = lane; i < L.neq; i += 32) M.y[i] = 0.;
  __syncwarp();
  if (lane == 0) {
    M.y[L.delta_g] = delta_g;
    M.y[L.theta_g] = theta_g;
    M.y[L.delta_b] = 3. / 4. * delta_g;
    M.y[L.theta_b] = theta_g;
    M.y[L.delta_cdm] = 3. / 4. * delta_g;
    M.y[L.eta] = eta;
    if (P.has_ur) {
      M.y[L.delta_ur] = delta_ur;
      M.y[L.theta_ur] = theta_ur;
      M.y[L.shear_ur] = shear_ur;
      M.y[L.l3_ur] = l3_ur;
    }
  }
  if (P.has_ncdm) {
    const int stride = L.l_max_ncdm + 1;
    for (int s = 0; s < P.N_ncdm; s++) {
      const double Ms = P.ncdm_M[s];
      for (int iq = lane; iq < L.q_size_ncdm[s]; iq += 32) {
        const int idx = L.ncdm_off[s] + iq * stride;
        const double q = P.ncdm_q[P.ncdm_q_off[s] + iq];
        const double dlnf0 = P.ncdm_dlnf0[P.ncdm_q_off[s] + iq];
        const double eps = sqrt(q * q + a * a * Ms * Ms);
        M.y[idx + 0] = -0.25 * delta_ur * dlnf0;
        M.y[idx + 1] = -eps / 3. / q / k * theta_ur * dlnf0;
        M.y[idx + 2] = -0.5 * shear_ur * dlnf0;
        M.y[idx + 3] = -0.25 * l3_ur * dlnf0;
      }
    }
  }
  __syncwarp();
}

// perturb_vector_init (switching part): move the state from the old layout (in M.y) to the new one
__device__ void remap_state(const PtParams& P, Mode& M, const Layout& Lo, const Approx& apo, const Layout& Ln,
                            const Approx& apn) {
  const int lane = M.lane;
  double* yo = M.y;
  double* yn = M.ynew;
  const double k = M.k;
  for (int i = lane; i < Ln.neq; i += 32) yn[i] = 0.;
  __syncwarp();
  if (lane == 0) {
    yn[Ln.delta_b] = yo[Lo.delta_b];
    yn[Ln.theta_b] = yo[Lo.theta_b];
    yn[Ln.delta_cdm] = yo[Lo.delta_cdm];
    yn[Ln.eta] = yo[Lo.eta];
    if (Ln.delta_g >= 0 && Lo.delta_g >= 0) { yn[Ln.delta_g] = yo[Lo.delta_g]; yn[Ln.theta_g] = yo[Lo.theta_g]; }
    if (Ln.delta_ur >= 0 && Lo.delta_ur >= 0) {
      yn[Ln.delta_ur] = yo[Lo.delta_ur]; yn[Ln.theta_ur] = yo[Lo.theta_ur]; yn[Ln.shear_ur] = yo[Lo.shear_ur];
    }
    if (Ln.shear_g >= 0 && Lo.shear_g < 0) {
      // tight coupling switched off: seed the hierarchy from the TCA expressions (:3909-3915);
      // tca_shear_g and kappa' are those of the last RHS call of the previous interval
      const double sg = M.m.tca_shear_g, dk = M.e.dkappa;
      yn[Ln.shear_g] = sg;
      yn[Ln.l3_g] = 6. / 7. * k / dk * sg;
      yn[Ln.pol0_g] = 2.5 * sg;
      yn[Ln.pol0_g + 1] = k / dk * (5. - 2.) / 6. * sg;
      yn[Ln.pol0_g + 2] = 0.5 * sg;
      yn[Ln.pol0_g + 3] = k / dk * 3. / 14. * sg;
    }
  }
  if (Ln.shear_g >= 0 && Lo.shear_g >= 0) {
    for (int l = 2 + lane; l <= Ln.l_max_g; l += 32) yn[Ln.delta_g + l] = yo[Lo.delta_g + l];
    for (int l = lane; l <= Ln.l_max_pol_g; l += 32) yn[Ln.pol0_g + l] = yo[Lo.pol0_g + l];
  }
  if (Ln.l3_ur >= 0 && Lo.l3_ur >= 0)
    for (int l = 3 + lane; l <= Ln.l_max_ur; l += 32) yn[Ln.delta_ur + l] = yo[Lo.delta_ur + l];
  if (P.has_ncdm) {
    if (apn.ncdmfa_on == apo.ncdmfa_on) {
      const int tot = Ln.eta - Ln.psi0_ncdm1;
      for (int i = lane; i < tot; i += 32) yn[Ln.psi0_ncdm1 + i] = yo[Lo.psi0_ncdm1 + i];
    } else {
      // ncdm fluid approximation switched on: integrate the momentum hierarchy (:4478-4518)
      const double a = M.e.a;
      const int stride = Lo.l_max_ncdm + 1;
      for (int s = 0; s < P.N_ncdm; s++) {
        const double rho_n = M.pvb[P.irho_ncdm1 + s], p_n = M.pvb[P.ip_ncdm1 + s];
        const double factor = P.ncdm_factor[s] * pow(P.a_today / a, 4);
        const double Ms = P.ncdm_M[s];
        double d = 0., th = 0., sh = 0.;
        for (int iq = lane; iq < Lo.q_size_ncdm[s]; iq += 32) {
          const int idx = Lo.ncdm_off[s] + iq * stride;
          const double q = P.ncdm_q[P.ncdm_q_off[s] + iq], w0 = P.ncdm_w[P.ncdm_q_off[s] + iq];
          const double eps = sqrt(q * q + a * a * Ms * Ms);
          d += w0 * pow(q, 2) * eps * yo[idx];
          th += w0 * pow(q, 3) * yo[idx + 1];
          sh += w0 * pow(q, 4) / eps * yo[idx + 2];
        }
        d = wsum(d) * factor / rho_n;
        th = wsum(th) * k * factor / (rho_n + p_n);
        sh = wsum(sh) * 2. / 3. * factor / (rho_n + p_n);
        if (lane == 0) {
          yn[Ln.ncdm_off[s]] = d;
          yn[Ln.ncdm_off[s] + 1] = th;
          yn[Ln.ncdm_off[s] + 2] = sh;
        }
      }
    }
  }
  __syncwarp();
  for (int i = lane; i < Ln.neq; i += 32) yo[i] = yn[i];
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) perturb_kernel(const PtParams P) {
  extern __shared__ double smem[];
  if ((int)blockIdx.x >= P.n_modes) return;
  Mode M;
  M.lane = threadIdx.x;
  const int np = (P.neq_max + 31) & ~31;
  M.neq_pad = np;
  double* p = smem;
  M.pvb = p; p += 32;
  M.pvt = p; p += 32;
  M.y = p; p += np; M.ynew = p; p += np; M.f0 = p; p += np; M.pred = p; p += np; M.psi = p; p += np;
  M.difkp1 = p; p += np; M.del = p; p += np; M.invwt = p; p += np; M.tmp = p; p += np; M.yi = p; p += np;
  M.ypi = p; p += np;
  M.dif = p; p += 7 * np;
  M.LU = p; p += (size_t)P.ld * P.neq_max;
  M.piv = (int*)p;
  M.J = P.jac + (size_t)blockIdx.x * P.neq_max * P.neq_max;
  M.ik = P.order[blockIdx.x];
  M.k = P.k[M.ik];
  M.k2 = M.k * M.k;
  M.cur_bg = 0; M.cur_th = P.tt_size - 2;
  M.status = 0;
  M.tca_shear_last = 0.;
  memset(&M.st, 0, sizeof(M.st));
  memset(&M.m, 0, sizeof(M.m));
  const int lane = M.lane;

  // ---- start time: bisection on (tau_c/tau_h, tau_h/tau_k, ncdm still relativistic)  (:2592-2635)
  double tau_lower = P.bg_tau[0], tau_upper = P.tau[0];
  {
    env_at(P, M, tau_lower, false);
    if (M.e.a * M.e.H / M.e.dkappa > P.start_small_k_at_tau_c_over_tau_h) M.status = 3;
    if (M.k / M.e.a / M.e.H > P.start_large_k_at_tau_h_over_tau_k) M.status = 4;
    for (int s = 0; s < P.N_ncdm; s++)
      if (fabs(M.pvb[P.ip_ncdm1 + s] / M.pvb[P.irho_ncdm1 + s] - 1. / 3.) > P.tol_ncdm_initial_w) M.status = 5;
  }
  double tau_mid = 0.5 * (tau_lower + tau_upper);
  if (M.status == 0) {
    while ((tau_upper - tau_lower) / tau_lower > P.tol_tau_approx) {
      env_at(P, M, tau_mid, false);
      bool early = true;
      for (int s = 0; s < P.N_ncdm; s++)
        if (fabs(M.pvb[P.ip_ncdm1 + s] / M.pvb[P.irho_ncdm1 + s] - 1. / 3.) > P.tol_ncdm_initial_w) early = false;
      if (early) {
        if ((M.e.a * M.e.H / M.e.dkappa > P.start_small_k_at_tau_c_over_tau_h) ||
            (M.k / M.e.a / M.e.H > P.start_large_k_at_tau_h_over_tau_k))
          early = false;
      }
      if (early) tau_lower = tau_mid; else tau_upper = tau_mid;
      tau_mid = 0.5 * (tau_lower + tau_upper);
    }
  }
  const double tau_ini = tau_mid;
  const double tau_end = P.tau[P.tau_size - 1];
  M.st.tau_ini = tau_ini;

  // ---- schedule of approximation switches (:2940-3231)
  double limit[PT_MAX_INTERVALS + 1];
  Approx sched[PT_MAX_INTERVALS];
  int n_int = 1;
  if (M.status == 0) {
    const Approx a_ini = approximations_at(P, M, tau_ini);
    const Approx a_end = approximations_at(P, M, tau_end);
    double sw[4];
    int nsw = 0;
    for (int w = 0; w < 4; w++) {
      const int f0 = approx_flag(a_ini, w), f1 = approx_flag(a_end, w);
      if (f1 < f0) { M.status = 6; break; }
      if (f1 > f0) {
        double lo = tau_ini, hi = tau_end, mid = 0.5 * (lo + hi);
        while (hi - lo > P.tol_tau_approx) {
          const Approx am = approximations_at(P, M, mid);
          if (approx_flag(am, w) > f0) hi = mid; else lo = mid;
          mid = 0.5 * (lo + hi);
        }
        sw[nsw++] = mid;
      }
    }
    n_int = nsw + 1;
    limit[0] = tau_ini;
    for (int i = 1; i < n_int; i++) {
      double nxt = tau_end;
      for (int j = 0; j < nsw; j++)
        if ((sw[j] > limit[i - 1]) && (sw[j] < nxt)) nxt = sw[j];
      limit[i] = nxt;
    }
    limit[n_int] = tau_end;
    sched[0] = a_ini;
    for (int i = 1; i < n_int && M.status == 0; i++) {
      sched[i] = approximations_at(P, M, 0.5 * (limit[i] + limit[i + 1]));
      int nchange = 0;
      for (int w = 0; w < 4; w++) {
        if (approx_flag(sched[i], w) < approx_flag(sched[i - 1], w)) M.status = 6;
        if (approx_flag(sched[i], w) != approx_flag(sched[i - 1], w)) nchange++;
      }
      if (nchange != 1) M.status = 7;
    }
    if (a_ini.tca_off || a_ini.rsa_on || a_ini.ufa_on || a_ini.ncdmfa_on) M.status = 8;
  }
  M.st.intervals = n_int;

  // ---- integrate interval by interval
  int next = 0;
  for (int iv = 0; iv < n_int && M.status == 0; iv++) {
    const Approx apn = sched[iv];
    const Layout Ln = make_layout(P, apn);
    if (iv == 0) {
      M.ap = apn;
      M.L = Ln;
      initial_conditions(P, M, limit[0]);
    } else {
      const Layout Lo = M.L;
      const Approx apo = M.ap;
      remap_state(P, M, Lo, apo, Ln, apn);
      M.ap = apn;
      M.L = Ln;
    }
    if (!ndf15(P, M, limit[iv], limit[iv + 1], &next)) break;
  }
  // zero-fill the samples that were not reached (failure only; normally next == tau_size)
  if (lane == 0) {
    const size_t stride_tp = (size_t)P.k_size * P.tau_size;
    double* out = P.sources + (size_t)M.ik * P.tau_size;
    const int tps[7] = {P.tp_t0, P.tp_t1, P.tp_t2, P.tp_p, P.tp_delta_m, P.tp_delta_cb, P.tp_phi_plus_psi};
    for (int it = next; it < P.tau_size; it++)
      for (int j = 0; j < 7; j++)
        if (tps[j] >= 0) out[tps[j] * stride_tp + it] = 0.;
    M.st.status = M.status;
    P.kstat[M.ik] = M.st;
  }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static int dev_alloc(T** p, size_t n, char* err) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  CLPP_CUDA(cudaMalloc((void**)p, n * sizeof(T)), err);
  return CLPP_SUCCESS;
}

int clpp_dev_perturb_solve(clpp_ctx* c, int k_begin, int k_end, char* err) {
  clpp_ctx::Dev* d = c->dev;
  const clpp_perturb_desc& pd = c->pd;
  const clpp_background_desc& bg = c->bg;
  const clpp_thermo_desc& th = c->th;
  const clpp_perturb_info& I = c->pinfo;
  cudaStream_t st = d->stream;
  CLPP_CHECK(c->N_ncdm <= PT_MAX_NCDM, err, "at most %d ncdm species are supported on the device", PT_MAX_NCDM);
  CLPP_CHECK(bg.bg_size_normal <= 32 && th.th_size <= 32, err, "background/thermo vectors wider than a warp");

  const int nk = I.k_size, nt = I.tau_size, ntp = I.tp_size;
  const size_t nsrc = (size_t)ntp * nk * nt;
  if (!d->sources || d->sources_count != nsrc) {
    if (dev_alloc(&d->sources, nsrc, err)) return CLPP_FAILURE;
    d->sources_count = nsrc;
    CLPP_CUDA(cudaMemsetAsync(d->sources, 0, nsrc * sizeof(double), st), err);
  }
  if (dev_alloc(&d->k, nk, err) || dev_alloc(&d->tau, nt, err) || dev_alloc(&d->kstat, nk, err)) return CLPP_FAILURE;
  CLPP_CUDA(cudaMemcpyAsync(d->k, c->k.data(), nk * sizeof(double), cudaMemcpyHostToDevice, st), err);
  CLPP_CUDA(cudaMemcpyAsync(d->tau, c->tau.data(), nt * sizeof(double), cudaMemcpyHostToDevice, st), err);
  CLPP_CUDA(cudaMemsetAsync(d->kstat, 0, nk * sizeof(clpp_kstat), st), err);

  // ncdm arrays
  double *d_q = nullptr, *d_w = nullptr, *d_dl = nullptr;
  if (c->N_ncdm > 0) {
    const size_t tot = c->ncdm_q.size();
    CLPP_CUDA(cudaMalloc((void**)&d_q, 3 * tot * sizeof(double)), err);
    d_w = d_q + tot;
    d_dl = d_w + tot;
    CLPP_CUDA(cudaMemcpyAsync(d_q, c->ncdm_q.data(), tot * sizeof(double), cudaMemcpyHostToDevice, st), err);
    CLPP_CUDA(cudaMemcpyAsync(d_w, c->ncdm_w.data(), tot * sizeof(double), cudaMemcpyHostToDevice, st), err);
    CLPP_CUDA(cudaMemcpyAsync(d_dl, c->ncdm_dlnf0.data(), tot * sizeof(double), cudaMemcpyHostToDevice, st), err);
  }

  PtParams P;
  memset(&P, 0, sizeof(P));
  P.bg_tau = d->bg_tau; P.bg_y = d->bg_y; P.bg_dd = d->bg_dd;
  P.bt_size = bg.bt_size; P.bg_size = bg.bg_size; P.bg_size_normal = bg.bg_size_normal;
  P.th_z = d->th_z; P.th_y = d->th_y; P.th_dd = d->th_dd;
  P.tt_size = th.tt_size; P.th_size = th.th_size;
  P.ia = bg.index_bg_a; P.iH = bg.index_bg_H; P.iHp = bg.index_bg_H_prime;
  P.irho_g = bg.index_bg_rho_g; P.irho_b = bg.index_bg_rho_b; P.irho_cdm = bg.index_bg_rho_cdm;
  P.irho_ur = bg.index_bg_rho_ur; P.irho_ncdm1 = bg.index_bg_rho_ncdm1; P.ip_ncdm1 = bg.index_bg_p_ncdm1;
  P.ipseudo_p_ncdm1 = bg.index_bg_pseudo_p_ncdm1;
  P.ixe = th.index_th_xe; P.idkappa = th.index_th_dkappa; P.iddkappa = th.index_th_ddkappa;
  P.idddkappa = th.index_th_dddkappa; P.iexp_m_kappa = th.index_th_exp_m_kappa; P.ig = th.index_th_g;
  P.idg = th.index_th_dg; P.iddg = th.index_th_ddg; P.icb2 = th.index_th_cb2; P.iwb = th.index_th_wb;
  P.iTb = th.index_th_Tb; P.itau_d = th.index_th_tau_d; P.irate = th.index_th_rate; P.ir_d = th.index_th_r_d;
  P.idcb2 = th.index_th_dcb2; P.iddcb2 = th.index_th_ddcb2;
  P.compute_cb2_derivatives = th.compute_cb2_derivatives; P.compute_damping_scale = th.compute_damping_scale;
  P.th_linear_below_z = -1.;
  if (th.reio_parametrization == CLPP_REIO_HALF_TANH) P.th_linear_below_z = 2 * th.z_reionization;
  if (th.reio_parametrization == CLPP_REIO_INTER) P.th_linear_below_z = 50.;
  P.n_e = th.n_e; P.YHe = th.YHe; P.T_cmb = bg.T_cmb; P.tau_free_streaming = th.tau_free_streaming;
  P.has_ur = bg.has_ur; P.has_ncdm = bg.has_ncdm; P.N_ncdm = bg.has_ncdm ? bg.N_ncdm : 0;
  int off = 0, nq_tot_l = 0;
  for (int s = 0; s < P.N_ncdm; s++) {
    P.ncdm_q_size[s] = c->ncdm_q_size[s];
    P.ncdm_q_off[s] = off;
    off += c->ncdm_q_size[s];
    P.ncdm_M[s] = c->ncdm_M[s];
    P.ncdm_factor[s] = c->ncdm_factor[s];
    nq_tot_l += c->ncdm_q_size[s] * (pd.l_max_ncdm + 1);
  }
  P.ncdm_q = d_q; P.ncdm_w = d_w; P.ncdm_dlnf0 = d_dl;
  P.a_today = bg.a_today;
  P.start_small_k_at_tau_c_over_tau_h = pd.start_small_k_at_tau_c_over_tau_h;
  P.start_large_k_at_tau_h_over_tau_k = pd.start_large_k_at_tau_h_over_tau_k;
  P.tca_trigger_tau_c_over_tau_h = pd.tight_coupling_trigger_tau_c_over_tau_h;
  P.tca_trigger_tau_c_over_tau_k = pd.tight_coupling_trigger_tau_c_over_tau_k;
  P.tca_method = pd.tight_coupling_approximation; P.rsa_method = pd.radiation_streaming_approximation;
  P.ufa_method = pd.ur_fluid_approximation; P.ncdmfa_method = pd.ncdm_fluid_approximation;
  P.rsa_trigger = pd.radiation_streaming_trigger_tau_over_tau_k; P.ufa_trigger = pd.ur_fluid_trigger_tau_over_tau_k;
  P.ncdmfa_trigger = pd.ncdm_fluid_trigger_tau_over_tau_k;
  P.l_max_g = pd.l_max_g; P.l_max_pol_g = pd.l_max_pol_g; P.l_max_ur = pd.l_max_ur; P.l_max_ncdm = pd.l_max_ncdm;
  P.tol_ncdm_initial_w = pd.tol_ncdm_initial_w; P.tol_tau_approx = pd.tol_tau_approx;
  P.rtol = pd.tol_perturb_integration; P.hmin_allowed = pd.smallest_allowed_variation;
  P.curvature_ini = pd.curvature_ini; P.three_ceff2_ur = pd.three_ceff2_ur; P.three_cvis2_ur = pd.three_cvis2_ur;
  P.switch_sw = pd.switch_sw; P.switch_eisw = pd.switch_eisw; P.switch_lisw = pd.switch_lisw;
  P.switch_dop = pd.switch_dop; P.switch_pol = pd.switch_pol; P.eisw_lisw_split_z = pd.eisw_lisw_split_z;
  P.k = d->k; P.k_size = nk; P.tau = d->tau; P.tau_size = nt; P.sources = d->sources;
  P.tp_t0 = I.index_tp_t0; P.tp_t1 = I.index_tp_t1; P.tp_t2 = I.index_tp_t2; P.tp_p = I.index_tp_p;
  P.tp_delta_m = I.index_tp_delta_m; P.tp_delta_cb = I.index_tp_delta_cb; P.tp_phi_plus_psi = I.index_tp_phi_plus_psi;

  // largest state vector over the approximation phases (full hierarchy, everything off)
  int neq_max = 2 + 1 + (pd.l_max_g - 2) + (pd.l_max_pol_g + 1) + 3 + 1;
  if (bg.has_ur) neq_max += 3 + (pd.l_max_ur - 2);
  neq_max += nq_tot_l;
  P.neq_max = neq_max;
  P.ld = neq_max | 1;
  const int np = (neq_max + 31) & ~31;
  const size_t smem = (size_t)(64 + 18 * np + (size_t)P.ld * neq_max) * sizeof(double) + (size_t)np * sizeof(int);
  CLPP_CHECK(smem <= 227 * 1024, err,
             "state vector of %d equations needs %zu bytes of shared memory per k-mode (> 227 KB): reduce l_max_ncdm / "
             "the number of ncdm momentum bins", neq_max, smem);

  // mode order: decreasing k (the expensive modes first)
  const int n_modes = k_end - k_begin;
  std::vector<int> order(n_modes);
  for (int i = 0; i < n_modes; i++) order[i] = k_end - 1 - i;
  if (dev_alloc(&d->k_order, std::max(n_modes, 1), err)) return CLPP_FAILURE;
  CLPP_CUDA(cudaMemcpyAsync(d->k_order, order.data(), n_modes * sizeof(int), cudaMemcpyHostToDevice, st), err);
  if (dev_alloc(&d->jac_scratch, (size_t)std::max(n_modes, 1) * neq_max * neq_max, err)) return CLPP_FAILURE;
  P.order = d->k_order; P.n_modes = n_modes; P.kstat = d->kstat; P.jac = d->jac_scratch;

  CLPP_CUDA(cudaFuncSetAttribute(perturb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), err);
  cudaEventRecord(d->ev[0], st);
  if (n_modes > 0) {
    perturb_kernel<<<n_modes, 32, smem, st>>>(P);
    c->launches++;
  }
  cudaEventRecord(d->ev[1], st);
  CLPP_CUDA(cudaGetLastError(), err);
  c->kstat.assign(nk, clpp_kstat{});
  CLPP_CUDA(cudaMemcpyAsync(c->kstat.data(), d->kstat, nk * sizeof(clpp_kstat), cudaMemcpyDeviceToHost, st), err);
  CLPP_CUDA(cudaStreamSynchronize(st), err);
  {
    float ms = 0;
    cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]);
    d->t_perturb_ms = ms;
  }
  if (d_q) cudaFree(d_q);
  for (int ik = k_begin; ik < k_end; ik++) {
    const int s = c->kstat[ik].status;
    if (s != 0) {
      const char* what = s == 2 ? "Step size too small in the NDF15 evolver"
                       : s == 3 ? "your choice of initial time for integrating wavenumbers is inappropriate: it corresponds to a time before that at which the background has been integrated. You should increase 'start_small_k_at_tau_c_over_tau_h'"
                       : s == 4 ? "your choice of initial time for integrating wavenumbers is inappropriate: it corresponds to a time before that at which the background has been integrated. You should increase 'start_large_k_at_tau_h_over_tau_k'"
                       : s == 5 ? "your choice of initial time for integrating wavenumbers is inappropriate: ncdm species not ultra-relativistic"
                       : s == 6 ? "an approximation flag goes backward in time, this cannot be handled"
                       : s == 7 ? "you switch several approximations at the same time, this cannot be handled"
                       : s == 8 ? "scalar initial conditions assume tight coupling on and all other approximations off"
                                : "unknown device error";
      return clpp_fail(err, "perturb_solve failed for k=%e (index %d): %s", c->k[ik], ik, what);
    }
  }
  c->has_sources = true;
  return CLPP_SUCCESS;
}
