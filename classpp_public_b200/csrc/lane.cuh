// Stage 1, lane kernels: ONE THREAD integrates one (cosmology, k) mode from its initial time to today.
//
// A warp therefore advances 32 modes at once -- the same k of 32 neighbouring cosmologies of a sweep, or 32 neighbouring
// k of one cosmology -- and every instruction does useful work for 32 modes (the warp-per-mode kernels of perturb.cu
// evaluate all scalar "hub" mathematics once per warp).  The per-mode state (state vector, NDF backward differences,
// Newton-matrix factors; 20..100 KB) lives in a per-thread slab of GLOBAL memory laid out element-major
// (element e of the 64 threads of a CTA are 512 contiguous bytes), so every access of a warp is coalesced; the
// working set streams through L1/L2.  The time stepping is a flat state machine -- one step ATTEMPT per loop
// iteration, rejected attempts simply come round again -- so that lanes whose steps fail or whose Newton iteration
// needs a new Jacobian do not stall the other 31 lanes of the warp.
//
// The physics and the integrator are the ones of perturb.cu (same reference functions, cited there):
//   perturb_solve / perturb_find_approximation_switches / perturb_approximations  (perturbations_module.cpp:2463-3231, 5443-5670)
//   perturb_vector_init / perturb_initial_conditions (:3271-5408), perturb_einstein (:5840), perturb_total_stress_energy (:6047),
//   perturb_sources_member (:6731), perturb_derivs_member (:7861), perturb_tca_slip_and_shear (:9229),
//   perturb_rsa_delta_and_theta (:9530), evolver_ndf15 (tools/evolver_ndf15.cpp:62-705).
// Own design (not in the reference): state vector ordered HUB FIRST (densities, velocities, shears, the l <= 2 moments
// of every hierarchy, eta), then the multipole CHAINS (l >= 3 of photon temperature, photon polarisation, ur, each ncdm
// momentum bin).  I - cJ is factorised as chains -> Schur complement on their l = 2 roots -> hub block.  The hub block
// itself is D + U V^T: D block diagonal (photon-baryon plasma, ur, one 3x3 block per ncdm momentum bin), the rest is
// the coupling through the four metric scalars (h', eta', alpha', eta), so a Newton solve costs O(n) instead of O(nh^2)
// and a refactorisation O(n) instead of O(nh^3) (ln_factor / ln_solve).
//
// The file compiles for the device (included by perturb.cu) and, with -DCLPP_HOST_SIM, as plain C++ for the
// developer harness tests/hostsim (CPU execution of the SAME lane program for debugging; test infrastructure, never
// linked into libclpp.so).
#ifndef CLPP_LANE_CUH
#define CLPP_LANE_CUH

#include "pt_types.h"

#ifndef LN_CTA
#define LN_CTA 64
#endif

#ifdef CLPP_HOST_SIM
#include <cmath>
#define LN_FN static inline
#define LN_NOINLINE static
#define LN_STRIDE 1
#define LN_LDG(p) (*(p))
#define LN_CLOCK() 0LL
static const double c_G[5] = {1.0, 3.0 / 2.0, 11.0 / 6.0, 25.0 / 12.0, 137.0 / 60.0};
static const double c_invGa[5] = {1.0 / (1.0 * (1.0 + 37.0 / 200)), 1.0 / (1.5 * (1.0 + 1.0 / 9.0)),
                                  1.0 / (11.0 / 6.0 * (1.0 + 8.23e-2)), 1.0 / (25.0 / 12.0 * (1.0 + 4.15e-2)),
                                  1.0 / (137.0 / 60.0)};
static const double c_erconst[5] = {-37.0 / 200 * 1.0 + 1.0 / 2.0, -1.0 / 9.0 * 1.5 + 1.0 / 3.0,
                                    -8.23e-2 * (11.0 / 6.0) + 1.0 / 4.0, -4.15e-2 * (25.0 / 12.0) + 1.0 / 5.0, 1.0 / 6.0};
static const double c_invint[7] = {0., 1.0, 0.5, 1.0 / 3.0, 0.25, 0.2, 1.0 / 6.0};
static const double c_U[5][5] = {{-1, -2, -3, -4, -5}, {0, 1, 3, 6, 10}, {0, 0, -1, -4, -10}, {0, 0, 0, 1, 5}, {0, 0, 0, 0, -1}};
static inline double ln_root_n(double x, double n) { return (double)exp2f(log2f((float)x) / (float)n); }
#else
#define LN_FN __device__ __forceinline__
#define LN_NOINLINE __device__ __noinline__
#define LN_STRIDE LN_CTA
#define LN_LDG(p) __ldg(p)
#define LN_CLOCK() clock64()
// x^(1/n) for the step-size heuristics: float accuracy is ample for a controller that only compares and clamps the result
__device__ __forceinline__ double ln_root_n(double x, double n) {
  const float xf = fminf(fmaxf((float)x, 1e-30f), 1e30f);
  return (double)exp2f(__log2f(xf) / (float)n);
}
#endif

// ---- per-thread scratch (global memory, element-major)
#define LM(off) mem[(size_t)(off) * LN_STRIDE]
enum {
  LV_Y = 0, LV_YNEW, LV_F, LV_PRED, LV_PSI, LV_DIFKP1, LV_DEL, LV_INVWT,
  LV_TMP = LV_PSI, LV_YPI = LV_PRED,  // source output / Jacobian probes: psi and pred are dead there
  LV_DIF0 = LV_INVWT + 1,             // 7 slots: dif[0..6]
  LV_JD = LV_DIF0 + 7, LV_JL, LV_JU,  // chain rows of J: diagonal, coupling to l-1 (to the root for the first element), to l+1
  LV_IP, LV_MU,                       // chain factors: 1/pivot, T[i,i+1]/p[i+1]
  LV_COUNT
};
#define LVEC(slot, i) LM(P.lo_vec + (slot) * P.np + (i))
// per-chain scalars: J[root, first], its eliminated multiplier
#define LCH_JUR(c) LM(P.lo_ch + (c))
#define LCH_MUR(c) LM(P.lo_ch + PT_MAX_CHAINS + (c))
// metric scalars the hub rows depend on (the low-rank part of the hub Jacobian)
enum { MS_HP = 0, MS_EP, MS_AP, MS_ETA, MS_COUNT };
#define LN_BS_MAX 8  // largest diagonal block of the hub (photon-baryon plasma: delta_g theta_g shear_g pol0 pol1 pol2 delta_b theta_b)

struct LnLayout {
  int neq, nh, nch;
  int delta_g, theta_g, shear_g, pol0_g;  // pol0, pol1, pol2 are consecutive hub variables; -1 when absent
  int delta_b, theta_b, delta_cdm;
  int delta_ur, theta_ur, shear_ur;
  int psi0_ncdm1, nbin;  // ncdm triples (l = 0, 1, 2 of each momentum bin, or delta/theta/shear of each species in the fluid approx.)
  int eta;
  int c_g, c_pol, c_ur, c_ncdm;  // first element (l = 3) of each chain family in the state vector, -1 when absent
  int len_g, len_pol, len_ur, len_ncdm;
};

struct LnEnv {
  double tau, a, H, Hp;
  double rho_g, rho_b, rho_cdm, rho_ur;
  double rho_n[PT_MAX_NCDM], p_n[PT_MAX_NCDM], pp_n[PT_MAX_NCDM];
  double dkappa, ddkappa, cb2;
  double g, dg, exp_m_kappa;  // filled on request (sources)
  double R, inv_R, inv_1pR, inv_half_aH, inv_tau, tau_c, fac_ncdm;
  double nf[PT_MAX_NCDM][8];  // ncdm fluid constants
};

struct LnMetric {
  double h_prime, eta_prime, alpha, alpha_prime;
  double rsa_delta_g, rsa_theta_g;
  double delta_m, delta_cb;
  double tca_shear_g;
};

struct LnStat {
  int steps, failed, fevals, jacobians, factorizations, solves;
};

struct Lane {
  const PtCosmo* C;
  const double *bg_tau, *bg_y, *bg_dd, *th_z, *th_y, *th_dd;
  double k, k2, ik, ik2;
  double z_last, th_lin, a_today;
  double fac_c;
  double tca_shear_last;
  int bt_size, tt_size, bg_cur, th_cur;
  int ik_index, need_nw, status, next;
  LnEnv e;
  LnMetric m;
  Approx ap, apprev;
  LnLayout L, Lprev;
  LnStat st;
};

enum { LR_MATTER = 1, LR_HUB = 2, LR_CHAINS = 4, LR_GIVEN_METRIC = 8 };

// -------------------------------------------------------------------------------------------------
// table lookups: largest inf <= n-2 with X[inf] <= x (X growing); cursor walk first (array_interpolate_spline_growing_closeby,
// arrays.c:2173-2232), bisection when the target is far
LN_FN int ln_locate(const double* __restrict__ X, int n, double x, int cur) {
  int inf = cur < 0 ? 0 : (cur > n - 2 ? n - 2 : cur);
  int guard = 0;
  bool far_away = false;
  while (inf > 0 && x < LN_LDG(X + inf)) {
    inf--;
    if (++guard > 8) { far_away = true; break; }
  }
  if (!far_away) {
    int sup = inf + 1;
    while (sup < n - 1 && x > LN_LDG(X + sup)) {
      sup++;
      if (++guard > 8) { far_away = true; break; }
    }
    if (!far_away) return sup - 1;
  }
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (x >= LN_LDG(X + mid)) lo = mid; else hi = mid;
  }
  return lo;
}

// background_at_tau (the columns the path needs) + thermodynamics_at_z + everything that only depends on time.
// want_src: also the visibility function columns (source output).
LN_NOINLINE void ln_env(const PtParams& P, Lane& M, double* __restrict__ mem, double tau, int want_src) {
  LnEnv& e = M.e;
  // ---- background
  {
    const int inf = ln_locate(M.bg_tau, M.bt_size, tau, M.bg_cur);
    M.bg_cur = inf;
    const double x0 = LN_LDG(M.bg_tau + inf), x1 = LN_LDG(M.bg_tau + inf + 1);
    const double h = x1 - x0, ih = 1.0 / h, h26 = h * h / 6.;
    const double b = (tau - x0) * ih, a = 1 - b;
    const double ca = (a * a * a - a), cb = (b * b * b - b);
    const double* __restrict__ y0 = M.bg_y + (size_t)inf * P.bg_size;
    const double* __restrict__ y1 = y0 + P.bg_size;
    const double* __restrict__ d0 = M.bg_dd + (size_t)inf * P.bg_size;
    const double* __restrict__ d1 = d0 + P.bg_size;
#define LN_BG(col) (a * LN_LDG(y0 + (col)) + b * LN_LDG(y1 + (col)) + (ca * LN_LDG(d0 + (col)) + cb * LN_LDG(d1 + (col))) * h26)
    e.a = LN_BG(P.ia); e.H = LN_BG(P.iH); e.Hp = LN_BG(P.iHp);
    e.rho_g = LN_BG(P.irho_g); e.rho_b = LN_BG(P.irho_b); e.rho_cdm = LN_BG(P.irho_cdm);
    e.rho_ur = P.has_ur ? LN_BG(P.irho_ur) : 0.;
    for (int s = 0; s < P.N_ncdm; s++) {
      e.rho_n[s] = LN_BG(P.irho_ncdm1 + s); e.p_n[s] = LN_BG(P.ip_ncdm1 + s); e.pp_n[s] = LN_BG(P.ipseudo_p_ncdm1 + s);
    }
#undef LN_BG
  }
  e.tau = tau;
  const double av = e.a, Hv = e.H, Hp = e.Hp;
  const double inv_a = 1. / av;
  const double z = inv_a - 1.;
  // ---- thermodynamics
  if (z >= M.z_last) {
    const PtCosmo* C = M.C;
    const double* row = M.th_y + (size_t)(M.tt_size - 1) * P.th_size;
    const double xe0 = LN_LDG(row + P.ixe);
    e.dkappa = (1. + z) * (1. + z) * C->n_e * xe0 * CLPP_sigma * CLPP_Mpc_over_m;
    e.ddkappa = -Hv * 2. / (1. + z) * e.dkappa;
    e.exp_m_kappa = 0.; e.g = 0.; e.dg = 0.;
    const double wb = CLPP_k_B / (CLPP_c * CLPP_c * CLPP_m_H) * (1. + (1. / CLPP_not4 - 1.) * C->YHe + xe0 * (1. - C->YHe)) *
                      C->T_cmb * (1. + z);
    e.cb2 = wb * 4. / 3.;
  } else {
    const bool linear = (z < M.th_lin);
    const int inf = ln_locate(M.th_z, M.tt_size, z, M.th_cur);
    M.th_cur = inf;
    const double x0 = LN_LDG(M.th_z + inf), x1 = LN_LDG(M.th_z + inf + 1);
    const double h = x1 - x0, ih = 1.0 / h, h26 = linear ? 0. : h * h / 6.;
    const double b = (z - x0) * ih, a = 1 - b;
    const double ca = (a * a * a - a), cb = (b * b * b - b);
    const double* __restrict__ y0 = M.th_y + (size_t)inf * P.th_size;
    const double* __restrict__ y1 = y0 + P.th_size;
    const double* __restrict__ d0 = M.th_dd + (size_t)inf * P.th_size;
    const double* __restrict__ d1 = d0 + P.th_size;
#define LN_TH(col) (a * LN_LDG(y0 + (col)) + b * LN_LDG(y1 + (col)) + (ca * LN_LDG(d0 + (col)) + cb * LN_LDG(d1 + (col))) * h26)
    e.dkappa = LN_TH(P.idkappa); e.ddkappa = LN_TH(P.iddkappa); e.cb2 = LN_TH(P.icb2);
    if (want_src) { e.g = LN_TH(P.ig); e.dg = LN_TH(P.idg); e.exp_m_kappa = LN_TH(P.iexp_m_kappa); }
#undef LN_TH
  }
  // ---- derived quantities (every division by a time-only quantity is done here, once per step)
  const double aH = Hv * av;
  e.inv_R = 0.75 * e.rho_b / e.rho_g;
  e.R = 4. / 3. * e.rho_g / e.rho_b;
  e.inv_1pR = e.rho_b / (e.rho_b + 4. / 3. * e.rho_g);
  e.inv_half_aH = 2. / aH;
  e.inv_tau = 1. / tau;
  e.tau_c = 1. / e.dkappa;
  {
    const double a_rel = M.a_today * inv_a;
    e.fac_ncdm = (a_rel * a_rel) * (a_rel * a_rel);
  }
  if (P.has_ncdm && M.ap.ncdmfa_on) {
    for (int s = 0; s < P.N_ncdm; s++) {
      const double rho_n = e.rho_n[s], p_n = e.p_n[s], pseudo = e.pp_n[s];
      const double w_n = p_n / rho_n, pseudo_p_over_p = pseudo / p_n, i1w = rho_n / (rho_n + p_n), inv_w = rho_n / p_n;
      const double cg2 = w_n * (1.0 - i1w * (1. / 3.) * (3.0 * w_n - 2.0 + pseudo_p_over_p));
      const double ca2 = w_n * (1. / 3.) * i1w * (5.0 - pseudo_p_over_p);
      const double cvis2 = (P.ncdmfa_method == CLPP_NCDMFA_HU) ? w_n : 3. * w_n * ca2;
      const double damp = (P.ncdmfa_method == CLPP_NCDMFA_HU) ? 3.0 * aH * ca2 * inv_w
                                                              : 3.0 * (aH * (2. / 3. - ca2 - pseudo_p_over_p * (1. / 3.)) + 1.0 / tau);
      double* nf = e.nf[s];
      nf[0] = rho_n; nf[1] = rho_n + p_n; nf[2] = w_n; nf[3] = cg2 * rho_n; nf[4] = ca2; nf[5] = ca2 * i1w;
      nf[6] = 8.0 / 3.0 * cvis2 * i1w; nf[7] = damp;
    }
  }
  // momentum-dependent ncdm weights at this scale factor
  if (M.need_nw) {
    const double a2 = av * av;
    const int nq = P.nq_tot;
    for (int s = 0; s < P.N_ncdm; s++) {
      const double Ms = M.C->ncdm_M[s];
      for (int j = P.ncdm_q_off[s]; j < P.ncdm_q_off[s] + P.ncdm_q_size[s]; j++) {
        const double q = LN_LDG(M.C->ncdm_q + j), w0 = LN_LDG(M.C->ncdm_w + j);
        const double q2 = q * q, eps = sqrt(q2 + Ms * Ms * a2);
        const double ieps = 1.0 / eps;
        LM(P.lo_nw + j) = q * ieps;
        LM(P.lo_nw + nq + j) = q2 * eps * w0;
        LM(P.lo_nw + 2 * nq + j) = q2 * q * w0;
        LM(P.lo_nw + 3 * nq + j) = q2 * q2 * ieps * w0;
      }
    }
  }
}

// perturb_approximations: flags at time tau
LN_NOINLINE Approx ln_approximations_at(const PtParams& P, Lane& M, double* mem, double tau) {
  M.bg_cur = -1000000; M.th_cur = -1000000;  // inter_normal: bisection
  {
    // a cursor far outside forces the bisection path of ln_locate without a second code copy
    M.bg_cur = M.bt_size / 2; M.th_cur = M.tt_size / 2;
  }
  ln_env(P, M, mem, tau, 0);
  const LnEnv& e = M.e;
  Approx a;
  const double tau_k = 1. / M.k, tau_h = 1. / (e.H * e.a);
  const double dkappa = e.dkappa;
  if (dkappa == 0.) a.tca_off = 1;
  else {
    const double tau_c = 1. / dkappa;
    a.tca_off = ((tau_c / tau_h < P.tca_trigger_tau_c_over_tau_h) && (tau_c / tau_k < P.tca_trigger_tau_c_over_tau_k)) ? 0 : 1;
  }
  a.rsa_on = ((tau / tau_k > P.rsa_trigger) && (tau > M.C->tau_free_streaming) && (P.rsa_method != CLPP_RSA_NONE)) ? 1 : 0;
  a.ufa_on = (P.has_ur && (tau / tau_k > P.ufa_trigger) && (P.ufa_method != CLPP_UFA_NONE)) ? 1 : 0;
  a.ncdmfa_on = (P.has_ncdm && (tau / tau_k > P.ncdmfa_trigger) && (P.ncdmfa_method != CLPP_NCDMFA_NONE)) ? 1 : 0;
  return a;
}

LN_FN int ln_approx_flag(const Approx& a, int which) {
  return which == 0 ? a.tca_off : which == 1 ? a.rsa_on : which == 2 ? a.ufa_on : a.ncdmfa_on;
}

// state-vector layout for a set of approximations: hub first, then the chains
LN_FN void ln_make_layout(const PtParams& P, const Approx& ap, LnLayout& L) {
  int n = 0;
  L.delta_g = L.theta_g = L.shear_g = L.pol0_g = -1;
  L.delta_ur = L.theta_ur = L.shear_ur = -1;
  L.psi0_ncdm1 = -1; L.nbin = 0;
  L.c_g = L.c_pol = L.c_ur = L.c_ncdm = -1;
  L.len_g = L.len_pol = L.len_ur = L.len_ncdm = 0;
  const bool full_g = !ap.rsa_on && ap.tca_off;
  if (!ap.rsa_on) {
    L.delta_g = n++; L.theta_g = n++;
    if (ap.tca_off) { L.shear_g = n++; L.pol0_g = n; n += 3; }
  }
  L.delta_b = n++; L.theta_b = n++;
  L.delta_cdm = n++;
  const bool has_ur = P.has_ur && !ap.rsa_on;
  if (has_ur) { L.delta_ur = n++; L.theta_ur = n++; L.shear_ur = n++; }
  if (P.has_ncdm) {
    L.psi0_ncdm1 = n;
    L.nbin = ap.ncdmfa_on ? P.N_ncdm : P.nq_tot;
    n += 3 * L.nbin;
  }
  L.eta = n++;
  L.nh = n;
  int nch = 0;
  if (full_g) {
    L.c_g = n; L.len_g = P.l_max_g - 2; n += L.len_g;
    L.c_pol = n; L.len_pol = P.l_max_pol_g - 2; n += L.len_pol;
    nch += 2;
  }
  if (has_ur && !ap.ufa_on) { L.c_ur = n; L.len_ur = P.l_max_ur - 2; n += L.len_ur; nch++; }
  if (P.has_ncdm && !ap.ncdmfa_on) { L.c_ncdm = n; L.len_ncdm = P.l_max_ncdm - 2; n += L.len_ncdm * P.nq_tot; nch += P.nq_tot; }
  L.nch = nch;
  L.neq = n;
}

// chain c: first element, length, hub slot of its root (the l = 2 moment)
LN_FN void ln_chain(const LnLayout& L, int c, int& start, int& len, int& root) {
  if (L.c_g >= 0) {
    if (c == 0) { start = L.c_g; len = L.len_g; root = L.shear_g; return; }
    if (c == 1) { start = L.c_pol; len = L.len_pol; root = L.pol0_g + 2; return; }
    c -= 2;
  }
  if (L.c_ur >= 0) {
    if (c == 0) { start = L.c_ur; len = L.len_ur; root = L.shear_ur; return; }
    c -= 1;
  }
  start = L.c_ncdm + c * L.len_ncdm; len = L.len_ncdm; root = L.psi0_ncdm1 + 3 * c + 2;
}

// -------------------------------------------------------------------------------------------------
// Right-hand side f(tau, y) for the environment in M.e.  Fills M.m (metric and by-products).
//  LR_MATTER: also delta_m, delta_cb.  LR_HUB / LR_CHAINS: which rows of dy to write (none: metric only).
//  LR_GIVEN_METRIC: the four metric scalars ms[] are imposed instead of being computed from y (Jacobian probes:
//  the hub block is D + U V^T with V^T = d(metric)/dy, U = d(rows)/d(metric), D = d(rows)/dy at fixed metric).
LN_NOINLINE void ln_rhs(const PtParams& P, Lane& M, double* __restrict__ mem, int sy, int sdy, int flags, const double* ms_in) {
  const LnLayout& L = M.L;
  const Approx ap = M.ap;
  const LnEnv& e = M.e;
  const double k = M.k, k2 = M.k2, ik2 = M.ik2;
#define Y(i) LVEC(sy, i)
#define DY(i) LVEC(sdy, i)
  const double a2 = e.a * e.a, aH = e.H * e.a, R = e.R;
  const double rho_g = e.rho_g, rho_b = e.rho_b, rho_cdm = e.rho_cdm, rho_ur = e.rho_ur;
  const double dkappa = e.dkappa, cb2 = e.cb2;
  LnMetric& m = M.m;
  const bool has_g = !ap.rsa_on;
  const bool has_ur = P.has_ur && !ap.rsa_on;
  const bool full_g = has_g && ap.tca_off;

  double delta_g = 0., theta_g = 0., shear_g = 0.;
  if (has_g) { delta_g = Y(L.delta_g); theta_g = Y(L.theta_g); }
  double g3 = 0., p0 = 0., p1 = 0., p2 = 0., p3 = 0.;
  if (full_g) {
    shear_g = Y(L.shear_g); g3 = Y(L.c_g);
    p0 = Y(L.pol0_g); p1 = Y(L.pol0_g + 1); p2 = Y(L.pol0_g + 2); p3 = Y(L.c_pol);
  }
  double delta_ur = 0., theta_ur = 0., shear_ur = 0., u3 = 0.;
  if (has_ur) {
    delta_ur = Y(L.delta_ur); theta_ur = Y(L.theta_ur); shear_ur = Y(L.shear_ur);
    if (!ap.ufa_on) u3 = Y(L.c_ur);
  }
  const double delta_b = Y(L.delta_b), theta_b = Y(L.theta_b), delta_cdm = Y(L.delta_cdm), eta_y = Y(L.eta);
  const double delta_p_b_over_rho_b = cb2 * delta_b;

  // ---- perturb_total_stress_energy
  double delta_rho = rho_g * delta_g + rho_b * delta_b;
  double rpt = 4. / 3. * rho_g * theta_g + rho_b * theta_b;
  double rps = 4. / 3. * rho_g * shear_g;
  double delta_rho_m = rho_b * delta_b, rho_m = rho_b, rpt_m = rho_b * theta_b, rpm = rho_b;
  delta_rho += rho_cdm * delta_cdm;
  delta_rho_m += rho_cdm * delta_cdm; rho_m += rho_cdm; rpm += rho_cdm;
  if (P.has_ur) {
    delta_rho = delta_rho + rho_ur * delta_ur;
    rpt = rpt + 4. / 3. * rho_ur * theta_ur;
    rps = rps + 4. / 3. * rho_ur * shear_ur;
  }
  if (flags & LR_MATTER) m.delta_cb = delta_rho_m / rho_m + 3. * aH * (rpt_m / rpm) * ik2;
  if (P.has_ncdm) {
    if (ap.ncdmfa_on) {
      for (int s = 0; s < P.N_ncdm; s++) {
        const double* nf = e.nf[s];
        const int idx = L.psi0_ncdm1 + 3 * s;
        const double y0 = Y(idx), y1 = Y(idx + 1), y2 = Y(idx + 2);
        delta_rho += nf[0] * y0;
        rpt += nf[1] * y1;
        rps += nf[1] * y2;
        delta_rho_m += nf[0] * y0; rho_m += nf[0];
        rpt_m += nf[1] * y1; rpm += nf[1];
      }
    } else {
      const int nqt = P.nq_tot;
      for (int s = 0; s < P.N_ncdm; s++) {
        const double rho_n = e.rho_n[s], p_n = e.p_n[s];
        const double factor = M.C->ncdm_factor[s] * e.fac_ncdm;
        double s_rho = 0., s_theta = 0., s_shear = 0.;
        const int nq = P.ncdm_q_size[s], q0 = P.ncdm_q_off[s];
        for (int iq = 0; iq < nq; iq++) {
          const int idx = L.psi0_ncdm1 + 3 * (q0 + iq);
          const double y0 = Y(idx), y1 = Y(idx + 1), y2 = Y(idx + 2);
          s_rho += LM(P.lo_nw + nqt + q0 + iq) * y0;
          s_theta += LM(P.lo_nw + 2 * nqt + q0 + iq) * y1;
          s_shear += LM(P.lo_nw + 3 * nqt + q0 + iq) * y2;
        }
        s_rho *= factor;
        s_theta *= k * factor;
        s_shear *= 2.0 / 3.0 * factor;
        delta_rho += s_rho; rpt += s_theta; rps += s_shear;
        delta_rho_m += s_rho; rho_m += rho_n;
        rpt_m += s_theta; rpm += (rho_n + p_n);
      }
    }
  }
  if (flags & LR_MATTER) m.delta_m = delta_rho_m / rho_m + 3. * aH * (rpt_m / rpm) * ik2;

  // ---- perturb_einstein (synchronous gauge, K = 0)
  double h_prime = (k2 * eta_y + 1.5 * a2 * delta_rho) * e.inv_half_aH;
  double eta = eta_y;
  if (flags & LR_GIVEN_METRIC) { h_prime = ms_in[MS_HP]; eta = ms_in[MS_ETA]; }
  double rsa_delta_g = 0., rsa_theta_g = 0.;
  if (ap.rsa_on) {
    double rsa_delta_ur = 0., rsa_theta_ur = 0.;
    if (P.rsa_method != CLPP_RSA_NULL) {
      rsa_delta_g = 4. * ik2 * (aH * h_prime - k2 * eta);
      rsa_theta_g = -0.5 * h_prime;
    }
    if (P.rsa_method == CLPP_RSA_MD_WITH_REIO) {
      rsa_delta_g += -4. * ik2 * dkappa * (theta_b + 0.5 * h_prime);
      rsa_theta_g += 3. * ik2 * (e.ddkappa * (theta_b + 0.5 * h_prime) +
                                 dkappa * (-aH * theta_b + cb2 * k2 * delta_b - aH * h_prime + k2 * eta));
    }
    if (P.has_ur && P.rsa_method != CLPP_RSA_NULL) {
      rsa_delta_ur = 4. * ik2 * (aH * h_prime - k2 * eta);
      rsa_theta_ur = -0.5 * h_prime;
    }
    delta_rho += rho_g * rsa_delta_g;
    rpt += 4. / 3. * rho_g * rsa_theta_g;
    if (P.has_ur) {
      delta_rho += rho_ur * rsa_delta_ur;
      rpt += 4. / 3. * rho_ur * rsa_theta_ur;
    }
  }
  double eta_prime = (1.5 * a2 * rpt) * ik2;
  if (flags & LR_GIVEN_METRIC) eta_prime = ms_in[MS_EP];
  const double alpha = (h_prime + 6. * eta_prime) * 0.5 * ik2;
  if (!ap.tca_off) {
    const double sg = 16. / 45. * e.tau_c * (theta_g + k2 * alpha);
    rps += 4. / 3. * rho_g * sg;
  }
  double alpha_prime = -2. * aH * alpha + eta - 4.5 * (a2 * ik2) * rps;
  if (flags & LR_GIVEN_METRIC) alpha_prime = ms_in[MS_AP];
  m.h_prime = h_prime; m.eta_prime = eta_prime; m.alpha = alpha; m.alpha_prime = alpha_prime;
  m.rsa_delta_g = rsa_delta_g; m.rsa_theta_g = rsa_theta_g;
  if (!(flags & (LR_HUB | LR_CHAINS))) return;

  const double cotKgen = e.inv_tau * M.ik;
  const double metric_continuity = h_prime * 0.5;
  const double metric_shear = k2 * alpha;
  const double metric_ufa_class = h_prime * 0.5;
  if (ap.rsa_on) { delta_g = rsa_delta_g; theta_g = rsa_theta_g; }

  // ---- hub rows
  if (flags & LR_HUB) {
    double dtheta_b, dtheta_g = 0.;
    if (ap.tca_off) {
      dtheta_b = -aH * theta_b + k2 * delta_p_b_over_rho_b + R * dkappa * (theta_g - theta_b);
      if (full_g) {
        const double P0 = (p0 + p2 + 2. * shear_g) * 0.125;
        dtheta_g = k2 * (delta_g * 0.25 - shear_g) + dkappa * (theta_b - theta_g);
        DY(L.shear_g) = 0.5 * (8. / 15. * (theta_g + metric_shear) - 3. / 5. * k * g3 - dkappa * (2. * shear_g - 4. / 5. * P0));
        DY(L.pol0_g) = -k * p1 - dkappa * (p0 - 4. * P0);
        DY(L.pol0_g + 1) = k * (1. / 3.) * (p0 - 2. * p2) - dkappa * p1;
        DY(L.pol0_g + 2) = k * (1. / 5.) * (2. * p1 - 3. * p3) - dkappa * (p2 - 4. / 5. * P0);
      }
    } else {
      // ---- perturb_tca_slip_and_shear
      const double a_primeprime_over_a = e.Hp * e.a + 2. * aH * aH;
      const double tau_c = e.tau_c;
      const double dtau_c = -e.ddkappa * tau_c * tau_c;
      const double i1pR = e.inv_1pR;
      const double F = tau_c * i1pR;
      double F_prime = 0.;
      if (P.tca_method >= CLPP_TCA_SECOND_ORDER_CLASS) F_prime = dtau_c * i1pR + tau_c * aH * R * i1pR * i1pR;
      const double metric_shear_prime = k2 * alpha_prime;
      const double common = F * (-a_primeprime_over_a * theta_b +
                                 k2 * (-aH * delta_g * 0.5 + cb2 * (-theta_b - metric_continuity) -
                                       4. / 3. * (-theta_g - metric_continuity) * 0.25));
      double slip;
      if (P.tca_method == CLPP_TCA_FIRST_ORDER_MB) slip = 2. * R * i1pR * aH * (theta_b - theta_g) + common;
      else slip = (dtau_c * dkappa - 2. * aH * i1pR) * (theta_b - theta_g) + common;
      double sg = 16. / 45. * tau_c * (theta_g + metric_shear);
      const double theta_prime = (-aH * theta_b + k2 * (cb2 * delta_b + R * 0.25 * delta_g)) * i1pR;
      const double shear_g_prime = 16. / 45. * (tau_c * (theta_prime + metric_shear_prime) + dtau_c * (theta_g + metric_shear));
      if (P.tca_method == CLPP_TCA_COMPROMISE_CLASS) {
        slip = (1. - 2. * aH * F) * slip +
               F * k2 * (2. * aH * sg + shear_g_prime - (1. / 3. - cb2) * (F * theta_prime + 2. * F_prime * theta_b));
        sg = (1. - 11. / 6. * dtau_c) * sg - 11. / 6. * tau_c * 16. / 45. * tau_c * (theta_prime + metric_shear_prime);
      }
      m.tca_shear_g = sg;
      dtheta_b = (-aH * theta_b + k2 * (delta_p_b_over_rho_b + R * (delta_g * 0.25 - sg)) + R * slip) * i1pR;
      dtheta_g = -(dtheta_b + aH * theta_b - k2 * delta_p_b_over_rho_b) * e.inv_R + k2 * (0.25 * delta_g - sg);
    }
    if (has_g) {
      DY(L.delta_g) = -4. / 3. * (theta_g + metric_continuity);
      DY(L.theta_g) = dtheta_g;
    }
    DY(L.delta_b) = -(theta_b + metric_continuity);
    DY(L.theta_b) = dtheta_b;
    DY(L.delta_cdm) = -metric_continuity;
    DY(L.eta) = eta_prime;
    if (has_ur) {
      DY(L.delta_ur) = -4. / 3. * (theta_ur + metric_continuity) +
                       (1. - P.three_ceff2_ur) * aH * (delta_ur + 4. * aH * theta_ur * ik2);
      DY(L.theta_ur) = k2 * (P.three_ceff2_ur * delta_ur * 0.25 - shear_ur) - (1. - P.three_ceff2_ur) * aH * theta_ur;
      double dshear_ur;
      if (!ap.ufa_on) {
        dshear_ur = 0.5 * (8. / 15. * (theta_ur + metric_shear) - 3. / 5. * k * u3 -
                           (1. - P.three_cvis2_ur) * (8. / 15. * (theta_ur + metric_shear)));
      } else {
        if (P.ufa_method == CLPP_UFA_MB) dshear_ur = -3. * e.inv_tau * shear_ur + 2. / 3. * (theta_ur + metric_shear);
        else if (P.ufa_method == CLPP_UFA_HU) dshear_ur = -3. * aH * shear_ur + 2. / 3. * (theta_ur + metric_shear);
        else dshear_ur = -3. * e.inv_tau * shear_ur + 2. / 3. * (theta_ur + metric_ufa_class);
      }
      DY(L.shear_ur) = dshear_ur;
    }
    if (P.has_ncdm) {
      if (ap.ncdmfa_on) {
        for (int s = 0; s < P.N_ncdm; s++) {
          const double* nf = e.nf[s];
          const int idx = L.psi0_ncdm1 + 3 * s;
          const double y0 = Y(idx), y1 = Y(idx + 1), y2 = Y(idx + 2);
          const double w_n = nf[2], ca2 = nf[4];
          DY(idx) = -(1.0 + w_n) * (y1 + metric_continuity) - 3.0 * aH * (ca2 - w_n) * y0;
          DY(idx + 1) = -aH * (1.0 - 3.0 * ca2) * y1 + nf[5] * k2 * y0 - k2 * y2;
          const double msn = (P.ncdmfa_method == CLPP_NCDMFA_CLASS) ? metric_ufa_class : metric_shear;
          DY(idx + 2) = -nf[7] * y2 + nf[6] * (y1 + msn);
        }
      } else {
        for (int j = 0; j < P.nq_tot; j++) {
          const int idx = L.psi0_ncdm1 + 3 * j;
          const double qk = k * LM(P.lo_nw + j);
          const double dlnf0 = LN_LDG(M.C->ncdm_dlnf0 + j);
          const double y0 = Y(idx), y1 = Y(idx + 1), y2 = Y(idx + 2), y3 = Y(L.c_ncdm + j * L.len_ncdm);
          DY(idx) = -qk * y1 + metric_continuity * dlnf0 * (1. / 3.);
          DY(idx + 1) = qk * (1. / 3.0) * (y0 - 2 * y2);
          DY(idx + 2) = qk * (1. / 5.0) * (2 * y1 - 3. * y3) - metric_shear * 2. / 15. * dlnf0;
        }
      }
    }
  }
  // ---- multipole chains (l >= 3)
  if (flags & LR_CHAINS) {
    const double* __restrict__ i2l1 = P.i2l1;
    if (full_g) {
      {
        const int c0 = L.c_g, len = L.len_g;
        double ym = 2. * shear_g, yl = Y(c0);
        for (int p = 0; p < len; p++) {
          const int l = 3 + p;
          if (p < len - 1) {
            const double yp = Y(c0 + p + 1);
            DY(c0 + p) = k * LN_LDG(i2l1 + l) * (l * ym - (l + 1) * yp) - dkappa * yl;
            ym = yl; yl = yp;
          } else {
            DY(c0 + p) = k * (ym - (1. + l) * cotKgen * yl) - dkappa * yl;
          }
        }
      }
      {
        const int c0 = L.c_pol, len = L.len_pol;
        double ym = p2, yl = Y(c0);
        for (int p = 0; p < len; p++) {
          const int l = 3 + p;
          if (p < len - 1) {
            const double yp = Y(c0 + p + 1);
            DY(c0 + p) = k * LN_LDG(i2l1 + l) * (l * ym - (l + 1.) * yp) - dkappa * yl;
            ym = yl; yl = yp;
          } else {
            DY(c0 + p) = k * (ym - (l + 1) * cotKgen * yl) - dkappa * yl;
          }
        }
      }
    }
    if (has_ur && !ap.ufa_on) {
      const int c0 = L.c_ur, len = L.len_ur;
      double ym = 2. * shear_ur, yl = Y(c0);
      for (int p = 0; p < len; p++) {
        const int l = 3 + p;
        if (p < len - 1) {
          const double yp = Y(c0 + p + 1);
          DY(c0 + p) = k * LN_LDG(i2l1 + l) * (l * ym - (l + 1.) * yp);
          ym = yl; yl = yp;
        } else {
          DY(c0 + p) = k * (ym - (1. + l) * cotKgen * yl);
        }
      }
    }
    if (P.has_ncdm && !ap.ncdmfa_on) {
      const int len = L.len_ncdm;
      for (int j = 0; j < P.nq_tot; j++) {
        const int c0 = L.c_ncdm + j * len;
        const double qk = k * LM(P.lo_nw + j);
        double ym = Y(L.psi0_ncdm1 + 3 * j + 2), yl = Y(c0);
        for (int p = 0; p < len; p++) {
          const int l = 3 + p;
          if (p < len - 1) {
            const double yp = Y(c0 + p + 1);
            DY(c0 + p) = qk * LN_LDG(i2l1 + l) * (l * ym - (l + 1.) * yp);
            ym = yl; yl = yp;
          } else {
            DY(c0 + p) = qk * ym - (1. + l) * k * cotKgen * yl;
          }
        }
      }
    }
  }
#undef Y
#undef DY
}

// -------------------------------------------------------------------------------------------------
// perturb_sources_member: source functions at sample index_tau from (y, dy)
LN_NOINLINE void ln_write_sources(const PtParams& P, Lane& M, double* __restrict__ mem, double tau, int sy, int sdy, int index_tau) {
  ln_env(P, M, mem, tau, 1);
  ln_rhs(P, M, mem, sy, -1, LR_MATTER, nullptr);
  const LnLayout& L = M.L;
  const Approx& ap = M.ap;
  const LnEnv& e = M.e;
  const LnMetric& m = M.m;
  const double k = M.k;
  const double z = M.a_today / e.a - 1.;
  const double aH = e.a * e.H;
  const double aH_prime = e.Hp * e.a + (e.H * e.a) * (e.H * e.a);
  double delta_g, Pi;
  if (ap.rsa_on) { delta_g = m.rsa_delta_g; Pi = 0.; }
  else {
    delta_g = LVEC(sy, L.delta_g);
    if (!ap.tca_off) Pi = 5. * M.tca_shear_last / 8.;
    else Pi = (LVEC(sy, L.pol0_g) + LVEC(sy, L.pol0_g + 2) + 2. * LVEC(sy, L.shear_g)) / 8.;
  }
  const size_t stride_tp = (size_t)M.C->k_size * M.C->tau_size;
  double* out = M.C->sources + (size_t)M.ik_index * M.C->tau_size + index_tau;
  const double eta = LVEC(sy, L.eta);
  if (P.tp_t0 >= 0) {
    int switch_isw = 1;
    if ((P.switch_eisw == 0) && (z >= P.eisw_lisw_split_z)) switch_isw = 0;
    if ((P.switch_lisw == 0) && (z < P.eisw_lisw_split_z)) switch_isw = 0;
    const double theta_b = LVEC(sy, L.theta_b), dtheta_b = LVEC(sdy, L.theta_b);
    out[P.tp_t0 * stride_tp] =
        P.switch_sw * e.g * (delta_g / 4. + m.alpha_prime) +
        switch_isw * (e.g * (eta - m.alpha_prime - 2 * aH * m.alpha) +
                      e.exp_m_kappa * 2. * (m.eta_prime - aH_prime * m.alpha - aH * m.alpha_prime)) +
        P.switch_dop * (e.g * (dtheta_b / k / k + m.alpha_prime) + e.dg * (theta_b / k / k + m.alpha));
    out[P.tp_t1 * stride_tp] = switch_isw * e.exp_m_kappa * k * (m.alpha_prime + 2. * aH * m.alpha - eta);
    out[P.tp_t2 * stride_tp] = P.switch_pol * e.g * Pi;
  }
  if (P.tp_p >= 0) out[P.tp_p * stride_tp] = sqrt(6.) * e.g * Pi;
  if (P.tp_phi_plus_psi >= 0) out[P.tp_phi_plus_psi * stride_tp] = eta + m.alpha_prime;
  if (P.tp_delta_m >= 0) out[P.tp_delta_m * stride_tp] = m.delta_m;
  if (P.tp_delta_cb >= 0) out[P.tp_delta_cb * stride_tp] = m.delta_cb;
}

// -------------------------------------------------------------------------------------------------
// Jacobian J = A(tau) at the environment in M.e.
//  chains: closed form (each l >= 3 row couples to l-1, l, l+1 only);
//  hub block (nh x nh, row-major at lo_jhh): column j = hub rows of f(tau, e_j) (the system is linear and homogeneous).
LN_NOINLINE void ln_jacobian(const PtParams& P, Lane& M, double* __restrict__ mem) {
  const LnLayout& L = M.L;
  const LnEnv& e = M.e;
  const int n = L.neq, nh = L.nh;
  const double k = M.k;
  const double cotKgen = e.inv_tau * M.ik;
  const double* __restrict__ i2l1 = P.i2l1;
  // ---- chains
  int c = 0;
  if (L.c_g >= 0) {
    for (int fam = 0; fam < 2; fam++) {
      const int c0 = fam == 0 ? L.c_g : L.c_pol, len = fam == 0 ? L.len_g : L.len_pol;
      for (int p = 0; p < len; p++) {
        const int l = 3 + p;
        const double first = (p == 0 && fam == 0) ? 2. : 1.;  // F_2 = 2 shear_g
        if (p < len - 1) {
          LVEC(LV_JL, c0 + p) = first * k * LN_LDG(i2l1 + l) * l;
          LVEC(LV_JU, c0 + p) = -k * LN_LDG(i2l1 + l) * (l + 1);
          LVEC(LV_JD, c0 + p) = -e.dkappa;
        } else {
          LVEC(LV_JL, c0 + p) = first * k;
          LVEC(LV_JU, c0 + p) = 0.;
          LVEC(LV_JD, c0 + p) = -k * (1. + l) * cotKgen - e.dkappa;
        }
      }
      LCH_JUR(c) = fam == 0 ? -0.3 * k : -0.6 * k;
      c++;
    }
  }
  if (L.c_ur >= 0) {
    const int c0 = L.c_ur, len = L.len_ur;
    for (int p = 0; p < len; p++) {
      const int l = 3 + p;
      const double first = (p == 0) ? 2. : 1.;
      if (p < len - 1) {
        LVEC(LV_JL, c0 + p) = first * k * LN_LDG(i2l1 + l) * l;
        LVEC(LV_JU, c0 + p) = -k * LN_LDG(i2l1 + l) * (l + 1);
        LVEC(LV_JD, c0 + p) = 0.;
      } else {
        LVEC(LV_JL, c0 + p) = first * k;
        LVEC(LV_JU, c0 + p) = 0.;
        LVEC(LV_JD, c0 + p) = -k * (1. + l) * cotKgen;
      }
    }
    LCH_JUR(c) = -0.3 * k;
    c++;
  }
  if (L.c_ncdm >= 0) {
    const int len = L.len_ncdm;
    for (int j = 0; j < P.nq_tot; j++) {
      const int c0 = L.c_ncdm + j * len;
      const double qk = k * LM(P.lo_nw + j);
      for (int p = 0; p < len; p++) {
        const int l = 3 + p;
        if (p < len - 1) {
          LVEC(LV_JL, c0 + p) = qk * LN_LDG(i2l1 + l) * l;
          LVEC(LV_JU, c0 + p) = -qk * LN_LDG(i2l1 + l) * (l + 1);
          LVEC(LV_JD, c0 + p) = 0.;
        } else {
          LVEC(LV_JL, c0 + p) = qk;
          LVEC(LV_JU, c0 + p) = 0.;
          LVEC(LV_JD, c0 + p) = -(1. + l) * k * cotKgen;
        }
      }
      LCH_JUR(c) = -0.6 * qk;
      c++;
    }
  }
  // ---- hub block by probing (hub rows only; the chain parts of the probe vector stay zero)
  for (int i = 0; i < n; i++) LVEC(LV_TMP, i) = 0.;
  for (int j = 0; j < nh; j++) {
    LVEC(LV_TMP, j) = 1.;
    ln_rhs(P, M, mem, LV_TMP, LV_DEL, LR_HUB, nullptr);
    for (int i = 0; i < nh; i++) LM(P.lo_jhh + i * nh + j) = LVEC(LV_DEL, i);
    LVEC(LV_TMP, j) = 0.;
  }
  M.st.jacobians++;
  M.st.fevals += nh;
}

// Factorisation of A = I - c J: chains (backward elimination towards their root), Schur complement on the root
// diagonals, LU with partial pivoting of the hub block (row-major at lo_lu, 1/pivot on the diagonal).
LN_NOINLINE void ln_factor(const PtParams& P, Lane& M, double* __restrict__ mem, double c) {
  const LnLayout& L = M.L;
  const int nh = L.nh, nch = L.nch;
  M.fac_c = c;
  for (int i = 0; i < nh; i++)
    for (int j = 0; j < nh; j++) LM(P.lo_lu + i * nh + j) = (i == j ? 1.0 : 0.0) - c * LM(P.lo_jhh + i * nh + j);
  for (int ch = 0; ch < nch; ch++) {
    int s, len, root;
    ln_chain(L, ch, s, len, root);
    const int last = s + len - 1;
    double ipn = 1.0 / (1.0 - c * LVEC(LV_JD, last));
    LVEC(LV_IP, last) = ipn;
    double lon = -c * LVEC(LV_JL, last);
    for (int i = last - 1; i >= s; i--) {
      const double mui = -c * LVEC(LV_JU, i) * ipn;
      const double p = (1.0 - c * LVEC(LV_JD, i)) - mui * lon;
      ipn = 1.0 / p;
      lon = -c * LVEC(LV_JL, i);
      LVEC(LV_IP, i) = ipn; LVEC(LV_MU, i) = mui;
    }
    const double mur = -c * LCH_JUR(ch) * ipn;
    LCH_MUR(ch) = mur;
    LM(P.lo_lu + root * nh + root) += -mur * lon;
  }
  // LU, partial pivoting
  for (int j = 0; j < nh; j++) {
    double best = fabs(LM(P.lo_lu + j * nh + j));
    int bi = j;
    for (int i = j + 1; i < nh; i++) {
      const double v = fabs(LM(P.lo_lu + i * nh + j));
      if (v > best) { best = v; bi = i; }
    }
    LM(P.lo_piv + j) = (double)bi;
    if (bi != j) {
      for (int cc = 0; cc < nh; cc++) {
        const double t = LM(P.lo_lu + j * nh + cc);
        LM(P.lo_lu + j * nh + cc) = LM(P.lo_lu + bi * nh + cc);
        LM(P.lo_lu + bi * nh + cc) = t;
      }
    }
    double pv = LM(P.lo_lu + j * nh + j);
    if (pv == 0.) pv = 1e-50;  // TINY, as ludcmp does for a singular pivot
    const double pinv = 1.0 / pv;
    LM(P.lo_lu + j * nh + j) = pinv;
    for (int i = j + 1; i < nh; i++) {
      const double f = LM(P.lo_lu + i * nh + j) * pinv;
      LM(P.lo_lu + i * nh + j) = f;
      if (f != 0.)
        for (int cc = j + 1; cc < nh; cc++) LM(P.lo_lu + i * nh + cc) -= f * LM(P.lo_lu + j * nh + cc);
    }
  }
  M.st.factorizations++;
}

// solve A x = b in place (vector slot sb) with the factors of ln_factor
LN_NOINLINE void ln_solve(const PtParams& P, Lane& M, double* __restrict__ mem, int sb) {
  const LnLayout& L = M.L;
  const int nh = L.nh, nch = L.nch;
  const double c = M.fac_c;
#define B(i) LVEC(sb, i)
  for (int ch = 0; ch < nch; ch++) {
    int s, len, root;
    ln_chain(L, ch, s, len, root);
    const int last = s + len - 1;
    double r = B(last);
    for (int i = last - 1; i >= s; i--) {
      r = B(i) - LVEC(LV_MU, i) * r;
      B(i) = r;
    }
    B(root) -= LCH_MUR(ch) * r;
  }
  for (int j = 0; j < nh; j++) {
    const int p = (int)LM(P.lo_piv + j);
    if (p != j) { const double t = B(j); B(j) = B(p); B(p) = t; }
  }
  for (int i = 1; i < nh; i++) {
    double s = B(i);
    for (int j = 0; j < i; j++) s -= LM(P.lo_lu + i * nh + j) * B(j);
    B(i) = s;
  }
  for (int i = nh - 1; i >= 0; i--) {
    double s = B(i);
    for (int j = i + 1; j < nh; j++) s -= LM(P.lo_lu + i * nh + j) * B(j);
    B(i) = s * LM(P.lo_lu + i * nh + i);
  }
  for (int ch = 0; ch < nch; ch++) {
    int s, len, root;
    ln_chain(L, ch, s, len, root);
    const int last = s + len - 1;
    double xp = B(root);
    for (int i = s; i <= last; i++) {
      const double lo = -c * LVEC(LV_JL, i);
      xp = (B(i) - lo * xp) * LVEC(LV_IP, i);
      B(i) = xp;
    }
  }
#undef B
  M.st.solves++;
}

// rescale the backward differences when the step changes by the factor r (kord = current order)
LN_NOINLINE void ln_adjust_stepsize(const PtParams& P, Lane& M, double* __restrict__ mem, double r, int kord) {
  double RU[5][5];
  {
    double Rm[5][5];
#pragma unroll
    for (int kk = 0; kk < 5; kk++) {
      double Rv = 1.;
#pragma unroll
      for (int ii = 0; ii < 5; ii++) {
        Rv *= (ii - (kk + 1) * r) * c_invint[ii + 1];
        Rm[ii][kk] = Rv;
      }
    }
#pragma unroll
    for (int ii = 0; ii < 5; ii++)
#pragma unroll
      for (int jj = 0; jj < 5; jj++) {
        double s = 0.;
#pragma unroll
        for (int kk = 0; kk < 5; kk++) s += Rm[ii][kk] * c_U[kk][jj];
        RU[ii][jj] = s;
      }
  }
  const int n = M.L.neq;
  for (int i = 0; i < n; i++) {
    double row[5];
#pragma unroll
    for (int kk = 0; kk < 5; kk++) row[kk] = (kk < kord) ? LVEC(LV_DIF0 + kk, i) : 0.;
#pragma unroll
    for (int jj = 0; jj < 5; jj++) {
      if (jj < kord) {
        double s = 0.0;
#pragma unroll
        for (int kk = 0; kk < 5; kk++) s += row[kk] * RU[kk][jj];
        LVEC(LV_DIF0 + jj, i) = s;
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// NDF1-5 over [t0, tfinal] for the current layout (evolver_ndf15.cpp:302-651, same step/order control), written as a
// flat state machine: every pass of the loop is ONE step attempt.
LN_NOINLINE bool ln_ndf15(const PtParams& P, Lane& M, double* __restrict__ mem, double t0, double tfinal) {
  const double eps = 1e-16, threshold = 1e-15;
  const int maxit = 4, maxk = 5;
  const double rtol = P.rtol;
  const int n = M.L.neq;
  const double* t_vec = M.C->tau;
  const int tres = M.C->tau_size;
  int next = M.next;
  while (next < tres && LN_LDG(t_vec + next) < t0) next++;
  double tnext = (next < tres) ? LN_LDG(t_vec + next) : 1e300;
  for (int j = 0; j < 7; j++)
    for (int i = 0; i < n; i++) LVEC(LV_DIF0 + j, i) = 0.;
  const double htspan = fabs(tfinal - t0);
  double t = t0, tnew = t0;
  ln_env(P, M, mem, t0, 0);
  ln_rhs(P, M, mem, LV_Y, LV_F, LR_HUB | LR_CHAINS, nullptr);
  M.st.fevals++;
  const double hmax = (tfinal - t0) / 10.0;
  ln_jacobian(P, M, mem);
  bool Jcurrent = true;
  double hmin = 16.0 * eps * fabs(t);
  double rh = 0.0;
  for (int i = 0; i < n; i++) {
    const double wt = fmax(fabs(LVEC(LV_Y, i)), threshold);
    rh = fmax(rh, 1.25 / sqrt(rtol) * fabs(LVEC(LV_F, i) / wt));
  }
  double absh = fmin(hmax, htspan);
  if (absh * rh > 1.0) absh = 1.0 / rh;
  absh = fmax(absh, hmin);
  double h = absh;
  {
    ln_rhs(P, M, mem, LV_F, LV_PSI, LR_HUB | LR_CHAINS, nullptr);  // J*f0 = f(t0, f0): linear, homogeneous
    M.st.fevals++;
    const double tdel = (t + fmin(sqrt(eps) * fmax(fabs(t), fabs(t + h)), absh)) - t;
    ln_env(P, M, mem, t + tdel, 0);
    ln_rhs(P, M, mem, LV_Y, LV_DEL, LR_HUB | LR_CHAINS, nullptr);
    M.st.fevals++;
    rh = 0.0;
    for (int i = 0; i < n; i++) {
      const double wt = fmax(fabs(LVEC(LV_Y, i)), threshold);
      const double s = LVEC(LV_PSI, i) + (LVEC(LV_DEL, i) - LVEC(LV_F, i)) / tdel;
      rh = fmax(rh, 1.25 * sqrt(0.5 * fabs(s / wt) / rtol));
    }
    absh = fmin(hmax, htspan);
    if (absh * rh > 1.0) absh = 1.0 / rh;
    absh = fmax(absh, hmin);
    h = absh;
  }
  int k = 1, klast = k;
  double abshlast = absh;
  for (int i = 0; i < n; i++) LVEC(LV_DIF0, i) = h * LVEC(LV_F, i);
  double hinvGak = h * c_invGa[k - 1];
  int nconhk = 0;
  ln_factor(P, M, mem, hinvGak);
  bool havrate = false, done = false, at_hmin = false, new_step = true, nofailed = true;
  double rate = 0., oldnrm = 0., err = 0.;

  for (;;) {
    if (new_step) {
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
        ln_adjust_stepsize(P, M, mem, absh / abshlast, k);
        hinvGak = h * c_invGa[k - 1];
        nconhk = 0;
        ln_factor(P, M, mem, hinvGak);
        havrate = false;
      }
      nofailed = true;
      new_step = false;
    }
    // ---- one attempt
    tnew = t + h;
    if (done) tnew = tfinal;
    h = tnew - t;
    double minnrm = 0.0;
    {
      const double invGak = c_invGa[k - 1];
      for (int i = 0; i < n; i++) {
        double ps = 0.0;
        const double yi = LVEC(LV_Y, i);
        double pr = yi;
        for (int j = 0; j < k; j++) {
          const double d = LVEC(LV_DIF0 + j, i);
          ps += d * c_G[j] * invGak;
          pr += d;
        }
        LVEC(LV_PSI, i) = ps;
        LVEC(LV_PRED, i) = pr;
        LVEC(LV_YNEW, i) = pr;
        LVEC(LV_DIFKP1, i) = 0.0;
        const double iw = 1.0 / fmax(fmax(fabs(pr), fabs(yi)), threshold);
        LVEC(LV_INVWT, i) = iw;
        minnrm = fmax(minnrm, 100 * eps * fabs(pr * iw));
      }
    }
    ln_env(P, M, mem, tnew, 0);
    bool tooslow = false, gotynew = false;
    for (int iter = 1; iter <= maxit; iter++) {
      ln_rhs(P, M, mem, LV_YNEW, LV_F, LR_HUB | LR_CHAINS, nullptr);
      if (!M.ap.tca_off) M.tca_shear_last = M.m.tca_shear_g;
      M.st.fevals++;
      for (int i = 0; i < n; i++) LVEC(LV_DEL, i) = hinvGak * LVEC(LV_F, i) - (LVEC(LV_PSI, i) + LVEC(LV_DIFKP1, i));
      ln_solve(P, M, mem, LV_DEL);
      double newnrm = 0.0;
      for (int i = 0; i < n; i++) {
        const double d = LVEC(LV_DEL, i);
        newnrm = fmax(newnrm, fabs(d * LVEC(LV_INVWT, i)));
        const double dk = LVEC(LV_DIFKP1, i) + d;
        LVEC(LV_DIFKP1, i) = dk;
        LVEC(LV_YNEW, i) = LVEC(LV_PRED, i) + dk;
      }
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
        else {
          double rp = rate;  // rate^(maxit-iter)
          for (int q = 1; q < maxit - iter; q++) rp *= rate;
          if (0.5 * rtol < errit * rp) { tooslow = true; break; }
        }
      }
      oldnrm = newnrm;
    }
    if (!gotynew) {  // Newton iteration too slow (tooslow is implied)
      (void)tooslow;
      M.st.failed++;
      if (!Jcurrent) {
        ln_env(P, M, mem, t, 0);
        M.st.fevals++;  // the reference re-evaluates f(t, y) for numjac; the probes below need the environment only
        ln_jacobian(P, M, mem);
        Jcurrent = true;
      } else if (absh <= hmin) {
        M.status = 2;  // step size too small
        return false;
      } else {
        abshlast = absh;
        absh = fmax(0.3 * absh, hmin);
        h = absh;
        done = false;
        ln_adjust_stepsize(P, M, mem, absh / abshlast, k);
        hinvGak = h * c_invGa[k - 1];
        nconhk = 0;
      }
      ln_factor(P, M, mem, hinvGak);
      havrate = false;
      continue;
    }
    // ---- error estimate
    err = 0.0;
    for (int i = 0; i < n; i++) err = fmax(err, fabs(LVEC(LV_DIFKP1, i) * LVEC(LV_INVWT, i)));
    err *= c_erconst[k - 1];
    if (err > rtol) {
      M.st.failed++;
      if (absh <= hmin) {
        M.status = 2;
        return false;
      }
      abshlast = absh;
      if (nofailed) {
        nofailed = false;
        double hopt = absh * fmax(0.1, 0.833 * ln_root_n(rtol / err, k + 1.0));
        if (k > 1) {
          double errkm1 = 0.0;
          for (int i = 0; i < n; i++)
            errkm1 = fmax(errkm1, fabs((LVEC(LV_DIF0 + (k - 1), i) + LVEC(LV_DIFKP1, i)) * LVEC(LV_INVWT, i)));
          errkm1 *= c_erconst[k - 2];
          const double hkm1 = absh * fmax(0.1, 0.769 * ln_root_n(rtol / errkm1, (double)k));
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
      ln_adjust_stepsize(P, M, mem, absh / abshlast, k);
      hinvGak = h * c_invGa[k - 1];
      nconhk = 0;
      ln_factor(P, M, mem, hinvGak);
      havrate = false;
      continue;
    }
    // ---- step accepted
    M.st.steps++;
    {
      double e_km1 = 0., e_kp1 = 0.;
      const bool want_order = !done && (nconhk + 1 >= k + 2 || nconhk + 1 >= maxk + 2);
      for (int i = 0; i < n; i++) {
        const double dk = LVEC(LV_DIFKP1, i);
        const double dkp1 = dk - LVEC(LV_DIF0 + k, i);
        LVEC(LV_DIF0 + k + 1, i) = dkp1;
        double acc = dk;
        LVEC(LV_DIF0 + k, i) = acc;
        for (int j = k - 1; j >= 0; j--) {
          acc += LVEC(LV_DIF0 + j, i);
          LVEC(LV_DIF0 + j, i) = acc;
          if (j == k - 1 && want_order) e_km1 = fmax(e_km1, fabs(acc * LVEC(LV_INVWT, i)));
        }
        if (want_order) e_kp1 = fmax(e_kp1, fabs(dkp1 * LVEC(LV_INVWT, i)));
      }
      // ---- output at the sample times passed by this step
      while ((next < tres) && ((tnew - tnext) >= 0.0)) {
        if (tnew == tnext) {
          ln_write_sources(P, M, mem, tnext, LV_YNEW, LV_F, next);
        } else {
          const double s = (tnext - tnew) / h;
          double c1[5], c2[5];
          {
            double prod = 1.0, sumfrac = 0., fact = 1.0;
            for (int j = 0; j < k; j++) {
              prod *= (s + j);
              fact *= (j + 1);
              sumfrac += 1.0 / (s + j);
              c1[j] = prod / fact;
              c2[j] = prod * sumfrac / (h * fact);
            }
          }
          for (int i = 0; i < n; i++) {
            double a1 = 0, a2 = 0;
            for (int j = 0; j < k; j++) {
              const double d = LVEC(LV_DIF0 + j, i);
              a1 += c1[j] * d;
              a2 += c2[j] * d;
            }
            LVEC(LV_TMP, i) = LVEC(LV_YNEW, i) + a1;
            LVEC(LV_YPI, i) = a2;
          }
          ln_write_sources(P, M, mem, tnext, LV_TMP, LV_YPI, next);
        }
        next++;
        tnext = (next < tres) ? LN_LDG(t_vec + next) : 1e300;
      }
      if (done) break;
      klast = k;
      abshlast = absh;
      nconhk = nconhk + 1 < maxk + 2 ? nconhk + 1 : maxk + 2;
      if (nconhk >= k + 2) {
        double temp = 0.;
        if (err > 0.) temp = 1.2 * ln_root_n(err / rtol, k + 1.0);
        double hopt = (temp > 0.1) ? absh / temp : 10 * absh;
        int kopt = k;
        if (k > 1) {
          e_km1 *= c_erconst[k - 2];
          temp = 0.;
          if (e_km1 > 0.) temp = 1.3 * ln_root_n(e_km1 / rtol, (double)k);
          const double hkm1 = (temp > 0.1) ? absh / temp : 10 * absh;
          if (hkm1 > hopt) { hopt = hkm1; kopt = k - 1; }
        }
        if (k < maxk) {
          e_kp1 *= c_erconst[k];
          temp = 0.;
          if (e_kp1 > 0.) temp = 1.4 * ln_root_n(e_kp1 / rtol, k + 2.0);
          const double hkp1 = (temp > 0.1) ? absh / temp : 10 * absh;
          if (hkp1 > hopt) { hopt = hkp1; kopt = k + 1; }
        }
        if (hopt > absh) {
          absh = hopt;
          if (k != kopt) k = kopt;
        }
      }
    }
    t = tnew;
    for (int i = 0; i < n; i++) LVEC(LV_Y, i) = LVEC(LV_YNEW, i);
    Jcurrent = false;
    new_step = true;
  }
  // final state, and a last RHS call so that the environment and the TCA/RSA by-products are current at the end of
  // the interval (evolver_ndf15.cpp:653-662)
  for (int i = 0; i < n; i++) LVEC(LV_Y, i) = LVEC(LV_YNEW, i);
  ln_env(P, M, mem, tnew, 0);
  ln_rhs(P, M, mem, LV_Y, LV_F, LR_HUB | LR_CHAINS, nullptr);
  if (!M.ap.tca_off) M.tca_shear_last = M.m.tca_shear_g;
  M.st.fevals++;
  M.next = next;
  return true;
}

// -------------------------------------------------------------------------------------------------
// perturb_initial_conditions: adiabatic mode, synchronous gauge, flat space
LN_NOINLINE void ln_initial_conditions(const PtParams& P, Lane& M, double* __restrict__ mem, double tau) {
  M.bg_cur = M.bt_size / 2; M.th_cur = M.tt_size / 2;
  ln_env(P, M, mem, tau, 0);
  const LnEnv& e = M.e;
  const LnLayout& L = M.L;
  const double k = M.k, a = e.a;
  double rho_r = e.rho_g, rho_m = e.rho_b + e.rho_cdm, rho_nu = 0.;
  if (P.has_ur) { rho_r += e.rho_ur; rho_nu += e.rho_ur; }
  for (int s = 0; s < P.N_ncdm; s++) { rho_r += e.rho_n[s]; rho_nu += e.rho_n[s]; }
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
  for (int i = 0; i < L.neq; i++) LVEC(LV_Y, i) = 0.;
  LVEC(LV_Y, L.delta_g) = delta_g;
  LVEC(LV_Y, L.theta_g) = theta_g;
  LVEC(LV_Y, L.delta_b) = 3. / 4. * delta_g;
  LVEC(LV_Y, L.theta_b) = theta_g;
  LVEC(LV_Y, L.delta_cdm) = 3. / 4. * delta_g;
  LVEC(LV_Y, L.eta) = eta;
  if (P.has_ur) {
    LVEC(LV_Y, L.delta_ur) = delta_ur;
    LVEC(LV_Y, L.theta_ur) = theta_ur;
    LVEC(LV_Y, L.shear_ur) = shear_ur;
    LVEC(LV_Y, L.c_ur) = l3_ur;
  }
  if (P.has_ncdm) {
    for (int s = 0; s < P.N_ncdm; s++) {
      const double Ms = M.C->ncdm_M[s];
      for (int j = P.ncdm_q_off[s]; j < P.ncdm_q_off[s] + P.ncdm_q_size[s]; j++) {
        const int idx = L.psi0_ncdm1 + 3 * j;
        const double q = M.C->ncdm_q[j];
        const double dlnf0 = M.C->ncdm_dlnf0[j];
        const double epsq = sqrt(q * q + a * a * Ms * Ms);
        LVEC(LV_Y, idx + 0) = -0.25 * delta_ur * dlnf0;
        LVEC(LV_Y, idx + 1) = -epsq / 3. / q / k * theta_ur * dlnf0;
        LVEC(LV_Y, idx + 2) = -0.5 * shear_ur * dlnf0;
        LVEC(LV_Y, L.c_ncdm + j * L.len_ncdm) = -0.25 * l3_ur * dlnf0;
      }
    }
  }
}

// perturb_vector_init (switching part): move the state from the old layout (in LV_Y) to the new one
LN_NOINLINE void ln_remap_state(const PtParams& P, Lane& M, double* __restrict__ mem) {
  const LnLayout& Lo = M.Lprev;
  const LnLayout& Ln = M.L;
  const Approx& apo = M.apprev;
  const Approx& apn = M.ap;
  const double k = M.k;
#define YO(i) LVEC(LV_Y, i)
#define YN(i) LVEC(LV_YNEW, i)
  for (int i = 0; i < Ln.neq; i++) YN(i) = 0.;
  YN(Ln.delta_b) = YO(Lo.delta_b);
  YN(Ln.theta_b) = YO(Lo.theta_b);
  YN(Ln.delta_cdm) = YO(Lo.delta_cdm);
  YN(Ln.eta) = YO(Lo.eta);
  if (Ln.delta_g >= 0 && Lo.delta_g >= 0) { YN(Ln.delta_g) = YO(Lo.delta_g); YN(Ln.theta_g) = YO(Lo.theta_g); }
  if (Ln.delta_ur >= 0 && Lo.delta_ur >= 0) {
    YN(Ln.delta_ur) = YO(Lo.delta_ur); YN(Ln.theta_ur) = YO(Lo.theta_ur); YN(Ln.shear_ur) = YO(Lo.shear_ur);
  }
  if (Ln.shear_g >= 0 && Lo.shear_g < 0) {
    // tight coupling switched off: seed the hierarchy from the TCA expressions (:3909-3915); tca_shear_g and kappa'
    // are those of the last RHS call of the previous interval
    const double sg = M.m.tca_shear_g, dk = M.e.dkappa;
    YN(Ln.shear_g) = sg;
    YN(Ln.c_g) = 6. / 7. * k / dk * sg;
    YN(Ln.pol0_g) = 2.5 * sg;
    YN(Ln.pol0_g + 1) = k / dk * (5. - 2.) / 6. * sg;
    YN(Ln.pol0_g + 2) = 0.5 * sg;
    YN(Ln.c_pol) = k / dk * 3. / 14. * sg;
  }
  if (Ln.shear_g >= 0 && Lo.shear_g >= 0) {
    YN(Ln.shear_g) = YO(Lo.shear_g);
    for (int l = 0; l < 3; l++) YN(Ln.pol0_g + l) = YO(Lo.pol0_g + l);
    for (int p = 0; p < Ln.len_g; p++) YN(Ln.c_g + p) = YO(Lo.c_g + p);
    for (int p = 0; p < Ln.len_pol; p++) YN(Ln.c_pol + p) = YO(Lo.c_pol + p);
  }
  if (Ln.c_ur >= 0 && Lo.c_ur >= 0)
    for (int p = 0; p < Ln.len_ur; p++) YN(Ln.c_ur + p) = YO(Lo.c_ur + p);
  if (P.has_ncdm) {
    if (apn.ncdmfa_on == apo.ncdmfa_on) {
      for (int i = 0; i < 3 * Ln.nbin; i++) YN(Ln.psi0_ncdm1 + i) = YO(Lo.psi0_ncdm1 + i);
      if (Ln.c_ncdm >= 0)
        for (int i = 0; i < Ln.len_ncdm * P.nq_tot; i++) YN(Ln.c_ncdm + i) = YO(Lo.c_ncdm + i);
    } else {
      // ncdm fluid approximation switched on: integrate the momentum hierarchy (:4478-4518)
      const double a = M.e.a;
      const double a_rel = M.C->a_today / a, a_rel4 = (a_rel * a_rel) * (a_rel * a_rel);
      for (int s = 0; s < P.N_ncdm; s++) {
        const double rho_n = M.e.rho_n[s], p_n = M.e.p_n[s];
        const double factor = M.C->ncdm_factor[s] * a_rel4;
        const double Ms = M.C->ncdm_M[s];
        double d = 0., th = 0., sh = 0.;
        for (int j = P.ncdm_q_off[s]; j < P.ncdm_q_off[s] + P.ncdm_q_size[s]; j++) {
          const int idx = Lo.psi0_ncdm1 + 3 * j;
          const double q = M.C->ncdm_q[j], w0 = M.C->ncdm_w[j];
          const double q2 = q * q, epsq = sqrt(q2 + a * a * Ms * Ms);
          d += w0 * q2 * epsq * YO(idx);
          th += w0 * q2 * q * YO(idx + 1);
          sh += w0 * q2 * q2 / epsq * YO(idx + 2);
        }
        YN(Ln.psi0_ncdm1 + 3 * s) = d * factor / rho_n;
        YN(Ln.psi0_ncdm1 + 3 * s + 1) = th * k * factor / (rho_n + p_n);
        YN(Ln.psi0_ncdm1 + 3 * s + 2) = sh * 2. / 3. * factor / (rho_n + p_n);
      }
    }
  }
  for (int i = 0; i < Ln.neq; i++) YO(i) = YN(i);
#undef YO
#undef YN
}

// -------------------------------------------------------------------------------------------------
// one mode from its initial time to today (perturb_solve)
LN_FN void ln_mode(const PtParams& P, double* __restrict__ mem, const PtCosmo* C, int ik) {
  Lane M;
  M.C = C;
  M.ik_index = ik;
  M.k = C->k[ik];
  M.k2 = M.k * M.k;
  M.ik = 1.0 / M.k;
  M.ik2 = 1.0 / M.k2;
  M.bg_tau = C->bg_tau; M.bg_y = C->bg_y; M.bg_dd = C->bg_dd;
  M.th_z = C->th_z; M.th_y = C->th_y; M.th_dd = C->th_dd;
  M.bt_size = C->bt_size; M.tt_size = C->tt_size;
  M.z_last = C->th_z[C->tt_size - 1]; M.th_lin = C->th_linear_below_z; M.a_today = C->a_today;
  M.bg_cur = 0; M.th_cur = C->tt_size - 2;
  M.need_nw = 0; M.status = 0; M.next = 0; M.tca_shear_last = 0.; M.fac_c = 0.;
  M.st.steps = M.st.failed = M.st.fevals = M.st.jacobians = M.st.factorizations = M.st.solves = 0;
  M.m.h_prime = M.m.eta_prime = M.m.alpha = M.m.alpha_prime = M.m.rsa_delta_g = M.m.rsa_theta_g = 0.;
  M.m.delta_m = M.m.delta_cb = M.m.tca_shear_g = 0.;
  M.ap.tca_off = M.ap.rsa_on = M.ap.ufa_on = M.ap.ncdmfa_on = 0;
  ln_make_layout(P, M.ap, M.L);
  const double tau_first = C->tau[0];
  const int tau_size = C->tau_size;
  double limit[PT_MAX_INTERVALS + 1];
  Approx sched[PT_MAX_INTERVALS];

  // ---- start time: bisection on (tau_c/tau_h, tau_h/tau_k, ncdm still relativistic)  (:2592-2635)
  double tau_lower = C->bg_tau[0], tau_upper = tau_first;
  int status = 0;
  {
    ln_env(P, M, mem, tau_lower, 0);
    if (M.e.a * M.e.H / M.e.dkappa > P.start_small_k_at_tau_c_over_tau_h) status = 3;
    if (M.k / M.e.a / M.e.H > P.start_large_k_at_tau_h_over_tau_k) status = 4;
    for (int s = 0; s < P.N_ncdm; s++)
      if (fabs(M.e.p_n[s] / M.e.rho_n[s] - 1. / 3.) > P.tol_ncdm_initial_w) status = 5;
  }
  double tau_mid = 0.5 * (tau_lower + tau_upper);
  if (status == 0) {
    while ((tau_upper - tau_lower) / tau_lower > P.tol_tau_approx) {
      M.bg_cur = M.bt_size / 2; M.th_cur = M.tt_size / 2;
      ln_env(P, M, mem, tau_mid, 0);
      bool early = true;
      for (int s = 0; s < P.N_ncdm; s++)
        if (fabs(M.e.p_n[s] / M.e.rho_n[s] - 1. / 3.) > P.tol_ncdm_initial_w) early = false;
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
  const double tau_end = C->tau[tau_size - 1];

  // ---- schedule of approximation switches (:2940-3231)
  int n_int = 1;
  if (status == 0) {
    const Approx a_ini = ln_approximations_at(P, M, mem, tau_ini);
    const Approx a_end = ln_approximations_at(P, M, mem, tau_end);
    double sw[4];
    int nsw = 0;
    for (int w = 0; w < 4; w++) {
      const int f0 = ln_approx_flag(a_ini, w), f1 = ln_approx_flag(a_end, w);
      if (f1 < f0) { status = 6; break; }
      if (f1 > f0) {
        double lo = tau_ini, hi = tau_end, mid = 0.5 * (lo + hi);
        while (hi - lo > P.tol_tau_approx) {
          const Approx am = ln_approximations_at(P, M, mem, mid);
          if (ln_approx_flag(am, w) > f0) hi = mid; else lo = mid;
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
    for (int i = 1; i < n_int && status == 0; i++) {
      const Approx ai = ln_approximations_at(P, M, mem, 0.5 * (limit[i] + limit[i + 1]));
      const Approx ap = sched[i - 1];
      sched[i] = ai;
      int nchange = 0;
      for (int w = 0; w < 4; w++) {
        if (ln_approx_flag(ai, w) < ln_approx_flag(ap, w)) status = 6;
        if (ln_approx_flag(ai, w) != ln_approx_flag(ap, w)) nchange++;
      }
      if (nchange != 1) status = 7;
    }
    if (a_ini.tca_off || a_ini.rsa_on || a_ini.ufa_on || a_ini.ncdmfa_on) status = 8;
  }
  M.status = status;
  clpp_kstat* ks = C->kstat + ik;

  // ---- integrate interval by interval
  for (int iv = 0; iv < n_int && status == 0; iv++) {
    const Approx apn = sched[iv];
    M.Lprev = M.L;
    M.apprev = M.ap;
    M.ap = apn;
    ln_make_layout(P, apn, M.L);
    if (iv == 0) ln_initial_conditions(P, M, mem, limit[0]);
    else ln_remap_state(P, M, mem);
    M.need_nw = P.has_ncdm && !apn.ncdmfa_on;
    const long long c0 = LN_CLOCK();
    const int s0 = M.st.steps;
    const bool ok = ln_ndf15(P, M, mem, limit[iv], limit[iv + 1]);
    ks->iv_neq[iv] = M.L.neq;
    ks->iv_steps[iv] = M.st.steps - s0;
    ks->iv_cycles[iv] = LN_CLOCK() - c0;
    status = M.status;
    if (!ok) break;
  }
  // zero-fill the samples that were not reached (failure only) and publish the counters
  {
    const size_t stride_tp = (size_t)C->k_size * tau_size;
    double* out = C->sources + (size_t)ik * tau_size;
    const int tps[7] = {P.tp_t0, P.tp_t1, P.tp_t2, P.tp_p, P.tp_delta_m, P.tp_delta_cb, P.tp_phi_plus_psi};
    for (int it = M.next; it < tau_size; it++)
      for (int j = 0; j < 7; j++)
        if (tps[j] >= 0) out[tps[j] * stride_tp + it] = 0.;
    ks->steps = M.st.steps; ks->failed = M.st.failed; ks->fevals = M.st.fevals; ks->jacobians = M.st.jacobians;
    ks->factorizations = M.st.factorizations; ks->solves = M.st.solves;
    ks->intervals = n_int; ks->status = status; ks->tau_ini = tau_ini;
  }
}

#endif
