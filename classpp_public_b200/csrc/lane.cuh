// Stage 1, lane kernels: ONE THREAD integrates one (cosmology, k) mode from its initial time to today.
//
// A warp therefore advances 32 modes at once -- the same k of 32 neighbouring cosmologies of a sweep, or 32 neighbouring
// k of one cosmology -- and every instruction does useful work for 32 modes (the warp-per-mode kernels of perturb.cu
// evaluate all scalar "hub" mathematics once per warp).  The per-mode state (state vector, NDF backward differences,
// Newton-matrix factors; 10..100 KB) lives in a per-thread slab of LOCAL memory (a stack array of the kernel): the
// hardware interleaves local memory by thread, so element e of the 32 lanes of a warp is 256 contiguous bytes and
// every access of a warp is coalesced, and -- unlike global memory, whose stores write through and invalidate the
// line -- local memory is cached WRITE-BACK in L1 (a global-memory slab made every load after a store an L2 round
// trip: 120 k cycles per step for a 7-equation system).  The working set streams through L1/L2.  The time stepping is a flat state machine -- one step ATTEMPT per loop
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
#define LN_CTA 32  // one warp per CTA: the long chains of a launch spread over as many SMs (and L1 caches) as possible
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
#define LN_STRIDE 1
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
#define LVP(slot) (mem + (size_t)(P.lo_vec + (slot) * P.np))
// per-chain scalars: J[root, first], its eliminated multiplier
#define LCH_JUR(c) LM(P.lo_ch + (c))
#define LCH_MUR(c) LM(P.lo_ch + PT_MAX_CHAINS + (c))
// The vector loops below are written as small functions over __restrict__ pointers, one per vector: with a single
// base pointer the compiler must assume that every store may alias every later load and serialises the loop on the
// load latency; a lone warp (single-cosmology runs, the long high-k chains of a batch) has nothing else to hide it.
#define LN_UNROLL _Pragma("unroll 4")
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
//  LR_GIVEN_METRIC: the metric scalars (h', eta', alpha', eta) are imposed through ms_in[] instead of being computed
//  from y: the hub block of the Jacobian is D + U V with V = d(metric)/dy, U = d(rows)/d(metric) and D = d(rows)/dy at
//  fixed metric, which is block diagonal (ln_jacobian_s).
LN_NOINLINE void ln_rhs(const PtParams& P, Lane& M, const double* __restrict__ y, double* __restrict__ dy,
                        const double* __restrict__ nw, int flags, const double* ms_in = nullptr) {
  const LnLayout& L = M.L;
  const Approx ap = M.ap;
  const LnEnv& e = M.e;
  const double k = M.k, k2 = M.k2, ik2 = M.ik2;
  const double a2 = e.a * e.a, aH = e.H * e.a, R = e.R;
  const double rho_g = e.rho_g, rho_b = e.rho_b, rho_cdm = e.rho_cdm, rho_ur = e.rho_ur;
  const double dkappa = e.dkappa, cb2 = e.cb2;
  LnMetric& m = M.m;
  const bool has_g = !ap.rsa_on;
  const bool has_ur = P.has_ur && !ap.rsa_on;
  const bool full_g = has_g && ap.tca_off;
  const int nqt = P.nq_tot;

  double delta_g = 0., theta_g = 0., shear_g = 0.;
  if (has_g) { delta_g = y[L.delta_g]; theta_g = y[L.theta_g]; }
  double g3 = 0., p0 = 0., p1 = 0., p2 = 0., p3 = 0.;
  if (full_g) {
    shear_g = y[L.shear_g]; g3 = y[L.c_g];
    p0 = y[L.pol0_g]; p1 = y[L.pol0_g + 1]; p2 = y[L.pol0_g + 2]; p3 = y[L.c_pol];
  }
  double delta_ur = 0., theta_ur = 0., shear_ur = 0., u3 = 0.;
  if (has_ur) {
    delta_ur = y[L.delta_ur]; theta_ur = y[L.theta_ur]; shear_ur = y[L.shear_ur];
    if (!ap.ufa_on) u3 = y[L.c_ur];
  }
  const double delta_b = y[L.delta_b], theta_b = y[L.theta_b], delta_cdm = y[L.delta_cdm];
  const bool given = (flags & LR_GIVEN_METRIC) != 0;
  const double eta = given ? ms_in[MS_ETA] : y[L.eta];
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
        const double y0 = y[idx], y1 = y[idx + 1], y2 = y[idx + 2];
        delta_rho += nf[0] * y0;
        rpt += nf[1] * y1;
        rps += nf[1] * y2;
        delta_rho_m += nf[0] * y0; rho_m += nf[0];
        rpt_m += nf[1] * y1; rpm += nf[1];
      }
    } else {
      for (int s = 0; s < P.N_ncdm; s++) {
        const double rho_n = e.rho_n[s], p_n = e.p_n[s];
        const double factor = M.C->ncdm_factor[s] * e.fac_ncdm;
        double s_rho = 0., s_theta = 0., s_shear = 0.;
        const int nq = P.ncdm_q_size[s], q0 = P.ncdm_q_off[s];
        LN_UNROLL
        for (int iq = 0; iq < nq; iq++) {
          const int idx = L.psi0_ncdm1 + 3 * (q0 + iq);
          const double y0 = y[idx], y1 = y[idx + 1], y2 = y[idx + 2];
          s_rho += nw[nqt + q0 + iq] * y0;
          s_theta += nw[2 * nqt + q0 + iq] * y1;
          s_shear += nw[3 * nqt + q0 + iq] * y2;
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
  const double h_prime = given ? ms_in[MS_HP] : (k2 * eta + 1.5 * a2 * delta_rho) * e.inv_half_aH;
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
  const double eta_prime = given ? ms_in[MS_EP] : (1.5 * a2 * rpt) * ik2;
  const double alpha = (h_prime + 6. * eta_prime) * 0.5 * ik2;
  if (!ap.tca_off) {
    const double sg = 16. / 45. * e.tau_c * (theta_g + k2 * alpha);
    rps += 4. / 3. * rho_g * sg;
  }
  const double alpha_prime = given ? ms_in[MS_AP] : -2. * aH * alpha + eta - 4.5 * (a2 * ik2) * rps;
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
        dy[L.shear_g] = 0.5 * (8. / 15. * (theta_g + metric_shear) - 3. / 5. * k * g3 - dkappa * (2. * shear_g - 4. / 5. * P0));
        dy[L.pol0_g] = -k * p1 - dkappa * (p0 - 4. * P0);
        dy[L.pol0_g + 1] = k * (1. / 3.) * (p0 - 2. * p2) - dkappa * p1;
        dy[L.pol0_g + 2] = k * (1. / 5.) * (2. * p1 - 3. * p3) - dkappa * (p2 - 4. / 5. * P0);
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
      dy[L.delta_g] = -4. / 3. * (theta_g + metric_continuity);
      dy[L.theta_g] = dtheta_g;
    }
    dy[L.delta_b] = -(theta_b + metric_continuity);
    dy[L.theta_b] = dtheta_b;
    dy[L.delta_cdm] = -metric_continuity;
    dy[L.eta] = eta_prime;
    if (has_ur) {
      dy[L.delta_ur] = -4. / 3. * (theta_ur + metric_continuity) +
                       (1. - P.three_ceff2_ur) * aH * (delta_ur + 4. * aH * theta_ur * ik2);
      dy[L.theta_ur] = k2 * (P.three_ceff2_ur * delta_ur * 0.25 - shear_ur) - (1. - P.three_ceff2_ur) * aH * theta_ur;
      double dshear_ur;
      if (!ap.ufa_on) {
        dshear_ur = 0.5 * (8. / 15. * (theta_ur + metric_shear) - 3. / 5. * k * u3 -
                           (1. - P.three_cvis2_ur) * (8. / 15. * (theta_ur + metric_shear)));
      } else {
        if (P.ufa_method == CLPP_UFA_MB) dshear_ur = -3. * e.inv_tau * shear_ur + 2. / 3. * (theta_ur + metric_shear);
        else if (P.ufa_method == CLPP_UFA_HU) dshear_ur = -3. * aH * shear_ur + 2. / 3. * (theta_ur + metric_shear);
        else dshear_ur = -3. * e.inv_tau * shear_ur + 2. / 3. * (theta_ur + metric_ufa_class);
      }
      dy[L.shear_ur] = dshear_ur;
    }
    if (P.has_ncdm) {
      if (ap.ncdmfa_on) {
        for (int s = 0; s < P.N_ncdm; s++) {
          const double* nf = e.nf[s];
          const int idx = L.psi0_ncdm1 + 3 * s;
          const double y0 = y[idx], y1 = y[idx + 1], y2 = y[idx + 2];
          const double w_n = nf[2], ca2 = nf[4];
          dy[idx] = -(1.0 + w_n) * (y1 + metric_continuity) - 3.0 * aH * (ca2 - w_n) * y0;
          dy[idx + 1] = -aH * (1.0 - 3.0 * ca2) * y1 + nf[5] * k2 * y0 - k2 * y2;
          const double msn = (P.ncdmfa_method == CLPP_NCDMFA_CLASS) ? metric_ufa_class : metric_shear;
          dy[idx + 2] = -nf[7] * y2 + nf[6] * (y1 + msn);
        }
      } else {
        const double* __restrict__ dlnf0_tab = M.C->ncdm_dlnf0;
        LN_UNROLL
        for (int j = 0; j < nqt; j++) {
          const int idx = L.psi0_ncdm1 + 3 * j;
          const double qk = k * nw[j];
          const double dlnf0 = LN_LDG(dlnf0_tab + j);
          const double y0 = y[idx], y1 = y[idx + 1], y2 = y[idx + 2], y3 = y[L.c_ncdm + j * L.len_ncdm];
          dy[idx] = -qk * y1 + metric_continuity * dlnf0 * (1. / 3.);
          dy[idx + 1] = qk * (1. / 3.0) * (y0 - 2 * y2);
          dy[idx + 2] = qk * (1. / 5.0) * (2 * y1 - 3. * y3) - metric_shear * 2. / 15. * dlnf0;
        }
      }
    }
  }
  // ---- multipole chains (l >= 3): row l couples to l-1 (the hub root for l = 3), l, l+1
  if (flags & LR_CHAINS) {
    const double* __restrict__ i2l1 = P.i2l1;
    for (int fam = 0; fam < 3; fam++) {
      int c0, len;
      double ym, damp;
      if (fam == 0) { if (!full_g) continue; c0 = L.c_g; len = L.len_g; ym = 2. * shear_g; damp = dkappa; }
      else if (fam == 1) { if (!full_g) continue; c0 = L.c_pol; len = L.len_pol; ym = p2; damp = dkappa; }
      else { if (!(has_ur && !ap.ufa_on)) continue; c0 = L.c_ur; len = L.len_ur; ym = 2. * shear_ur; damp = 0.; }
      const double* __restrict__ yc = y + c0;
      double* __restrict__ dc = dy + c0;
      LN_UNROLL
      for (int p = 0; p < len - 1; p++) {
        const int l = 3 + p;
        const double ylm = (p == 0) ? ym : yc[p - 1];
        dc[p] = k * LN_LDG(i2l1 + l) * (l * ylm - (l + 1) * yc[p + 1]) - damp * yc[p];
      }
      {
        const int p = len - 1, l = 3 + p;
        const double ylm = (p == 0) ? ym : yc[p - 1];
        dc[p] = k * (ylm - (1. + l) * cotKgen * yc[p]) - damp * yc[p];
      }
    }
    if (P.has_ncdm && !ap.ncdmfa_on) {
      const int len = L.len_ncdm;
      for (int j = 0; j < nqt; j++) {
        const double* __restrict__ yc = y + L.c_ncdm + j * len;
        double* __restrict__ dc = dy + L.c_ncdm + j * len;
        const double qk = k * nw[j];
        const double ym = y[L.psi0_ncdm1 + 3 * j + 2];
        LN_UNROLL
        for (int p = 0; p < len - 1; p++) {
          const int l = 3 + p;
          const double ylm = (p == 0) ? ym : yc[p - 1];
          dc[p] = qk * LN_LDG(i2l1 + l) * (l * ylm - (l + 1.) * yc[p + 1]);
        }
        {
          const int p = len - 1, l = 3 + p;
          const double ylm = (p == 0) ? ym : yc[p - 1];
          dc[p] = qk * ylm - (1. + l) * k * cotKgen * yc[p];
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// perturb_sources_member: source functions at sample index_tau from (y, dy)
LN_NOINLINE void ln_write_sources(const PtParams& P, Lane& M, double* __restrict__ mem, double tau, const double* __restrict__ y,
                                  const double* __restrict__ dy, int index_tau) {
  ln_env(P, M, mem, tau, 1);
  ln_rhs(P, M, y, nullptr, mem + P.lo_nw, LR_MATTER);
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
    delta_g = y[L.delta_g];
    if (!ap.tca_off) Pi = 5. * M.tca_shear_last / 8.;
    else Pi = (y[L.pol0_g] + y[L.pol0_g + 2] + 2. * y[L.shear_g]) / 8.;
  }
  const size_t stride_tp = (size_t)M.C->k_size * M.C->tau_size;
  double* out = M.C->sources + (size_t)M.ik_index * M.C->tau_size + index_tau;
  const double eta = y[L.eta];
  if (P.tp_t0 >= 0) {
    int switch_isw = 1;
    if ((P.switch_eisw == 0) && (z >= P.eisw_lisw_split_z)) switch_isw = 0;
    if ((P.switch_lisw == 0) && (z < P.eisw_lisw_split_z)) switch_isw = 0;
    const double theta_b = y[L.theta_b], dtheta_b = dy[L.theta_b];
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
LN_FN void ln_chain_rows(double* __restrict__ jd, double* __restrict__ jl, double* __restrict__ ju, const double* __restrict__ i2l1,
                         int len, double kk, double first, double damp, double last_diag) {
  LN_UNROLL
  for (int p = 0; p < len - 1; p++) {
    const int l = 3 + p;
    const double c = kk * LN_LDG(i2l1 + l);
    jl[p] = (p == 0 ? first : 1.) * c * l;
    ju[p] = -c * (l + 1);
    jd[p] = -damp;
  }
  const int p = len - 1;
  jl[p] = (p == 0 ? first : 1.) * kk;
  ju[p] = 0.;
  jd[p] = last_diag - damp;
}

LN_FN void ln_jacobian_chains(const PtParams& P, Lane& M, double* __restrict__ mem) {
  const LnLayout& L = M.L;
  const LnEnv& e = M.e;
  const double k = M.k;
  const double cotKgen = e.inv_tau * M.ik;
  const double* __restrict__ i2l1 = P.i2l1;
  double *jd = LVP(LV_JD), *jl = LVP(LV_JL), *ju = LVP(LV_JU);
  const double* nw = mem + P.lo_nw;
  // ---- chains
  int c = 0;
  if (L.c_g >= 0) {
    ln_chain_rows(jd + L.c_g, jl + L.c_g, ju + L.c_g, i2l1, L.len_g, k, 2., e.dkappa, -k * (1. + (2 + L.len_g)) * cotKgen);
    LCH_JUR(c) = -0.3 * k; c++;
    ln_chain_rows(jd + L.c_pol, jl + L.c_pol, ju + L.c_pol, i2l1, L.len_pol, k, 1., e.dkappa, -k * (1. + (2 + L.len_pol)) * cotKgen);
    LCH_JUR(c) = -0.6 * k; c++;
  }
  if (L.c_ur >= 0) {
    ln_chain_rows(jd + L.c_ur, jl + L.c_ur, ju + L.c_ur, i2l1, L.len_ur, k, 2., 0., -k * (1. + (2 + L.len_ur)) * cotKgen);
    LCH_JUR(c) = -0.3 * k; c++;
  }
  if (L.c_ncdm >= 0) {
    const int len = L.len_ncdm;
    for (int j = 0; j < P.nq_tot; j++) {
      const int c0 = L.c_ncdm + j * len;
      const double qk = k * nw[j];
      ln_chain_rows(jd + c0, jl + c0, ju + c0, i2l1, len, qk, 1., 0., -(1. + (2 + len)) * k * cotKgen);
      LCH_JUR(c) = -0.6 * qk; c++;
    }
  }
}

LN_NOINLINE void ln_jacobian(const PtParams& P, Lane& M, double* __restrict__ mem) {
  const LnLayout& L = M.L;
  const int n = L.neq, nh = L.nh;
  const double* nw = mem + P.lo_nw;
  ln_jacobian_chains(P, M, mem);
  // ---- hub block by probing (hub rows only; the chain parts of the probe vector stay zero)
  double* __restrict__ ej = LVP(LV_TMP);
  double* __restrict__ col = LVP(LV_DEL);
  double* __restrict__ jhh = mem + P.lo_jhh;
  LN_UNROLL
  for (int i = 0; i < n; i++) ej[i] = 0.;
  for (int j = 0; j < nh; j++) {
    ej[j] = 1.;
    ln_rhs(P, M, ej, col, nw, LR_HUB);
    LN_UNROLL
    for (int i = 0; i < nh; i++) jhh[i * nh + j] = col[i];
    ej[j] = 0.;
  }
  M.st.jacobians++;
  M.st.fevals += nh;
}

// Factorisation of A = I - c J: chains (backward elimination towards their root), Schur complement on the root
// diagonals, LU with partial pivoting of the hub block (row-major at lo_lu, 1/pivot on the diagonal).
LN_FN double ln_chain_factor(const double* __restrict__ jd, const double* __restrict__ jl, const double* __restrict__ ju,
                             double* __restrict__ ip, double* __restrict__ mu, int len, double c, double jur, double& mur_out) {
  const int last = len - 1;
  double ipn = 1.0 / (1.0 - c * jd[last]);
  ip[last] = ipn;
  double lon = -c * jl[last];
  for (int i = last - 1; i >= 0; i--) {
    const double mui = -c * ju[i] * ipn;
    const double p = (1.0 - c * jd[i]) - mui * lon;
    ipn = 1.0 / p;
    lon = -c * jl[i];
    ip[i] = ipn; mu[i] = mui;
  }
  const double mur = -c * jur * ipn;
  mur_out = mur;
  return -mur * lon;  // Schur term on the root diagonal
}

LN_NOINLINE void ln_factor(const PtParams& P, Lane& M, double* __restrict__ mem, double c) {
  const LnLayout& L = M.L;
  const int nh = L.nh, nch = L.nch;
  M.fac_c = c;
  const double* __restrict__ jhh = mem + P.lo_jhh;
  double* __restrict__ lu = mem + P.lo_lu;
  double* __restrict__ piv = mem + P.lo_piv;
  for (int i = 0; i < nh; i++) {
    LN_UNROLL
    for (int j = 0; j < nh; j++) lu[i * nh + j] = (i == j ? 1.0 : 0.0) - c * jhh[i * nh + j];
  }
  for (int ch = 0; ch < nch; ch++) {
    int s, len, root;
    ln_chain(L, ch, s, len, root);
    double mur;
    const double schur = ln_chain_factor(LVP(LV_JD) + s, LVP(LV_JL) + s, LVP(LV_JU) + s, LVP(LV_IP) + s, LVP(LV_MU) + s, len, c,
                                         LCH_JUR(ch), mur);
    LCH_MUR(ch) = mur;
    lu[root * nh + root] += schur;
  }
  // LU, partial pivoting
  for (int j = 0; j < nh; j++) {
    double best = fabs(lu[j * nh + j]);
    int bi = j;
    for (int i = j + 1; i < nh; i++) {
      const double v = fabs(lu[i * nh + j]);
      if (v > best) { best = v; bi = i; }
    }
    piv[j] = (double)bi;
    if (bi != j) {
      for (int cc = 0; cc < nh; cc++) {
        const double t = lu[j * nh + cc];
        lu[j * nh + cc] = lu[bi * nh + cc];
        lu[bi * nh + cc] = t;
      }
    }
    double pv = lu[j * nh + j];
    if (pv == 0.) pv = 1e-50;  // TINY, as ludcmp does for a singular pivot
    const double pinv = 1.0 / pv;
    lu[j * nh + j] = pinv;
    const double* __restrict__ rj = lu + j * nh;
    for (int i = j + 1; i < nh; i++) {
      double* ri = lu + i * nh;
      const double f = ri[j] * pinv;
      ri[j] = f;
      if (f != 0.) {
        LN_UNROLL
        for (int cc = j + 1; cc < nh; cc++) ri[cc] -= f * rj[cc];
      }
    }
  }
  M.st.factorizations++;
}

// solve A x = b in place with the factors of ln_factor
LN_NOINLINE void ln_solve(const PtParams& P, Lane& M, double* __restrict__ mem, double* __restrict__ b) {
  const LnLayout& L = M.L;
  const int nh = L.nh, nch = L.nch;
  const double c = M.fac_c;
  const double* __restrict__ lu = mem + P.lo_lu;
  const double* __restrict__ piv = mem + P.lo_piv;
  const double* __restrict__ mu = LVP(LV_MU);
  const double* __restrict__ ip = LVP(LV_IP);
  const double* __restrict__ jl = LVP(LV_JL);
  for (int ch = 0; ch < nch; ch++) {
    int s, len, root;
    ln_chain(L, ch, s, len, root);
    const int last = s + len - 1;
    double r = b[last];
    LN_UNROLL
    for (int i = last - 1; i >= s; i--) {
      r = b[i] - mu[i] * r;
      b[i] = r;
    }
    b[root] -= LCH_MUR(ch) * r;
  }
  for (int j = 0; j < nh; j++) {
    const int p = (int)piv[j];
    if (p != j) { const double t = b[j]; b[j] = b[p]; b[p] = t; }
  }
  for (int i = 1; i < nh; i++) {
    double s0 = b[i], s1 = 0.;
    const double* __restrict__ ri = lu + i * nh;
    int j = 0;
    for (; j + 1 < i; j += 2) { s0 -= ri[j] * b[j]; s1 -= ri[j + 1] * b[j + 1]; }
    if (j < i) s0 -= ri[j] * b[j];
    b[i] = s0 + s1;
  }
  for (int i = nh - 1; i >= 0; i--) {
    double s0 = b[i], s1 = 0.;
    const double* __restrict__ ri = lu + i * nh;
    int j = i + 1;
    for (; j + 1 < nh; j += 2) { s0 -= ri[j] * b[j]; s1 -= ri[j + 1] * b[j + 1]; }
    if (j < nh) s0 -= ri[j] * b[j];
    b[i] = (s0 + s1) * ri[i];
  }
  for (int ch = 0; ch < nch; ch++) {
    int s, len, root;
    ln_chain(L, ch, s, len, root);
    const int last = s + len - 1;
    double xp = b[root];
    LN_UNROLL
    for (int i = s; i <= last; i++) {
      const double lo = -c * jl[i];
      xp = (b[i] - lo * xp) * ip[i];
      b[i] = xp;
    }
  }
  M.st.solves++;
}

// -------------------------------------------------------------------------------------------------
// STRUCTURED hub solve.  With the metric scalars s = (h', eta', alpha', eta) held fixed, a hub row only depends on the
// variables of its own physical block (photon-baryon plasma; cdm; ur; each ncdm momentum bin / fluid; eta), so the hub
// block of the Jacobian is  J_hh = D + U V,  D block diagonal (blocks of <= 8), U = d(rows)/ds (nh x 4), V = ds/dy (4 x nh).
// The Newton matrix  A = (I - cD + Schur terms of the chains) - c U V  is solved by block LU + a 4x4 capacitance matrix
// (Sherman-Morrison-Woodbury):  x = w + Z t,  w = B^-1 b,  Z = B^-1 U,  (I - c V Z) t = c V w.
// A refactorisation costs O(nh) instead of O(nh^3/3), a solve O(nh) instead of O(nh^2) -- what keeps one refactorisation
// per step attempt (some lane of a warp changes its step size at almost every attempt) affordable.
// Slab layout (doubles from lo_jhh): DB[nh][8] | Uc[4][nh] | V[4][nh] | Zc[4][nh] | BL[nh][8] | bpiv[nh] | Minv[16] | active[4]
#define LN_SB 8
LN_FN void ln_block(const LnLayout& L, int i, int& b0, int& bs) {
  if (i < L.delta_cdm) { b0 = 0; bs = L.delta_cdm; }
  else if (i == L.delta_cdm) { b0 = i; bs = 1; }
  else if (L.delta_ur >= 0 && i <= L.shear_ur) { b0 = L.delta_ur; bs = 3; }
  else if (i < L.eta) { b0 = L.psi0_ncdm1 + 3 * ((i - L.psi0_ncdm1) / 3); bs = 3; }
  else { b0 = i; bs = 1; }
}
#define LS_DB(P) (mem + (P).lo_jhh)
#define LS_UC(P, nh) (LS_DB(P) + LN_SB * (nh))
#define LS_V(P, nh) (LS_UC(P, nh) + 4 * (nh))
#define LS_ZC(P, nh) (LS_V(P, nh) + 4 * (nh))
#define LS_BL(P, nh) (LS_ZC(P, nh) + 4 * (nh))
#define LS_BP(P, nh) (LS_BL(P, nh) + LN_SB * (nh))
#define LS_MI(P, nh) (LS_BP(P, nh) + (nh))
#define LS_ACT(P, nh) (LS_MI(P, nh) + 16)

LN_NOINLINE void ln_jacobian_s(const PtParams& P, Lane& M, double* __restrict__ mem) {
  const LnLayout& L = M.L;
  const int n = L.neq, nh = L.nh;
  const double* nw = mem + P.lo_nw;
  ln_jacobian_chains(P, M, mem);
  double* __restrict__ ej = LVP(LV_TMP);
  double* __restrict__ col = LVP(LV_DEL);
  double* __restrict__ DB = LS_DB(P);
  double* __restrict__ Uc = LS_UC(P, nh);
  double* __restrict__ V = LS_V(P, nh);
  double* __restrict__ act = LS_ACT(P, nh);
  double ms[MS_COUNT] = {0., 0., 0., 0.};
  LN_UNROLL
  for (int i = 0; i < n; i++) ej[i] = 0.;
  // D: one probe per position inside the blocks (all blocks at once: they do not see each other at fixed metric)
  const int bs_max = L.delta_cdm > 3 ? L.delta_cdm : 3;
  for (int g = 0; g < LN_SB; g++) {
    if (g < bs_max) {
      for (int i = 0; i < nh; i++) {
        int b0, bs;
        ln_block(L, i, b0, bs);
        ej[i] = (i - b0 == g) ? 1. : 0.;
      }
      ln_rhs(P, M, ej, col, nw, LR_HUB | LR_GIVEN_METRIC, ms);
      for (int i = 0; i < nh; i++) {
        int b0, bs;
        ln_block(L, i, b0, bs);
        DB[i * LN_SB + g] = (g < bs) ? col[i] : 0.;
      }
    } else {
      for (int i = 0; i < nh; i++) DB[i * LN_SB + g] = 0.;
    }
  }
  for (int i = 0; i < nh; i++) ej[i] = 0.;
  // U: response of the rows to each metric scalar
  for (int mm = 0; mm < MS_COUNT; mm++) {
    ms[mm] = 1.;
    ln_rhs(P, M, ej, col, nw, LR_HUB | LR_GIVEN_METRIC, ms);
    ms[mm] = 0.;
    double any = 0.;
    for (int i = 0; i < nh; i++) { Uc[mm * nh + i] = col[i]; any = fmax(any, fabs(col[i])); }
    act[mm] = any > 0. ? 1. : 0.;
  }
  // V: the metric scalars as functionals of the hub variables
  for (int j = 0; j < nh; j++) {
    ej[j] = 1.;
    ln_rhs(P, M, ej, nullptr, nw, 0);
    ej[j] = 0.;
    V[0 * nh + j] = M.m.h_prime; V[1 * nh + j] = M.m.eta_prime; V[2 * nh + j] = M.m.alpha_prime;
    V[3 * nh + j] = (j == L.eta) ? 1. : 0.;
  }
  M.st.jacobians++;
  M.st.fevals += bs_max + MS_COUNT;  // (the nh metric probes are ~1/10 of an RHS each)
}

// in-place LU with partial pivoting of one diagonal block (rows r0..r0+bs-1 of BL, LN_SB columns each)
LN_FN void ln_block_lu(double* __restrict__ BL, double* __restrict__ bpiv, int r0, int bs) {
  for (int j = 0; j < bs; j++) {
    double best = fabs(BL[(r0 + j) * LN_SB + j]);
    int bi = j;
    for (int i = j + 1; i < bs; i++) {
      const double v = fabs(BL[(r0 + i) * LN_SB + j]);
      if (v > best) { best = v; bi = i; }
    }
    bpiv[r0 + j] = (double)bi;
    if (bi != j) {
      for (int cc = 0; cc < bs; cc++) {
        const double t = BL[(r0 + j) * LN_SB + cc];
        BL[(r0 + j) * LN_SB + cc] = BL[(r0 + bi) * LN_SB + cc];
        BL[(r0 + bi) * LN_SB + cc] = t;
      }
    }
    double pv = BL[(r0 + j) * LN_SB + j];
    if (pv == 0.) pv = 1e-50;
    const double pinv = 1.0 / pv;
    BL[(r0 + j) * LN_SB + j] = pinv;
    for (int i = j + 1; i < bs; i++) {
      const double f = BL[(r0 + i) * LN_SB + j] * pinv;
      BL[(r0 + i) * LN_SB + j] = f;
      for (int cc = j + 1; cc < bs; cc++) BL[(r0 + i) * LN_SB + cc] -= f * BL[(r0 + j) * LN_SB + cc];
    }
  }
}
// x <- B^-1 x for the block-diagonal B factorised by ln_block_lu
LN_FN void ln_blocks_solve(const LnLayout& L, const double* __restrict__ BL, const double* __restrict__ bpiv, double* __restrict__ x) {
  const int nh = L.nh;
  int r0 = 0;
  while (r0 < nh) {
    int b0, bs;
    ln_block(L, r0, b0, bs);
    if (bs == 1) {
      x[r0] *= BL[r0 * LN_SB];
    } else {
      for (int j = 0; j < bs; j++) {
        const int p = (int)bpiv[r0 + j];
        if (p != j) { const double t = x[r0 + j]; x[r0 + j] = x[r0 + p]; x[r0 + p] = t; }
      }
      for (int i = 1; i < bs; i++) {
        double s = x[r0 + i];
        for (int j = 0; j < i; j++) s -= BL[(r0 + i) * LN_SB + j] * x[r0 + j];
        x[r0 + i] = s;
      }
      for (int i = bs - 1; i >= 0; i--) {
        double s = x[r0 + i];
        for (int j = i + 1; j < bs; j++) s -= BL[(r0 + i) * LN_SB + j] * x[r0 + j];
        x[r0 + i] = s * BL[(r0 + i) * LN_SB + i];
      }
    }
    r0 += bs;
  }
}

LN_NOINLINE void ln_factor_s(const PtParams& P, Lane& M, double* __restrict__ mem, double c) {
  const LnLayout& L = M.L;
  const int nh = L.nh, nch = L.nch;
  M.fac_c = c;
  const double* __restrict__ DB = LS_DB(P);
  const double* __restrict__ Uc = LS_UC(P, nh);
  const double* __restrict__ V = LS_V(P, nh);
  double* __restrict__ Zc = LS_ZC(P, nh);
  double* __restrict__ BL = LS_BL(P, nh);
  double* __restrict__ bpiv = LS_BP(P, nh);
  double* __restrict__ Mi = LS_MI(P, nh);
  const double* __restrict__ act = LS_ACT(P, nh);
  for (int i = 0; i < nh; i++) {
    int b0, bs;
    ln_block(L, i, b0, bs);
#pragma unroll
    for (int cc = 0; cc < LN_SB; cc++) BL[i * LN_SB + cc] = ((b0 + cc == i) ? 1.0 : 0.0) - c * DB[i * LN_SB + cc];
  }
  for (int ch = 0; ch < nch; ch++) {
    int s, len, root;
    ln_chain(L, ch, s, len, root);
    double mur;
    const double schur = ln_chain_factor(LVP(LV_JD) + s, LVP(LV_JL) + s, LVP(LV_JU) + s, LVP(LV_IP) + s, LVP(LV_MU) + s, len, c,
                                         LCH_JUR(ch), mur);
    LCH_MUR(ch) = mur;
    int b0, bs;
    ln_block(L, root, b0, bs);
    BL[root * LN_SB + (root - b0)] += schur;
  }
  {
    int r0 = 0;
    while (r0 < nh) {
      int b0, bs;
      ln_block(L, r0, b0, bs);
      if (bs == 1) BL[r0 * LN_SB] = 1.0 / BL[r0 * LN_SB];
      else ln_block_lu(BL, bpiv, r0, bs);
      r0 += bs;
    }
  }
  // Z = B^-1 U and the capacitance matrix G = I - c V Z (inactive metric scalars: identity rows/columns)
  double G[MS_COUNT][MS_COUNT];
  for (int mm = 0; mm < MS_COUNT; mm++) {
    if (act[mm] != 0.) {
      double* __restrict__ z = Zc + mm * nh;
      const double* __restrict__ u = Uc + mm * nh;
      for (int i = 0; i < nh; i++) z[i] = u[i];
      ln_blocks_solve(L, BL, bpiv, z);
    }
  }
#pragma unroll
  for (int a = 0; a < MS_COUNT; a++)
#pragma unroll
    for (int b = 0; b < MS_COUNT; b++) {
      double sum = 0.;
      if (act[a] != 0. && act[b] != 0.) {
        const double* __restrict__ v = V + a * nh;
        const double* __restrict__ z = Zc + b * nh;
        LN_UNROLL
        for (int j = 0; j < nh; j++) sum += v[j] * z[j];
      }
      G[a][b] = (a == b ? 1.0 : 0.0) - c * sum;
    }
  // invert G (Gauss-Jordan, partial pivoting; 4 x 4 in registers)
  double Inv[MS_COUNT][MS_COUNT];
#pragma unroll
  for (int a = 0; a < MS_COUNT; a++)
#pragma unroll
    for (int b = 0; b < MS_COUNT; b++) Inv[a][b] = (a == b) ? 1. : 0.;
#pragma unroll
  for (int j = 0; j < MS_COUNT; j++) {
    int p = j;
    double best = fabs(G[j][j]);
#pragma unroll
    for (int i = 0; i < MS_COUNT; i++)
      if (i > j && fabs(G[i][j]) > best) { best = fabs(G[i][j]); p = i; }
#pragma unroll
    for (int i = 0; i < MS_COUNT; i++) {
      if (i > j && i == p) {
#pragma unroll
        for (int b = 0; b < MS_COUNT; b++) {
          double t = G[j][b]; G[j][b] = G[i][b]; G[i][b] = t;
          t = Inv[j][b]; Inv[j][b] = Inv[i][b]; Inv[i][b] = t;
        }
      }
    }
    double pv = G[j][j];
    if (pv == 0.) pv = 1e-50;
    const double pinv = 1.0 / pv;
#pragma unroll
    for (int b = 0; b < MS_COUNT; b++) { G[j][b] *= pinv; Inv[j][b] *= pinv; }
#pragma unroll
    for (int i = 0; i < MS_COUNT; i++) {
      if (i != j) {
        const double f = G[i][j];
#pragma unroll
        for (int b = 0; b < MS_COUNT; b++) { G[i][b] -= f * G[j][b]; Inv[i][b] -= f * Inv[j][b]; }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < MS_COUNT; a++)
#pragma unroll
    for (int b = 0; b < MS_COUNT; b++) Mi[a * MS_COUNT + b] = Inv[a][b];
  M.st.factorizations++;
}

LN_NOINLINE void ln_solve_s(const PtParams& P, Lane& M, double* __restrict__ mem, double* __restrict__ b) {
  const LnLayout& L = M.L;
  const int nh = L.nh, nch = L.nch;
  const double c = M.fac_c;
  const double* __restrict__ V = LS_V(P, nh);
  const double* __restrict__ Zc = LS_ZC(P, nh);
  const double* __restrict__ BL = LS_BL(P, nh);
  const double* __restrict__ bpiv = LS_BP(P, nh);
  const double* __restrict__ Mi = LS_MI(P, nh);
  const double* __restrict__ act = LS_ACT(P, nh);
  const double* __restrict__ mu = LVP(LV_MU);
  const double* __restrict__ ip = LVP(LV_IP);
  const double* __restrict__ jl = LVP(LV_JL);
  for (int ch = 0; ch < nch; ch++) {
    int s, len, root;
    ln_chain(L, ch, s, len, root);
    const int last = s + len - 1;
    double r = b[last];
    LN_UNROLL
    for (int i = last - 1; i >= s; i--) {
      r = b[i] - mu[i] * r;
      b[i] = r;
    }
    b[root] -= LCH_MUR(ch) * r;
  }
  ln_blocks_solve(L, BL, bpiv, b);
  double g[MS_COUNT], t[MS_COUNT];
#pragma unroll
  for (int mm = 0; mm < MS_COUNT; mm++) {
    double sum = 0.;
    if (act[mm] != 0.) {
      const double* __restrict__ v = V + mm * nh;
      LN_UNROLL
      for (int j = 0; j < nh; j++) sum += v[j] * b[j];
    }
    g[mm] = c * sum;
  }
#pragma unroll
  for (int a = 0; a < MS_COUNT; a++) {
    double sum = 0.;
#pragma unroll
    for (int bb = 0; bb < MS_COUNT; bb++) sum += Mi[a * MS_COUNT + bb] * g[bb];
    t[a] = sum;
  }
#pragma unroll
  for (int mm = 0; mm < MS_COUNT; mm++) {
    if (act[mm] != 0.) {
      const double* __restrict__ z = Zc + mm * nh;
      const double tm = t[mm];
      LN_UNROLL
      for (int j = 0; j < nh; j++) b[j] += z[j] * tm;
    }
  }
  for (int ch = 0; ch < nch; ch++) {
    int s, len, root;
    ln_chain(L, ch, s, len, root);
    const int last = s + len - 1;
    double xp = b[root];
    LN_UNROLL
    for (int i = s; i <= last; i++) {
      const double lo = -c * jl[i];
      xp = (b[i] - lo * xp) * ip[i];
      b[i] = xp;
    }
  }
  M.st.solves++;
}

// dense or structured hub algebra (P.ln_structured; the dense path remains as the cross-check)
LN_FN void ln_do_jacobian(const PtParams& P, Lane& M, double* __restrict__ mem) {
  if (P.ln_structured) ln_jacobian_s(P, M, mem); else ln_jacobian(P, M, mem);
}
LN_FN void ln_do_factor(const PtParams& P, Lane& M, double* __restrict__ mem, double c) {
  if (P.ln_structured) ln_factor_s(P, M, mem, c); else ln_factor(P, M, mem, c);
}
LN_FN void ln_do_solve(const PtParams& P, Lane& M, double* __restrict__ mem, double* __restrict__ b) {
  if (P.ln_structured) ln_solve_s(P, M, mem, b); else ln_solve(P, M, mem, b);
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
  double* __restrict__ d0 = LVP(LV_DIF0);
  double* __restrict__ d1 = LVP(LV_DIF0 + 1);
  double* __restrict__ d2 = LVP(LV_DIF0 + 2);
  double* __restrict__ d3 = LVP(LV_DIF0 + 3);
  double* __restrict__ d4 = LVP(LV_DIF0 + 4);
  LN_UNROLL
  for (int i = 0; i < n; i++) {
    double row[5];
    row[0] = d0[i];
    row[1] = (1 < kord) ? d1[i] : 0.;
    row[2] = (2 < kord) ? d2[i] : 0.;
    row[3] = (3 < kord) ? d3[i] : 0.;
    row[4] = (4 < kord) ? d4[i] : 0.;
    double o[5];
#pragma unroll
    for (int jj = 0; jj < 5; jj++) {
      double s = 0.0;
#pragma unroll
      for (int kk = 0; kk < 5; kk++) s += row[kk] * RU[kk][jj];
      o[jj] = s;
    }
    d0[i] = o[0];
    if (1 < kord) d1[i] = o[1];
    if (2 < kord) d2[i] = o[2];
    if (3 < kord) d3[i] = o[3];
    if (4 < kord) d4[i] = o[4];
  }
}

// ---- vector loops of the NDF step (restrict pointers: see LN_UNROLL above)
LN_FN double ln_predict(const double* __restrict__ y, const double* __restrict__ dif, int np, int k, int n, double* __restrict__ psi,
                        double* __restrict__ pred, double* __restrict__ ynew, double* __restrict__ difkp1,
                        double* __restrict__ invwt, double threshold, double eps) {
  const double invGak = c_invGa[k - 1];
  double g[5];
#pragma unroll
  for (int j = 0; j < 5; j++) g[j] = c_G[j] * invGak;
  double minnrm = 0.0;
  LN_UNROLL
  for (int i = 0; i < n; i++) {
    const double yi = y[i];
    double ps = 0.0, pr = yi;
#pragma unroll
    for (int j = 0; j < 5; j++) {
      if (j < k) {
        const double d = dif[j * np + i];
        ps += d * g[j];
        pr += d;
      }
    }
    psi[i] = ps;
    pred[i] = pr;
    ynew[i] = pr;
    difkp1[i] = 0.0;
    const double iw = 1.0 / fmax(fmax(fabs(pr), fabs(yi)), threshold);
    invwt[i] = iw;
    minnrm = fmax(minnrm, 100 * eps * fabs(pr * iw));
  }
  return minnrm;
}

LN_FN void ln_residual(double* __restrict__ del, const double* __restrict__ f, const double* __restrict__ psi,
                       const double* __restrict__ difkp1, double hinvGak, int n) {
  LN_UNROLL
  for (int i = 0; i < n; i++) del[i] = hinvGak * f[i] - (psi[i] + difkp1[i]);
}

LN_FN double ln_newton_update(const double* __restrict__ del, const double* __restrict__ invwt, const double* __restrict__ pred,
                              double* __restrict__ difkp1, double* __restrict__ ynew, int n) {
  double newnrm = 0.0;
  LN_UNROLL
  for (int i = 0; i < n; i++) {
    const double d = del[i];
    newnrm = fmax(newnrm, fabs(d * invwt[i]));
    const double dk = difkp1[i] + d;
    difkp1[i] = dk;
    ynew[i] = pred[i] + dk;
  }
  return newnrm;
}

LN_FN double ln_wnorm(const double* __restrict__ v, const double* __restrict__ invwt, int n) {
  double e = 0.0;
  LN_UNROLL
  for (int i = 0; i < n; i++) e = fmax(e, fabs(v[i] * invwt[i]));
  return e;
}

LN_FN double ln_wnorm2(const double* __restrict__ v, const double* __restrict__ w, const double* __restrict__ invwt, int n) {
  double e = 0.0;
  LN_UNROLL
  for (int i = 0; i < n; i++) e = fmax(e, fabs((v[i] + w[i]) * invwt[i]));
  return e;
}

// dif[k+1] = dk - dif[k]; dif[k] = dk; dif[j] += dif[j+1] (j < k); also the norms of the new dif[k-1] and dif[k+1].
// The column of differences of element i is loaded into registers first: a load-add-store recurrence through memory
// (stores to dif[] may alias the next load as far as the compiler knows) costs one memory round trip per order.
LN_FN void ln_dif_update(double* __restrict__ dif, int np, int k, int n, const double* __restrict__ difkp1,
                         const double* __restrict__ invwt, double& e_km1, double& e_kp1) {
  double a = 0., b = 0.;
#pragma unroll 2
  for (int i = 0; i < n; i++) {
    const double dk = difkp1[i], iw = invwt[i];
    double d[6];
#pragma unroll
    for (int j = 0; j < 6; j++) d[j] = (j <= k) ? dif[j * np + i] : 0.;
    double dkold = d[0];
#pragma unroll
    for (int j = 1; j < 6; j++) dkold = (j == k) ? d[j] : dkold;
    const double dkp1 = dk - dkold;
    double acc = dk;
#pragma unroll
    for (int j = 5; j >= 0; j--) {
      if (j == k) d[j] = dk;
      else if (j < k) { acc += d[j]; d[j] = acc; }
      if (j == k - 1) a = fmax(a, fabs(d[j] * iw));
    }
    dif[(k + 1) * np + i] = dkp1;
#pragma unroll
    for (int j = 0; j < 6; j++)
      if (j <= k) dif[j * np + i] = d[j];
    b = fmax(b, fabs(dkp1 * iw));
  }
  e_km1 = a; e_kp1 = b;
}

LN_FN void ln_copy(double* __restrict__ dst, const double* __restrict__ src, int n) {
  LN_UNROLL
  for (int i = 0; i < n; i++) dst[i] = src[i];
}

LN_FN void ln_interp(const double* __restrict__ dif, int np, int k, int n, const double* __restrict__ ynew, const double* c1,
                     const double* c2, double* __restrict__ yi, double* __restrict__ ypi) {
  LN_UNROLL
  for (int i = 0; i < n; i++) {
    double a1 = 0, a2 = 0;
    for (int j = 0; j < k; j++) {
      const double d = dif[j * np + i];
      a1 += c1[j] * d;
      a2 += c2[j] * d;
    }
    yi[i] = ynew[i] + a1;
    ypi[i] = a2;
  }
}

// -------------------------------------------------------------------------------------------------
// NDF1-5 over [t0, tfinal] for the current layout (evolver_ndf15.cpp:302-651, same step/order control), written as a
// flat state machine: every pass of the loop is ONE step attempt.
LN_NOINLINE bool ln_ndf15(const PtParams& P, Lane& M, double* __restrict__ mem, double t0, double tfinal) {
  const double eps = 1e-16, threshold = 1e-15;
  const int maxit = 4, maxk = 5;
  const double rtol = P.rtol;
  const int n = M.L.neq, np = P.np;
  double *Y = LVP(LV_Y), *YNEW = LVP(LV_YNEW), *F = LVP(LV_F), *PRED = LVP(LV_PRED), *PSI = LVP(LV_PSI),
         *DIFKP1 = LVP(LV_DIFKP1), *DEL = LVP(LV_DEL), *INVWT = LVP(LV_INVWT), *DIF = LVP(LV_DIF0);
  const double* NW = mem + P.lo_nw;
  const double* t_vec = M.C->tau;
  const int tres = M.C->tau_size;
  int next = M.next;
  while (next < tres && LN_LDG(t_vec + next) < t0) next++;
  double tnext = (next < tres) ? LN_LDG(t_vec + next) : 1e300;
  for (int j = 0; j < 7; j++) {
    double* __restrict__ dj = DIF + j * np;
    LN_UNROLL
    for (int i = 0; i < n; i++) dj[i] = 0.;
  }
  const double htspan = fabs(tfinal - t0);
  double t = t0, tnew = t0;
  ln_env(P, M, mem, t0, 0);
  ln_rhs(P, M, Y, F, NW, LR_HUB | LR_CHAINS);
  M.st.fevals++;
  const double hmax = (tfinal - t0) / 10.0;
  ln_do_jacobian(P, M, mem);
  bool Jcurrent = true;
  double hmin = 16.0 * eps * fabs(t);
  double rh = 0.0;
  for (int i = 0; i < n; i++) {
    const double wt = fmax(fabs(Y[i]), threshold);
    rh = fmax(rh, 1.25 / sqrt(rtol) * fabs(F[i] / wt));
  }
  double absh = fmin(hmax, htspan);
  if (absh * rh > 1.0) absh = 1.0 / rh;
  absh = fmax(absh, hmin);
  double h = absh;
  {
    ln_rhs(P, M, F, PSI, NW, LR_HUB | LR_CHAINS);  // J*f0 = f(t0, f0): linear, homogeneous
    M.st.fevals++;
    const double tdel = (t + fmin(sqrt(eps) * fmax(fabs(t), fabs(t + h)), absh)) - t;
    ln_env(P, M, mem, t + tdel, 0);
    ln_rhs(P, M, Y, DEL, NW, LR_HUB | LR_CHAINS);
    M.st.fevals++;
    rh = 0.0;
    for (int i = 0; i < n; i++) {
      const double wt = fmax(fabs(Y[i]), threshold);
      const double s = PSI[i] + (DEL[i] - F[i]) / tdel;
      rh = fmax(rh, 1.25 * sqrt(0.5 * fabs(s / wt) / rtol));
    }
    absh = fmin(hmax, htspan);
    if (absh * rh > 1.0) absh = 1.0 / rh;
    absh = fmax(absh, hmin);
    h = absh;
  }
  int k = 1, klast = k;
  double abshlast = absh;
  for (int i = 0; i < n; i++) DIF[i] = h * F[i];
  double hinvGak = h * c_invGa[k - 1];
  int nconhk = 0;
  ln_do_factor(P, M, mem, hinvGak);
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
        ln_do_factor(P, M, mem, hinvGak);
        havrate = false;
      }
      nofailed = true;
      new_step = false;
    }
    // ---- one attempt
    tnew = t + h;
    if (done) tnew = tfinal;
    h = tnew - t;
    const double minnrm = ln_predict(Y, DIF, np, k, n, PSI, PRED, YNEW, DIFKP1, INVWT, threshold, eps);
    ln_env(P, M, mem, tnew, 0);
    bool gotynew = false;
    for (int iter = 1; iter <= maxit; iter++) {
      ln_rhs(P, M, YNEW, F, NW, LR_HUB | LR_CHAINS);
      if (!M.ap.tca_off) M.tca_shear_last = M.m.tca_shear_g;
      M.st.fevals++;
      ln_residual(DEL, F, PSI, DIFKP1, hinvGak, n);
      ln_do_solve(P, M, mem, DEL);
      const double newnrm = ln_newton_update(DEL, INVWT, PRED, DIFKP1, YNEW, n);
      if (newnrm <= minnrm) { gotynew = true; break; }
      else if (iter == 1) {
        if (havrate) {
          const double errit = newnrm * rate / (1.0 - rate);
          if (errit <= 0.05 * rtol) { gotynew = true; break; }
        } else {
          rate = 0.0;
        }
      } else if (newnrm > 0.9 * oldnrm) {
        break;  // too slow
      } else {
        rate = fmax(0.9 * rate, newnrm / oldnrm);
        havrate = true;
        const double errit = newnrm * rate / (1.0 - rate);
        if (errit <= 0.5 * rtol) { gotynew = true; break; }
        else if (iter == maxit) break;
        else {
          double rp = rate;  // rate^(maxit-iter)
          for (int q = 1; q < maxit - iter; q++) rp *= rate;
          if (0.5 * rtol < errit * rp) break;
        }
      }
      oldnrm = newnrm;
    }
    if (!gotynew) {  // Newton iteration too slow
      M.st.failed++;
      if (!Jcurrent) {
        ln_env(P, M, mem, t, 0);
        M.st.fevals++;  // the reference re-evaluates f(t, y) for numjac; the probes below need the environment only
        ln_do_jacobian(P, M, mem);
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
      ln_do_factor(P, M, mem, hinvGak);
      havrate = false;
      continue;
    }
    // ---- error estimate
    err = ln_wnorm(DIFKP1, INVWT, n) * c_erconst[k - 1];
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
          const double errkm1 = ln_wnorm2(DIF + (k - 1) * np, DIFKP1, INVWT, n) * c_erconst[k - 2];
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
      ln_do_factor(P, M, mem, hinvGak);
      havrate = false;
      continue;
    }
    // ---- step accepted
    M.st.steps++;
    double e_km1 = 0., e_kp1 = 0.;
    ln_dif_update(DIF, np, k, n, DIFKP1, INVWT, e_km1, e_kp1);
    // ---- output at the sample times passed by this step
    while ((next < tres) && ((tnew - tnext) >= 0.0)) {
      if (tnew == tnext) {
        ln_write_sources(P, M, mem, tnext, YNEW, F, next);
      } else {
        const double s = (tnext - tnew) / h;
        double c1[5], c2[5];
        double prod = 1.0, sumfrac = 0., fact = 1.0;
        for (int j = 0; j < k; j++) {
          prod *= (s + j);
          fact *= (j + 1);
          sumfrac += 1.0 / (s + j);
          c1[j] = prod / fact;
          c2[j] = prod * sumfrac / (h * fact);
        }
        ln_interp(DIF, np, k, n, YNEW, c1, c2, LVP(LV_TMP), LVP(LV_YPI));
        ln_write_sources(P, M, mem, tnext, LVP(LV_TMP), LVP(LV_YPI), next);
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
    t = tnew;
    { double* sw = Y; Y = YNEW; YNEW = sw; }  // y <- ynew without a pass over memory
    Jcurrent = false;
    new_step = true;
  }
  // final state (into the LV_Y slot, where the next interval expects it), and a last RHS call so that the environment
  // and the TCA/RSA by-products are current at the end of the interval (evolver_ndf15.cpp:653-662)
  if (YNEW != LVP(LV_Y)) ln_copy(LVP(LV_Y), YNEW, n);
  Y = LVP(LV_Y);
  ln_env(P, M, mem, tnew, 0);
  ln_rhs(P, M, Y, F, NW, LR_HUB | LR_CHAINS);
  if (!M.ap.tca_off) M.tca_shear_last = M.m.tca_shear_g;
  M.st.fevals++;
  M.next = next;
  return true;
}

// =================================================================================================
// TAIL: the radiation-streaming interval (RSA on, ncdm fluid on if any): N = 4 + 3 N_ncdm equations, no chains.  It is
// the last interval of every mode with k tau_0 > 45 and holds 70 % of all steps (10^4 .. 3x10^5 per high-k mode): the
// serial critical path of a launch.  Everything of the integrator lives in REGISTERS (fully unrolled static arrays:
// state, backward differences, step-control vectors); only the LU factors of the N x N Newton matrix sit in local
// memory (partial pivoting needs dynamic row indices).  The environment is evaluated straight into registers.
// State order (ln_make_layout with rsa_on): delta_b, theta_b, delta_cdm, [delta, theta, shear](ncdm s), eta.
// =================================================================================================
template <int NN>
struct LnTailEnv {
  double a2x15, aH, inv_half_aH, rho_b, rho_cdm, rho_g43, rho_ur43, dkappa, ddkappa, cb2, Rdk;
  double nf[NN > 0 ? NN : 1][8];
};

template <int NN>
LN_FN void ln_env_tail(const PtParams& P, Lane& M, double tau, LnTailEnv<NN>& E) {
  double av, Hv, rho_g, rho_b;
  double rho_n[NN > 0 ? NN : 1], p_n[NN > 0 ? NN : 1], pp_n[NN > 0 ? NN : 1];
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
    av = LN_BG(P.ia); Hv = LN_BG(P.iH);
    rho_g = LN_BG(P.irho_g); rho_b = LN_BG(P.irho_b); E.rho_cdm = LN_BG(P.irho_cdm);
    E.rho_ur43 = P.has_ur ? 4. / 3. * LN_BG(P.irho_ur) : 0.;
#pragma unroll
    for (int s = 0; s < NN; s++) {
      rho_n[s] = LN_BG(P.irho_ncdm1 + s); p_n[s] = LN_BG(P.ip_ncdm1 + s); pp_n[s] = LN_BG(P.ipseudo_p_ncdm1 + s);
    }
#undef LN_BG
  }
  const double inv_a = 1. / av;
  const double z = inv_a - 1.;
  if (z >= M.z_last) {
    const PtCosmo* C = M.C;
    const double* row = M.th_y + (size_t)(M.tt_size - 1) * P.th_size;
    const double xe0 = LN_LDG(row + P.ixe);
    E.dkappa = (1. + z) * (1. + z) * C->n_e * xe0 * CLPP_sigma * CLPP_Mpc_over_m;
    E.ddkappa = -Hv * 2. / (1. + z) * E.dkappa;
    const double wb = CLPP_k_B / (CLPP_c * CLPP_c * CLPP_m_H) * (1. + (1. / CLPP_not4 - 1.) * C->YHe + xe0 * (1. - C->YHe)) *
                      C->T_cmb * (1. + z);
    E.cb2 = wb * 4. / 3.;
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
    E.dkappa = LN_TH(P.idkappa); E.ddkappa = LN_TH(P.iddkappa); E.cb2 = LN_TH(P.icb2);
#undef LN_TH
  }
  const double aH = Hv * av;
  E.a2x15 = 1.5 * av * av;
  E.aH = aH;
  E.inv_half_aH = 2. / aH;
  E.rho_b = rho_b;
  E.rho_g43 = 4. / 3. * rho_g;
  E.Rdk = 4. / 3. * rho_g / rho_b * E.dkappa;
#pragma unroll
  for (int s = 0; s < NN; s++) {
    const double rho = rho_n[s], p = p_n[s], pseudo = pp_n[s];
    const double w_n = p / rho, pseudo_p_over_p = pseudo / p, i1w = rho / (rho + p);
    const double cg2 = w_n * (1.0 - i1w * (1. / 3.) * (3.0 * w_n - 2.0 + pseudo_p_over_p));
    const double ca2 = w_n * (1. / 3.) * i1w * (5.0 - pseudo_p_over_p);
    const double cvis2 = (P.ncdmfa_method == CLPP_NCDMFA_HU) ? w_n : 3. * w_n * ca2;
    const double damp = (P.ncdmfa_method == CLPP_NCDMFA_HU) ? 3.0 * aH * ca2 * (rho / p)
                                                            : 3.0 * (aH * (2. / 3. - ca2 - pseudo_p_over_p * (1. / 3.)) + 1.0 / tau);
    E.nf[s][0] = rho; E.nf[s][1] = rho + p; E.nf[s][2] = w_n; E.nf[s][3] = cg2 * rho; E.nf[s][4] = ca2; E.nf[s][5] = ca2 * i1w;
    E.nf[s][6] = 8.0 / 3.0 * cvis2 * i1w; E.nf[s][7] = damp;
  }
}

// f(tau, y) of the radiation-streaming interval (perturb_total_stress_energy / perturb_einstein /
// perturb_rsa_delta_and_theta / perturb_derivs_member restricted to this approximation set)
template <int NN>
LN_FN void ln_rhs_tail(const PtParams& P, const LnTailEnv<NN>& E, double k2, double ik2, const double (&y)[4 + 3 * NN],
                       double (&dy)[4 + 3 * NN]) {
  constexpr int N = 4 + 3 * NN;
  const double delta_b = y[0], theta_b = y[1], delta_cdm = y[2], eta = y[N - 1];
  double delta_rho = E.rho_b * delta_b + E.rho_cdm * delta_cdm;
  double rpt = E.rho_b * theta_b;
#pragma unroll
  for (int s = 0; s < NN; s++) {
    delta_rho += E.nf[s][0] * y[3 + 3 * s];
    rpt += E.nf[s][1] * y[4 + 3 * s];
  }
  const double aH = E.aH;
  const double h_prime = (k2 * eta + E.a2x15 * delta_rho) * E.inv_half_aH;
  double rsa_theta_g = 0.;
  if (P.rsa_method != CLPP_RSA_NULL) rsa_theta_g = -0.5 * h_prime;
  const double rsa_theta_ur = rsa_theta_g;  // before the reionisation correction
  if (P.rsa_method == CLPP_RSA_MD_WITH_REIO)
    rsa_theta_g += 3. * ik2 * (E.ddkappa * (theta_b + 0.5 * h_prime) +
                               E.dkappa * (-aH * theta_b + E.cb2 * k2 * delta_b - aH * h_prime + k2 * eta));
  rpt += E.rho_g43 * rsa_theta_g + E.rho_ur43 * rsa_theta_ur;
  const double eta_prime = (E.a2x15 * rpt) * ik2;
  const double alpha = (h_prime + 6. * eta_prime) * 0.5 * ik2;
  const double metric_continuity = 0.5 * h_prime;
  const double metric_shear = k2 * alpha;
  dy[0] = -(theta_b + metric_continuity);
  dy[1] = -aH * theta_b + k2 * E.cb2 * delta_b + E.Rdk * (rsa_theta_g - theta_b);
  dy[2] = -metric_continuity;
  dy[N - 1] = eta_prime;
#pragma unroll
  for (int s = 0; s < NN; s++) {
    const double* nf = E.nf[s];
    const double y0 = y[3 + 3 * s], y1 = y[4 + 3 * s], y2 = y[5 + 3 * s];
    const double w_n = nf[2], ca2 = nf[4];
    const double msn = (P.ncdmfa_method == CLPP_NCDMFA_CLASS) ? metric_continuity : metric_shear;
    dy[3 + 3 * s] = -(1.0 + w_n) * (y1 + metric_continuity) - 3.0 * aH * (ca2 - w_n) * y0;
    dy[4 + 3 * s] = -aH * (1.0 - 3.0 * ca2) * y1 + nf[5] * k2 * y0 - k2 * y2;
    dy[5 + 3 * s] = -nf[7] * y2 + nf[6] * (y1 + msn);
  }
}

// order-specialised pieces of the tail integrator (static register indices instead of select chains)
template <int N, int K>
LN_FN void ln_tail_predict(const double (&y)[N], const double (&dif)[7][N], double (&psi)[N], double (&pred)[N]) {
  const double invGak = c_invGa[K - 1];
#pragma unroll
  for (int i = 0; i < N; i++) {
    double ps = 0., pr = y[i];
#pragma unroll
    for (int j = 0; j < K; j++) {
      ps += dif[j][i] * (c_G[j] * invGak);
      pr += dif[j][i];
    }
    psi[i] = ps; pred[i] = pr;
  }
}
template <int N, int K>
LN_FN void ln_tail_accept(double (&dif)[7][N], const double (&dk1)[N], const double (&iw)[N], double& e_km1, double& e_kp1) {
  double a = 0., b = 0.;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const double dkp1 = dk1[i] - dif[K][i];
    dif[K + 1][i] = dkp1;
    dif[K][i] = dk1[i];
#pragma unroll
    for (int j = K - 1; j >= 0; j--) dif[j][i] += dif[j + 1][i];
    a = fmax(a, fabs(dif[K - 1][i] * iw[i]));
    b = fmax(b, fabs(dkp1 * iw[i]));
  }
  e_km1 = a; e_kp1 = b;
}
template <int N, int K>
LN_FN double ln_tail_errkm1(const double (&dif)[7][N], const double (&dk1)[N], const double (&iw)[N]) {
  double e = 0.;
#pragma unroll
  for (int i = 0; i < N; i++) e = fmax(e, fabs((dif[K - 1][i] + dk1[i]) * iw[i]));
  return e;
}
#define LN_TAIL_BY_ORDER(k, CALL)                      \
  switch (k) {                                         \
    case 1: { constexpr int K_ = 1; CALL; } break;     \
    case 2: { constexpr int K_ = 2; CALL; } break;     \
    case 3: { constexpr int K_ = 3; CALL; } break;     \
    case 4: { constexpr int K_ = 4; CALL; } break;     \
    default: { constexpr int K_ = 5; CALL; } break;    \
  }

template <int NN>
LN_NOINLINE bool ln_ndf15_tail(const PtParams& P, Lane& M, double* __restrict__ mem, double t0, double tfinal) {
  constexpr int N = 4 + 3 * NN;
  const double eps = 1e-16, threshold = 1e-15;
  const int maxit = 4, maxk = 5;
  const double rtol = P.rtol;
  const double k2 = M.k2, ik2 = M.ik2;
  const double* t_vec = M.C->tau;
  const int tres = M.C->tau_size;
  int next = M.next;
  while (next < tres && LN_LDG(t_vec + next) < t0) next++;
  double tnext = (next < tres) ? LN_LDG(t_vec + next) : 1e300;
  double y[N], dif[7][N], psi[N], pred[N], dk1[N], iw[N], f[N], tmp[N];
  double J[N * N], A[N * N];  // Jacobian and LU factors of I - c J (row-major; local memory: dynamic pivot rows)
  int piv[N];
  LnTailEnv<NN> E;
  {
    const double* Ys = LVP(LV_Y);
#pragma unroll
    for (int i = 0; i < N; i++) y[i] = Ys[i];
  }
#pragma unroll
  for (int j = 0; j < 7; j++)
#pragma unroll
    for (int i = 0; i < N; i++) dif[j][i] = 0.;

  auto jacobian = [&]() {  // columns f(e_j): the system is linear and homogeneous; the environment must be current
#pragma unroll 1
    for (int j = 0; j < N; j++) {
      double ej[N], col[N];
#pragma unroll
      for (int i = 0; i < N; i++) ej[i] = (i == j) ? 1. : 0.;
      ln_rhs_tail<NN>(P, E, k2, ik2, ej, col);
#pragma unroll
      for (int i = 0; i < N; i++) J[i * N + j] = col[i];
    }
    M.st.jacobians++;
    M.st.fevals += N;
  };
  auto factor = [&](double c) {
#pragma unroll 1
    for (int i = 0; i < N * N; i++) A[i] = -c * J[i];
#pragma unroll 1
    for (int i = 0; i < N; i++) A[i * N + i] += 1.0;
#pragma unroll 1
    for (int j = 0; j < N; j++) {
      double best = fabs(A[j * N + j]);
      int bi = j;
      for (int i = j + 1; i < N; i++) {
        const double v = fabs(A[i * N + j]);
        if (v > best) { best = v; bi = i; }
      }
      piv[j] = bi;
      if (bi != j) {
        for (int cc = 0; cc < N; cc++) {
          const double t_ = A[j * N + cc];
          A[j * N + cc] = A[bi * N + cc];
          A[bi * N + cc] = t_;
        }
      }
      double pv = A[j * N + j];
      if (pv == 0.) pv = 1e-50;
      const double pinv = 1.0 / pv;
      A[j * N + j] = pinv;
      for (int i = j + 1; i < N; i++) {
        const double fm = A[i * N + j] * pinv;
        A[i * N + j] = fm;
        for (int cc = j + 1; cc < N; cc++) A[i * N + cc] -= fm * A[j * N + cc];
      }
    }
    M.st.factorizations++;
  };
  auto solve = [&](double (&b)[N]) {  // in place
#pragma unroll
    for (int i = 0; i < N; i++) tmp[i] = b[i];
#pragma unroll 1
    for (int j = 0; j < N; j++) {
      const int p = piv[j];
      if (p != j) { const double t_ = tmp[j]; tmp[j] = tmp[p]; tmp[p] = t_; }
    }
#pragma unroll
    for (int i = 0; i < N; i++) b[i] = tmp[i];
#pragma unroll
    for (int i = 1; i < N; i++) {
      double s = b[i];
#pragma unroll
      for (int j = 0; j < i; j++) s -= A[i * N + j] * b[j];
      b[i] = s;
    }
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
      double s = b[i];
#pragma unroll
      for (int j = i + 1; j < N; j++) s -= A[i * N + j] * b[j];
      b[i] = s * A[i * N + i];
    }
    M.st.solves++;
  };
  auto rescale = [&](double r, int kord) {  // adjust_stepsize
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
#pragma unroll
    for (int i = 0; i < N; i++) {
      double row[5], o[5];
#pragma unroll
      for (int kk = 0; kk < 5; kk++) row[kk] = (kk < kord) ? dif[kk][i] : 0.;
#pragma unroll
      for (int jj = 0; jj < 5; jj++) {
        double s = 0.0;
#pragma unroll
        for (int kk = 0; kk < 5; kk++) s += row[kk] * RU[kk][jj];
        o[jj] = s;
      }
#pragma unroll
      for (int jj = 0; jj < 5; jj++)
        if (jj < kord) dif[jj][i] = o[jj];
    }
  };

  const double htspan = fabs(tfinal - t0);
  double t = t0, tnew = t0;
  ln_env_tail<NN>(P, M, t0, E);
  ln_rhs_tail<NN>(P, E, k2, ik2, y, f);
  M.st.fevals++;
  const double hmax = (tfinal - t0) / 10.0;
  jacobian();
  bool Jcurrent = true;
  double hmin = 16.0 * eps * fabs(t);
  double rh = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const double wt = fmax(fabs(y[i]), threshold);
    rh = fmax(rh, 1.25 / sqrt(rtol) * fabs(f[i] / wt));
  }
  double absh = fmin(hmax, htspan);
  if (absh * rh > 1.0) absh = 1.0 / rh;
  absh = fmax(absh, hmin);
  double h = absh;
  {
    double jf[N], fdel[N];
    ln_rhs_tail<NN>(P, E, k2, ik2, f, jf);  // J*f0 = f(t0, f0)
    const double tdel = (t + fmin(sqrt(eps) * fmax(fabs(t), fabs(t + h)), absh)) - t;
    ln_env_tail<NN>(P, M, t + tdel, E);
    ln_rhs_tail<NN>(P, E, k2, ik2, y, fdel);
    M.st.fevals += 2;
    rh = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      const double wt = fmax(fabs(y[i]), threshold);
      const double s = jf[i] + (fdel[i] - f[i]) / tdel;
      rh = fmax(rh, 1.25 * sqrt(0.5 * fabs(s / wt) / rtol));
    }
    absh = fmin(hmax, htspan);
    if (absh * rh > 1.0) absh = 1.0 / rh;
    absh = fmax(absh, hmin);
    h = absh;
  }
  int k = 1, klast = k;
  double abshlast = absh;
#pragma unroll
  for (int i = 0; i < N; i++) dif[0][i] = h * f[i];
  double hinvGak = h * c_invGa[k - 1];
  int nconhk = 0;
  factor(hinvGak);
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
        rescale(absh / abshlast, k);
        hinvGak = h * c_invGa[k - 1];
        nconhk = 0;
        factor(hinvGak);
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
    LN_TAIL_BY_ORDER(k, (ln_tail_predict<N, K_>(y, dif, psi, pred)));
#pragma unroll
    for (int i = 0; i < N; i++) {
      dk1[i] = 0.;
      iw[i] = 1.0 / fmax(fmax(fabs(pred[i]), fabs(y[i])), threshold);
      minnrm = fmax(minnrm, 100 * eps * fabs(pred[i] * iw[i]));
    }
    ln_env_tail<NN>(P, M, tnew, E);
    bool gotynew = false;
#pragma unroll 1
    for (int iter = 1; iter <= maxit; iter++) {
      double yn[N], r[N];
#pragma unroll
      for (int i = 0; i < N; i++) yn[i] = pred[i] + dk1[i];
      ln_rhs_tail<NN>(P, E, k2, ik2, yn, f);
      M.st.fevals++;
#pragma unroll
      for (int i = 0; i < N; i++) r[i] = hinvGak * f[i] - (psi[i] + dk1[i]);
      solve(r);
      double newnrm = 0.0;
#pragma unroll
      for (int i = 0; i < N; i++) {
        newnrm = fmax(newnrm, fabs(r[i] * iw[i]));
        dk1[i] += r[i];
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
        break;
      } else {
        rate = fmax(0.9 * rate, newnrm / oldnrm);
        havrate = true;
        const double errit = newnrm * rate / (1.0 - rate);
        if (errit <= 0.5 * rtol) { gotynew = true; break; }
        else if (iter == maxit) break;
        else {
          double rp = rate;
          for (int q = 1; q < maxit - iter; q++) rp *= rate;
          if (0.5 * rtol < errit * rp) break;
        }
      }
      oldnrm = newnrm;
    }
    if (!gotynew) {
      M.st.failed++;
      if (!Jcurrent) {
        ln_env_tail<NN>(P, M, t, E);
        M.st.fevals++;
        jacobian();
        Jcurrent = true;
      } else if (absh <= hmin) {
        M.status = 2;
        return false;
      } else {
        abshlast = absh;
        absh = fmax(0.3 * absh, hmin);
        h = absh;
        done = false;
        rescale(absh / abshlast, k);
        hinvGak = h * c_invGa[k - 1];
        nconhk = 0;
      }
      factor(hinvGak);
      havrate = false;
      continue;
    }
    err = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) err = fmax(err, fabs(dk1[i] * iw[i]));
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
          LN_TAIL_BY_ORDER(k, (errkm1 = ln_tail_errkm1<N, (K_ > 1 ? K_ : 2)>(dif, dk1, iw)));
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
      rescale(absh / abshlast, k);
      hinvGak = h * c_invGa[k - 1];
      nconhk = 0;
      factor(hinvGak);
      havrate = false;
      continue;
    }
    // ---- step accepted: dif[k+1] = dk1 - dif[k]; dif[k] = dk1; dif[j] += dif[j+1] (j < k)
    M.st.steps++;
    double e_km1 = 0., e_kp1 = 0.;
    LN_TAIL_BY_ORDER(k, (ln_tail_accept<N, K_>(dif, dk1, iw, e_km1, e_kp1)));
    // ---- output at the sample times passed by this step (through the generic source routine)
    while ((next < tres) && ((tnew - tnext) >= 0.0)) {
      double* yo = LVP(LV_TMP);
      double* dyo = LVP(LV_YPI);
      if (tnew == tnext) {
#pragma unroll
        for (int i = 0; i < N; i++) { yo[i] = pred[i] + dk1[i]; dyo[i] = f[i]; }
      } else {
        const double s = (tnext - tnew) / h;
        double c1[5], c2[5];
        double prod = 1.0, sumfrac = 0., fact = 1.0;
#pragma unroll
        for (int j = 0; j < 5; j++) {
          prod *= (s + j);
          fact *= (j + 1);
          sumfrac += 1.0 / (s + j);
          c1[j] = (j < k) ? prod / fact : 0.;
          c2[j] = (j < k) ? prod * sumfrac / (h * fact) : 0.;
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
          double a1 = 0, a2 = 0;
#pragma unroll
          for (int j = 0; j < 5; j++) { a1 += c1[j] * dif[j][i]; a2 += c2[j] * dif[j][i]; }
          yo[i] = (pred[i] + dk1[i]) + a1;
          dyo[i] = a2;
        }
      }
      ln_write_sources(P, M, mem, tnext, yo, dyo, next);
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
    t = tnew;
#pragma unroll
    for (int i = 0; i < N; i++) y[i] = pred[i] + dk1[i];
    Jcurrent = false;
    new_step = true;
  }
  {
    double* Ys = LVP(LV_Y);
#pragma unroll
    for (int i = 0; i < N; i++) Ys[i] = pred[i] + dk1[i];
  }
  M.st.fevals++;  // the reference's final RHS call (nothing follows the last interval)
  M.next = next;
  return true;
}

// dispatch on the number of ncdm species; false: not a tail interval (generic integrator)
LN_FN bool ln_is_tail(const PtParams& P, const Approx& ap) {
  return ap.rsa_on && (!P.has_ncdm || ap.ncdmfa_on) && P.N_ncdm <= PT_MAX_NCDM && P.rsa_method != CLPP_RSA_NONE;
}
LN_FN bool ln_run_tail(const PtParams& P, Lane& M, double* __restrict__ mem, double t0, double tfinal) {
  switch (P.has_ncdm ? P.N_ncdm : 0) {
    case 0: return ln_ndf15_tail<0>(P, M, mem, t0, tfinal);
    case 1: return ln_ndf15_tail<1>(P, M, mem, t0, tfinal);
    case 2: return ln_ndf15_tail<2>(P, M, mem, t0, tfinal);
    default: return ln_ndf15_tail<3>(P, M, mem, t0, tfinal);
  }
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
  for (int i = 0; i < L.neq; i++) LVP(LV_Y)[i] = 0.;
  LVP(LV_Y)[L.delta_g] = delta_g;
  LVP(LV_Y)[L.theta_g] = theta_g;
  LVP(LV_Y)[L.delta_b] = 3. / 4. * delta_g;
  LVP(LV_Y)[L.theta_b] = theta_g;
  LVP(LV_Y)[L.delta_cdm] = 3. / 4. * delta_g;
  LVP(LV_Y)[L.eta] = eta;
  if (P.has_ur) {
    LVP(LV_Y)[L.delta_ur] = delta_ur;
    LVP(LV_Y)[L.theta_ur] = theta_ur;
    LVP(LV_Y)[L.shear_ur] = shear_ur;
    LVP(LV_Y)[L.c_ur] = l3_ur;
  }
  if (P.has_ncdm) {
    for (int s = 0; s < P.N_ncdm; s++) {
      const double Ms = M.C->ncdm_M[s];
      for (int j = P.ncdm_q_off[s]; j < P.ncdm_q_off[s] + P.ncdm_q_size[s]; j++) {
        const int idx = L.psi0_ncdm1 + 3 * j;
        const double q = M.C->ncdm_q[j];
        const double dlnf0 = M.C->ncdm_dlnf0[j];
        const double epsq = sqrt(q * q + a * a * Ms * Ms);
        LVP(LV_Y)[idx + 0] = -0.25 * delta_ur * dlnf0;
        LVP(LV_Y)[idx + 1] = -epsq / 3. / q / k * theta_ur * dlnf0;
        LVP(LV_Y)[idx + 2] = -0.5 * shear_ur * dlnf0;
        LVP(LV_Y)[L.c_ncdm + j * L.len_ncdm] = -0.25 * l3_ur * dlnf0;
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
#define YO(i) LVP(LV_Y)[i]
#define YN(i) LVP(LV_YNEW)[i]
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
    const bool ok = (!P.force_generic && ln_is_tail(P, apn)) ? ln_run_tail(P, M, mem, limit[iv], limit[iv + 1])
                                                             : ln_ndf15(P, M, mem, limit[iv], limit[iv + 1]);
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
