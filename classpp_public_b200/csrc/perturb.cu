// Stage 1: per-wavenumber Einstein-Boltzmann integration on the device.
//
// One WARP integrates one (cosmology, k) mode from its initial time to today.  A CTA is a cohort
// of 1..8 such warps working on similar modes in lockstep (see PT_COHORT_SYNC below), and the
// hardware block scheduler is the work queue; modes of all cosmologies of a batch
// are issued in decreasing-cost order (longest chains first).  Everything of a mode lives in
// shared memory (~150 B per equation + the hub inverse): the state vector, the NDF
// backward-difference array and the factors of the Newton matrix (I - h/(G(1-alpha)) J).
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
//   interp_from_dif :860-905, new_linearisation :945-998.
//
// What is and what is not re-designed: the data layout, the warp mapping, the chain/hub linear algebra, the Jacobian
// probing, the tail kernel, the cohorts and the launch structure are this repository's own.  The CONTROL SKELETON of the
// three ndf15* integrators (step-size and order selection, Newton convergence test, failure handling: the branches and
// the constants 0.3 / 0.833 / 0.769 / 1.2 / 1.3 / 1.4 / 0.05 rtol ...) follows tools/evolver_ndf15.cpp:302-651 branch by
// branch and keeps its variable names (absh, abshlast, hinvGak, nconhk, havrate, gotynew ...): a transliteration by
// necessity, not a re-derivation -- the step decisions must match the reference's for the work counters and the 1e-4
// parity to hold (the reference itself is a port of MATLAB's ode15s).  The right-hand side is the Ma-Bertschinger algebra
// of perturb_derivs_member re-ordered for lanes.
//
// Differences of METHOD (not of result) from the reference:
//  * Linear algebra.  The reference discovers the sparsity of J numerically and runs a generic
//    sparse LU (tools/sparse.c).  Here the structure of the Boltzmann hierarchy is used directly:
//    every multipole l >= 3 of a hierarchy (photon temperature, photon polarisation, ur, each
//    ncdm momentum bin) couples only to l-1 and l+1, so the state splits into CHAINS
//    (tridiagonal, rooted at their l = 2 moment) and a small dense HUB block (everything else:
//    densities, velocities, shears, eta; 7..28 variables).  (I - cJ) is factorised exactly as
//    chains -> Schur complement on the hub -> explicit hub inverse (Gauss-Jordan with partial
//    pivoting); the chain eliminations need no pivoting (all pivots >= 1).
//    A solve is two chain sweeps (one lane per chain) + one hub mat-vec (one lane per row).
//  * Jacobian.  The system is linear and homogeneous in y, so J's columns are f(t, e_j) exactly;
//    hub columns are probed one by one and all chain entries with 3 grouped probes
//    (l mod 3), instead of the reference's finite-difference numjac (:1213-1539).
//  * Background/thermodynamics lookups keep the current table interval in registers (one
//    column per lane) and only touch memory when the interval changes.
#include <algorithm>
#include <cmath>
#include <vector>

#include "device.h"

#include "pt_types.h"

extern __shared__ double smem_all[];
// A CTA is a COHORT of P.wpc warps (1..PT_MAX_WPC), one k mode per warp, each with its own shared-memory region.
// The warps of a cohort integrate similar modes (adjacent in the cost-sorted issue order: the same k of neighbouring
// cosmologies of a batch) and meet at a CTA barrier at the top of every step, so that they walk through the same code at
// the same time: the per-step code footprint (~70 KB) exceeds the 32 KB instruction cache, and instruction fetch, not
// issue slots, is what saturates an SM (profiles/r01_perturb_kernel_loaded_v3_footprint.txt).
// The barrier is bar.sync 0 over the whole CTA, reached from different call sites (three integrators) by warps that are
// each fully converged; a warp that finishes its mode (or hands it to the tail kernel) simply exits, which the hardware
// counts as arrived, so cohorts may be ragged (last CTA of a launch, modes of different length).
#ifndef PT_MAX_WPC
#define PT_MAX_WPC 8
#endif
#ifndef PT_GEN_MIN_BLOCKS
#define PT_GEN_MIN_BLOCKS 1  // 255 registers; build variants (-DPT_MAX_WPC=4 -DPT_GEN_MIN_BLOCKS=3) trade spills for occupancy
#endif
#define PT_LANE ((int)(threadIdx.x & 31))
#define PT_WARP ((int)(threadIdx.x >> 5))
#define PT_SLOT(P) ((int)(blockIdx.x * (P).wpc) + PT_WARP)
#define SMEM(P) (smem_all + (size_t)PT_WARP * (P).wstride)
// `sync_ctr` is a local counter of the calling integrator: with sync_every = n a warp joins the barrier on every n-th of its
// step attempts (the j-th barrier of one warp pairs with the j-th of the others; exited warps count as arrived)
#define PT_COHORT_SYNC(P) do { if ((P).wpc > 1 && ++sync_ctr >= (P).sync_every) { sync_ctr = 0; asm volatile("bar.sync 0;" ::: "memory"); } } while (0)
// every shared-memory access goes through these, so that the compiler sees the shared address
// space (LDS/STS with 32-bit addresses) instead of generic pointers
#define s_pvb(P) (SMEM(P))
#define s_pvt(P) (SMEM(P) + 32)
#define s_hubtmp(P) (SMEM(P) + (P).o_hubtmp)
#define s_nw(P) (SMEM(P) + (P).o_nw)
#define s_vec(P, slot) (SMEM(P) + ((P).o_vec + (slot) * (P).np))
#define s_sinv(P) (SMEM(P) + (P).o_sinv)
// bracketing rows of table `tab` (0 background, 1 thermodynamics), cache set `set`: [4][ncol] = y0, y1, dd0, dd1
#define s_tabc(P, tab, set) (SMEM(P) + (P).o_tabc + ((tab) * 4) * (P).ncol)
#define s_i2l1(P) (SMEM(P) + (P).o_i2l1)
#define s_hub_idx(P) ((int*)(SMEM(P) + (P).o_int))
#define s_piv(P) ((int*)(SMEM(P) + (P).o_int) + (P).nh_max)
#define s_ch_start(P) ((int*)(SMEM(P) + (P).o_int) + 2 * (P).nh_max)
#define s_ch_len(P) ((int*)(SMEM(P) + (P).o_int) + 2 * (P).nh_max + PT_MAX_CHAINS)
#define s_ch_rootslot(P) ((int*)(SMEM(P) + (P).o_int) + 2 * (P).nh_max + 2 * PT_MAX_CHAINS)

// NDF constants (evolver_ndf15.cpp:86-100), indexed by order-1
__constant__ double c_G[5] = {1.0, 3.0 / 2.0, 11.0 / 6.0, 25.0 / 12.0, 137.0 / 60.0};
__constant__ double c_invGa[5] = {1.0 / (1.0 * (1.0 + 37.0 / 200)), 1.0 / (1.5 * (1.0 + 1.0 / 9.0)),
                                  1.0 / (11.0 / 6.0 * (1.0 + 8.23e-2)), 1.0 / (25.0 / 12.0 * (1.0 + 4.15e-2)),
                                  1.0 / (137.0 / 60.0)};
__constant__ double c_erconst[5] = {-37.0 / 200 * 1.0 + 1.0 / 2.0, -1.0 / 9.0 * 1.5 + 1.0 / 3.0,
                                    -8.23e-2 * (11.0 / 6.0) + 1.0 / 4.0, -4.15e-2 * (25.0 / 12.0) + 1.0 / 5.0,
                                    1.0 / 6.0};
// difference-array rescaling matrix U (evolver_ndf15.cpp:907-943)
// 1/m for the R(r) recurrence of adjust_stepsize
__constant__ double c_invint[7] = {0., 1.0, 0.5, 1.0 / 3.0, 0.25, 0.2, 1.0 / 6.0};
__constant__ double c_U[5][5] = {{-1, -2, -3, -4, -5}, {0, 1, 3, 6, 10}, {0, 0, -1, -4, -10}, {0, 0, 0, 1, 5}, {0, 0, 0, 0, -1}};


struct Layout {
  int neq;
  int delta_g, theta_g, shear_g, l3_g, pol0_g;  // -1 when absent
  int delta_b, theta_b, delta_cdm;
  int delta_ur, theta_ur, shear_ur, l3_ur;
  int psi0_ncdm1;
  int eta;
  int l_max_ncdm, q_size_ncdm[PT_MAX_NCDM], ncdm_off[PT_MAX_NCDM];
};

// background + thermodynamics at the current time; identical in all lanes
enum { DRV_INV_R = 0, DRV_R, DRV_INV_1PR, DRV_INV_HALF_AH, DRV_INV_TAU, DRV_TAU_C, DRV_FAC_NCDM, DRV_COUNT = 8 };
struct Env {
  double tau, a, H, Hp;
  // hoisted out of the RHS: R = 4 rho_g / 3 rho_b, reciprocals, (a_today/a)^4 (see env_at)
  double drv[DRV_COUNT];
};
// interval cache of a table: bracketing rows (y0, y1, dd0, dd1), one column per lane
struct TabCache {
  double x0, x1, ih, h26;
  int cur, pad;  // current interval
};


struct Metric {
  double h_prime, eta_prime, alpha, alpha_prime;
  double rsa_delta_g, rsa_theta_g;
  double delta_m, delta_cb;
  double tca_shear_g;
};

// vector slots in shared memory (each np doubles)
enum {
  V_Y = 0, V_YNEW, V_F0, V_PRED, V_PSI, V_DIFKP1, V_DEL, V_INVWT,
  V_TMP = V_PSI, V_YPI = V_PRED,  // scratch of the Jacobian probes / source output: psi and pred are dead there
  V_DIF0 = V_INVWT + 1,  // 5 slots: dif[0..4]; dif[5], dif[6] (orders 4 and 5 only) live in the mode's global scratch
  V_JL = V_DIF0 + 5,     // chain rows of J: sub-diagonal J[i,i-1] (T[i,i-1] = -c J[i,i-1] is formed on the fly by solve);
                         // the diagonal and the super-diagonal are only read by factor(): global scratch (Mode::gJd, gJu)
  V_IP, V_MU,            // chain factors: 1/pivot, T[i,i+1]/p[i+1]
  V_COUNT
};
// Shared memory per mode decides how many cohorts fit on an SM: 16 slots (17.4 KB for 136 equations) instead of 21 puts
// two 4-mode cohorts of the largest Planck-18 system (136 equations, 28 hub variables) on one SM.

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double wmax_any(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(PT_FULL, v, o));
  return v;
}
// warp maximum of NON-NEGATIVE doubles (error norms): their bit patterns order like unsigned integers, so two
// hardware integer reductions (REDUX) replace five double shuffle+max rounds
__device__ __forceinline__ double wmax(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mhi = __reduce_max_sync(PT_FULL, hi);
  const unsigned mlo = __reduce_max_sync(PT_FULL, hi == mhi ? lo : 0u);
  return __hiloint2double((int)mhi, (int)mlo);
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PT_FULL, v, o);
  return v;
}
// x^(1/n), x > 0 (step-size heuristics; the reference calls pow)
// (the operands are ratios of error norms to the tolerance: float range and float accuracy are ample for a
// step-size heuristic; the reference calls pow())
__device__ __forceinline__ double root_n(double x, double n) { return (double)exp2f(__log2f((float)x) / (float)n); }

// largest index inf <= n-2 with X[inf] <= x (X growing): 32-ary search, one probe per lane
__device__ __forceinline__ int locate_warp(const double* __restrict__ X, int n, double x, int lane) {
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int step = (hi - lo + 31) >> 5;
    const int idx = lo + (lane + 1) * step;
    const bool ge = (idx < hi) && (x >= __ldg(X + idx));
    const int c = __popc(__ballot_sync(PT_FULL, ge));
    hi = min(lo + (c + 1) * step, hi);
    lo = lo + c * step;
  }
  return lo;
}
// cursor walk (array_interpolate_spline_growing_closeby, arrays.c:2173-2232), falling back to the
// warp search when the target is far
__device__ __forceinline__ int locate_closeby(const double* __restrict__ X, int n, double x, int cursor, int lane) {
  int inf = min(max(cursor, 0), n - 2);
  int guard = 0;
  while (inf > 0 && x < __ldg(X + inf)) {
    inf--;
    if (++guard > 6) return locate_warp(X, n, x, lane);
  }
  int sup = inf + 1;
  while (sup < n - 1 && x > __ldg(X + sup)) {
    sup++;
    if (++guard > 6) return locate_warp(X, n, x, lane);
  }
  return sup - 1;
}

// optional section profiler (cycles per code section, summed over the mode's lifetime)
#ifdef PT_PROF
#define PROF_DECL long long prof_t0_
#define PROF_BEGIN() (prof_t0_ = clock64())
#define PROF_END(slot) do { if (PT_LANE == 0) M.prof[slot] += clock64() - prof_t0_; } while (0)
#else
#define PROF_DECL
#define PROF_BEGIN()
#define PROF_END(slot)
#endif
enum { PF_ENV = 0, PF_RHS, PF_SOLVE, PF_PREDICT, PF_UPDATE, PF_CONTROL, PF_FACTOR, PF_ADJUST, PF_OUTPUT, PF_JAC, PF_DIFUPD, PF_X, PF_COUNT };

// per-mode statistics (copied to the public clpp_kstat at the end)
struct Stat {
  int steps, failed, fevals, jacobians, factorizations, solves, intervals, status;
  double tau_ini;
};

// State of the mode a warp integrates.  It lives in SHARED memory (one per warp of the CTA): every
// field is warp-uniform except the table caches, which hold one column per lane.  Keeping it out
// of local memory matters: with several warps per SM the per-thread stack frames thrash the L1.
struct Mode {
  const PtCosmo* C;
  double* Jhh;  // global scratch [nh][nh], column-major: Jhh[i + j*nh] = J[hub i, hub j]
  double *gJd, *gJu, *gdif;  // global scratch [np] each: chain diagonal / super-diagonal of J; dif[5], dif[6] ([2][np])
  double fac_c;              // c of the current factorisation of I - c J
  double k, k2, inv_k, inv_k2;
  double nf[PT_MAX_NCDM][8];  // ncdm fluid constants at the current time (env_at)
  TabCache bgc[2], thc[2];  // [0] time stepping, [1] source output
  // per-cosmology scalars and table pointers (copied from PtCosmo once per mode)
  const double *bg_tau, *bg_y, *bg_dd, *th_z, *th_y, *th_dd;
  double z_last, th_lin, a_today;
  double q[4 * PT_MAX_NCDM];
  int bt_size, tt_size;
  Env e;
  Metric m;
  double tca_shear_last;  // photon shear of the last Newton-iteration RHS call (used by the sources while TCA is on)
  double limit[PT_MAX_INTERVALS + 1];
  double sw[4];
  long long prof[PF_COUNT];
  Approx sched[PT_MAX_INTERVALS];
  Approx ap, apprev;
  Layout L, Lprev;
  Stat st;
  int ik, need_nw, nh, nch, status, next;
};
#define MODE(P) (*(Mode*)(SMEM(P) + (P).o_mode))



// ---------------------------------------------------------------------------------------------
// background_at_tau (normal_info columns) + thermodynamics_at_z, cooperative over lanes.
// `set` used to select one of two interval caches (0 = the time stepping, 1 = the source output); there is one
// cache now (shared memory per mode decides how many cohorts fit on an SM) and the argument is ignored (both
// sweep the tables monotonically, at different times).  Entering a new interval also prefetches
// the row the sweep will need next into L1.  Everything that only depends on time and that the RHS
// would otherwise divide by is computed here ONCE per step, the divisions side by side in lanes.
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __noinline__ void env_at(const PtParams& P, double tau, bool closeby, int set) {
  Mode& M = MODE(P);
  set = 0;  // one interval cache: the output times lie inside the current step, i.e. (almost always) in the cached interval
  const int lane = PT_LANE;
  double* pvb = s_pvb(P);
  double* pvt = s_pvt(P);
  PROF_DECL;
  PROF_BEGIN();
  // background
  {
    TabCache& T = M.bgc[set];
    double* tc = s_tabc(P, 0, set);
    const int nc = P.ncol;
    if (!(tau >= T.x0 && tau <= T.x1)) {
      __syncwarp();
      const double* X = M.bg_tau;
      const int n = M.bt_size;
      const int inf = closeby ? locate_closeby(X, n, tau, T.cur, lane) : locate_warp(X, n, tau, lane);
      const double x0 = __ldg(X + inf), x1 = __ldg(X + inf + 1);
      if (lane < P.bg_size_normal) {
        const size_t r0 = (size_t)inf * P.bg_size + lane, r1 = r0 + P.bg_size;
        tc[lane] = __ldg(M.bg_y + r0); tc[nc + lane] = __ldg(M.bg_y + r1);
        tc[2 * nc + lane] = __ldg(M.bg_dd + r0); tc[3 * nc + lane] = __ldg(M.bg_dd + r1);
        if (inf + 3 < n) {
          prefetch_l1(M.bg_y + r1 + P.bg_size); prefetch_l1(M.bg_dd + r1 + P.bg_size);
          prefetch_l1(M.bg_y + r1 + 2 * P.bg_size); prefetch_l1(M.bg_dd + r1 + 2 * P.bg_size);
        }
      }
      if (lane == 0) {
        T.cur = inf; T.x0 = x0; T.x1 = x1;
        const double h = x1 - x0;
        T.ih = 1.0 / h; T.h26 = h * h / 6.;
      }
      __syncwarp();
    }
    const double b = (tau - T.x0) * T.ih, a = 1 - b;
    if (lane < P.bg_size_normal)
      pvb[lane] = a * tc[lane] + b * tc[nc + lane] + ((a * a * a - a) * tc[2 * nc + lane] + (b * b * b - b) * tc[3 * nc + lane]) * T.h26;
  }
  __syncwarp();
  PROF_END(PF_JAC);
  PROF_BEGIN();
  const double av = pvb[P.ia], Hv = pvb[P.iH], Hp = pvb[P.iHp];
  const double inv_a = 1. / av;
  const double z = inv_a - 1.;
  // thermodynamics
  const double z_last = M.z_last;
  if (z >= z_last) {
    if (lane == 0) {
      const PtCosmo* C = M.C;
      const double* row = M.th_y + (size_t)(M.tt_size - 1) * P.th_size;
      const double x0 = row[P.ixe];
      double* pv = pvt;
      const double r = (1. + z) / (1. + z_last);
      pv[P.ixe] = x0;
      pv[P.idkappa] = (1. + z) * (1. + z) * C->n_e * x0 * CLPP_sigma * CLPP_Mpc_over_m;
      pv[P.itau_d] = row[P.itau_d] * r * r;
      if (P.compute_damping_scale) pv[P.ir_d] = row[P.ir_d] / (r * sqrt(r));
      pv[P.iddkappa] = -Hv * 2. / (1. + z) * pv[P.idkappa];
      pv[P.idddkappa] = (Hv * Hv / (1. + z) - Hp) * 2. / (1. + z) * pv[P.idkappa];
      pv[P.iexp_m_kappa] = 0.;
      pv[P.ig] = 0.;
      pv[P.idg] = 0.;
      pv[P.iddg] = 0.;
      pv[P.iTb] = C->T_cmb * (1. + z);
      pv[P.iwb] = CLPP_k_B / (CLPP_c * CLPP_c * CLPP_m_H) * (1. + (1. / CLPP_not4 - 1.) * C->YHe + x0 * (1. - C->YHe)) *
                  C->T_cmb * (1. + z);
      pv[P.icb2] = pv[P.iwb] * 4. / 3.;
      if (P.compute_cb2_derivatives) {
        pv[P.idcb2] = -Hv * av * pv[P.icb2];
        pv[P.iddcb2] = -Hp * av * pv[P.icb2];
      }
      pv[P.irate] = pv[P.idkappa];
    }
  } else {
    TabCache& T = M.thc[set];
    double* tc = s_tabc(P, 1, set);
    const int nc = P.ncol;
    const bool linear = (z < M.th_lin);
    if (!(z >= T.x0 && z <= T.x1)) {
      __syncwarp();
      const double* X = M.th_z;
      const int n = M.tt_size;
      const int inf = (closeby && !linear) ? locate_closeby(X, n, z, T.cur, lane) : locate_warp(X, n, z, lane);
      const double x0 = __ldg(X + inf), x1 = __ldg(X + inf + 1);
      if (lane < P.th_size) {
        const size_t r0 = (size_t)inf * P.th_size + lane, r1 = r0 + P.th_size;
        tc[lane] = __ldg(M.th_y + r0); tc[nc + lane] = __ldg(M.th_y + r1);
        tc[2 * nc + lane] = __ldg(M.th_dd + r0); tc[3 * nc + lane] = __ldg(M.th_dd + r1);
        if (inf > 1) {  // z decreases with time
          prefetch_l1(M.th_y + r0 - P.th_size); prefetch_l1(M.th_dd + r0 - P.th_size);
          prefetch_l1(M.th_y + r0 - 2 * P.th_size); prefetch_l1(M.th_dd + r0 - 2 * P.th_size);
        }
      }
      if (lane == 0) {
        T.cur = inf; T.x0 = x0; T.x1 = x1;
        const double h = x1 - x0;
        T.ih = 1.0 / h; T.h26 = h * h / 6.;
      }
      __syncwarp();
    }
    const double b = (z - T.x0) * T.ih, a = 1 - b;
    if (lane < P.th_size) {
      double v = a * tc[lane] + b * tc[nc + lane];
      if (!linear) v += ((a * a * a - a) * tc[2 * nc + lane] + (b * b * b - b) * tc[3 * nc + lane]) * T.h26;
      pvt[lane] = v;
    }
  }
  PROF_END(PF_X);
  // momentum-dependent ncdm weights at this scale factor (used by the RHS while the ncdm hierarchy is integrated)
  if (M.need_nw) {
    const double a2 = av * av;
    const int nq = P.nq_tot;
    double* nw = s_nw(P);
#pragma unroll 1
    for (int j = lane; j < nq; j += 32) {
      int s = 0;
#pragma unroll
      for (int t = 1; t < PT_MAX_NCDM; t++)
        if (t < P.N_ncdm && j >= P.ncdm_q_off[t]) s = t;
      const double Ms = M.C->ncdm_M[s];
      const double q = __ldg(M.C->ncdm_q + j), w0 = __ldg(M.C->ncdm_w + j);
      const double q2 = q * q, eps = sqrt(q2 + Ms * Ms * a2);
      const double ieps = 1.0 / eps;
      nw[j] = q * ieps;
      nw[nq + j] = q2 * eps * w0;
      nw[2 * nq + j] = q2 * q * w0;
      nw[3 * nq + j] = q2 * q2 * ieps * w0;
    }
  }
  __syncwarp();
  // derived quantities: ONE SIMD division; lane j < 6 produces DRV j, lanes 8+4s.. the ratios of ncdm
  // species s needed by the fluid approximation (w = p/rho, pseudo_p/p, 1/(1+w), 1/w).  Selects only,
  // no lane-divergent branches.
  {
    const double rho_g = pvb[P.irho_g], rho_b = pvb[P.irho_b], dkappa = pvt[P.idkappa];
    const double aH = Hv * av;
    double num = 1., den = 1.;
    num = (lane == DRV_INV_R) ? 0.75 * rho_b : num;          den = (lane == DRV_INV_R) ? rho_g : den;
    num = (lane == DRV_R) ? 4. / 3. * rho_g : num;           den = (lane == DRV_R) ? rho_b : den;
    num = (lane == DRV_INV_1PR) ? rho_b : num;               den = (lane == DRV_INV_1PR) ? rho_b + 4. / 3. * rho_g : den;
    num = (lane == DRV_INV_HALF_AH) ? 2. : num;              den = (lane == DRV_INV_HALF_AH) ? aH : den;
    den = (lane == DRV_INV_TAU) ? tau : den;
    den = (lane == DRV_TAU_C) ? dkappa : den;
    const bool fluid = P.has_ncdm && M.ap.ncdmfa_on;
    if (fluid) {
      const int s = min(max((lane - 8) >> 2, 0), P.N_ncdm - 1), r = (lane - 8) & 3;
      const bool mine = lane >= 8 && lane < 8 + 4 * P.N_ncdm;
      const double rho_n = pvb[P.irho_ncdm1 + s], p_n = pvb[P.ip_ncdm1 + s], pseudo = pvb[P.ipseudo_p_ncdm1 + s];
      const double nn = r == 0 ? p_n : r == 1 ? pseudo : rho_n;
      const double dd = r == 0 ? rho_n : r == 1 ? p_n : r == 2 ? rho_n + p_n : p_n;
      num = mine ? nn : num;
      den = mine ? dd : den;
    }
    const double q = num / den;
    if (lane < 6) M.e.drv[lane] = q;
    else if (lane >= 8 && lane < 8 + 4 * PT_MAX_NCDM) M.q[lane - 8] = q;
    if (lane == 6) {
      const double a_rel = M.a_today * inv_a;
      M.e.drv[DRV_FAC_NCDM] = (a_rel * a_rel) * (a_rel * a_rel);
      M.e.tau = tau; M.e.a = av; M.e.H = Hv; M.e.Hp = Hp;
    }
    __syncwarp();
    // ncdm fluid constants (perturb_derivs_member :8800-8850, perturb_total_stress_energy :6380-6400);
    // lane j < 8 N_ncdm writes constant (j & 7) of species (j >> 3)
    if (fluid) {
      const int s = min(lane >> 3, P.N_ncdm - 1), c8 = lane & 7;
      const double rho_n = pvb[P.irho_ncdm1 + s], p_n = pvb[P.ip_ncdm1 + s];
      const double w_n = M.q[4 * s], pseudo_p_over_p = M.q[4 * s + 1], i1w = M.q[4 * s + 2], inv_w = M.q[4 * s + 3];
      const double cg2 = w_n * (1.0 - i1w * (1. / 3.) * (3.0 * w_n - 2.0 + pseudo_p_over_p));
      const double ca2 = w_n * (1. / 3.) * i1w * (5.0 - pseudo_p_over_p);
      const double cvis2 = (P.ncdmfa_method == CLPP_NCDMFA_HU) ? w_n : 3. * w_n * ca2;
      const double damp = (P.ncdmfa_method == CLPP_NCDMFA_HU) ? 3.0 * aH * ca2 * inv_w
                                                              : 3.0 * (aH * (2. / 3. - ca2 - pseudo_p_over_p * (1. / 3.)) + 1.0 / tau);
      double v = rho_n;                          // [0] rho
      v = c8 == 1 ? rho_n + p_n : v;             // [1] rho + p
      v = c8 == 2 ? w_n : v;                     // [2] w
      v = c8 == 3 ? cg2 * rho_n : v;             // [3] cg2 rho
      v = c8 == 4 ? ca2 : v;                     // [4] ca2 (= ceff2)
      v = c8 == 5 ? ca2 * i1w : v;               // [5] ceff2 / (1+w)
      v = c8 == 6 ? 8.0 / 3.0 * cvis2 * i1w : v; // [6] 8/3 cvis2 / (1+w)
      v = c8 == 7 ? damp : v;                    // [7] shear damping rate
      if (lane < 8 * P.N_ncdm) M.nf[s][c8] = v;
      __syncwarp();
    }
  }
}

// perturb_approximations: flags at time tau (uses bisection lookups, inter_normal)
__device__ __noinline__ Approx approximations_at(const PtParams& P, double tau) {
  Mode& M = MODE(P);
  env_at(P, tau, false, 0);
  const Env& e = M.e;
  Approx a;
  const double tau_k = 1. / M.k, tau_h = 1. / (e.H * e.a);
  const double dkappa = s_pvt(P)[P.idkappa];
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

__device__ __forceinline__ int approx_flag(const Approx& a, int which) {
  return which == 0 ? a.tca_off : which == 1 ? a.rsa_on : which == 2 ? a.ufa_on : a.ncdmfa_on;
}

// perturb_vector_init (index part): layout of the state vector for a set of approximations
__device__ __forceinline__ void make_layout(const PtParams& P, const Approx& ap, Layout& L) {
  int n = 0;
  L.delta_g = L.theta_g = L.shear_g = L.l3_g = L.pol0_g = -1;
  L.delta_ur = L.theta_ur = L.shear_ur = L.l3_ur = -1;
  L.psi0_ncdm1 = -1;
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
#pragma unroll
  for (int s = 0; s < PT_MAX_NCDM; s++) { L.q_size_ncdm[s] = 0; L.ncdm_off[s] = 0; }
  if (P.has_ncdm) {
    L.psi0_ncdm1 = n;
    L.l_max_ncdm = ap.ncdmfa_on ? 2 : P.l_max_ncdm;
#pragma unroll
    for (int s = 0; s < PT_MAX_NCDM; s++) {
      if (s < P.N_ncdm) {
        L.q_size_ncdm[s] = ap.ncdmfa_on ? 1 : P.ncdm_q_size[s];
        L.ncdm_off[s] = n;
        n += (L.l_max_ncdm + 1) * L.q_size_ncdm[s];
      }
    }
  }
  L.eta = n++;
  L.neq = n;
}

// Split the state of the current layout into chains (multipoles l >= 3 of each hierarchy, rooted at
// their l = 2 moment, which sits right before them in the state vector) and hub variables.
__device__ __noinline__ void make_structure(const PtParams& P) {
  Mode& M = MODE(P);
  const Layout& L = M.L;
  const Approx& ap = M.ap;
  if (PT_LANE == 0) {
    int nch = 0;
    if (L.l3_g >= 0) {
      s_ch_start(P)[nch] = L.l3_g; s_ch_len(P)[nch] = P.l_max_g - 2; nch++;
      s_ch_start(P)[nch] = L.pol0_g + 3; s_ch_len(P)[nch] = P.l_max_pol_g - 2; nch++;
    }
    if (L.l3_ur >= 0) { s_ch_start(P)[nch] = L.l3_ur; s_ch_len(P)[nch] = P.l_max_ur - 2; nch++; }
    if (P.has_ncdm && !ap.ncdmfa_on) {
      const int stride = L.l_max_ncdm + 1;
      const int nq = (L.eta - L.psi0_ncdm1) / stride;
      for (int iq = 0; iq < nq; iq++) {
        s_ch_start(P)[nch] = L.psi0_ncdm1 + iq * stride + 3; s_ch_len(P)[nch] = L.l_max_ncdm - 2; nch++;
      }
    }
    // hub variables: everything that is not a chain interior (chains are sorted by start)
    int nh = 0, c = 0;
    for (int i = 0; i < L.neq; i++) {
      while (c < nch && i >= s_ch_start(P)[c] + s_ch_len(P)[c]) c++;
      if (c < nch && i >= s_ch_start(P)[c]) continue;
      s_hub_idx(P)[nh++] = i;
    }
    // hub slot of each chain root
    for (int cc = 0, s = 0; cc < nch; cc++) {
      while (s_hub_idx(P)[s] != s_ch_start(P)[cc] - 1) s++;
      s_ch_rootslot(P)[cc] = s;
    }
    s_piv(P)[0] = nh;
    s_piv(P)[1] = nch;
  }
  __syncwarp();
  M.nh = s_piv(P)[0];
  M.nch = s_piv(P)[1];
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Right-hand side f(tau, y) for the environment currently stored in M.e/pvb/pvt (one out-of-line
// copy shared by the Newton loop, the Jacobian probes and the source output: the instruction
// cache is the scarce resource of this kernel).  Also fills M.m (metric and derived quantities).
// All lanes evaluate the scalar (hub) equations redundantly from broadcast shared-memory loads --
// straight-line code without lane-divergent sections -- and the multipole chains are evaluated one
// element per lane.  Every division by a quantity that only depends on time or on k is hoisted into
// env_at / the mode set-up (reciprocals in M.e, M.ik2).
__device__ __noinline__ void rhs_apply(const PtParams& P, int slot_y, int slot_dy, int want_matter) {
  Mode& M = MODE(P);
  const double* __restrict__ y = s_vec(P, slot_y);
  double* __restrict__ dy = slot_dy >= 0 ? s_vec(P, slot_dy) : nullptr;
  const Layout& L = M.L;
  const Approx ap = M.ap;
  const Env& e = M.e;
  const int lane = PT_LANE;
  const double k = M.k, k2 = M.k2, ik2 = M.inv_k2;
  const double* pvb = s_pvb(P);
  const double* pvt = s_pvt(P);
  const double a2 = e.a * e.a, aH = e.H * e.a, R = e.drv[DRV_R];
  const double rho_g = pvb[P.irho_g], rho_b = pvb[P.irho_b], rho_cdm = pvb[P.irho_cdm];
  const double rho_ur = P.has_ur ? pvb[P.irho_ur] : 0.;
  const double dkappa = pvt[P.idkappa], cb2 = pvt[P.icb2];
  Metric& m = M.m;
  const bool has_g = !ap.rsa_on;
  const bool has_ur = P.has_ur && !ap.rsa_on;
  const bool full_g = has_g && ap.tca_off;

  // ---- loads of every hub value (independent, issued up front)
  double delta_g = 0., theta_g = 0., shear_g = 0.;
  if (has_g) { delta_g = y[L.delta_g]; theta_g = y[L.theta_g]; }
  double g3 = 0., g4 = 0., p0 = 0., p1 = 0., p2 = 0., p3 = 0.;
  if (full_g) {
    shear_g = y[L.shear_g]; g3 = y[L.l3_g]; g4 = y[L.l3_g + 1];
    p0 = y[L.pol0_g]; p1 = y[L.pol0_g + 1]; p2 = y[L.pol0_g + 2]; p3 = y[L.pol0_g + 3];
  }
  double delta_ur = 0., theta_ur = 0., shear_ur = 0., u3 = 0., u4 = 0.;
  if (has_ur) {
    delta_ur = y[L.delta_ur]; theta_ur = y[L.theta_ur]; shear_ur = y[L.shear_ur];
    if (!ap.ufa_on) { u3 = y[L.l3_ur]; u4 = y[L.l3_ur + 1]; }
  }
  const double delta_b = y[L.delta_b], theta_b = y[L.theta_b], delta_cdm = y[L.delta_cdm], eta = y[L.eta];
  const double delta_p_b_over_rho_b = cb2 * delta_b;

  // ---- perturb_total_stress_energy
  double delta_rho = rho_g * delta_g + rho_b * delta_b;
  double rpt = 4. / 3. * rho_g * theta_g + rho_b * theta_b;
  double rps = 4. / 3. * rho_g * shear_g;
  double delta_p = 1. / 3. * rho_g * delta_g + rho_b * delta_p_b_over_rho_b;
  double delta_rho_m = rho_b * delta_b, rho_m = rho_b, rpt_m = rho_b * theta_b, rpm = rho_b;
  delta_rho += rho_cdm * delta_cdm;
  delta_rho_m += rho_cdm * delta_cdm; rho_m += rho_cdm; rpm += rho_cdm;
  if (P.has_ur) {
    delta_rho = delta_rho + rho_ur * delta_ur;
    rpt = rpt + 4. / 3. * rho_ur * theta_ur;
    rps = rps + 4. / 3. * rho_ur * shear_ur;
    delta_p += 1. / 3. * rho_ur * delta_ur;
  }
  if (want_matter) m.delta_cb = delta_rho_m / rho_m + 3. * aH * (rpt_m / rpm) * ik2;
  if (P.has_ncdm) {
    if (ap.ncdmfa_on) {
      for (int s = 0; s < P.N_ncdm; s++) {
        const double* nf = M.nf[s];  // rho, rho+p, w, cg2*rho, ...
        const int idx = L.psi0_ncdm1 + 3 * s;
        const double y0 = y[idx], y1 = y[idx + 1], y2 = y[idx + 2];
        delta_rho += nf[0] * y0;
        rpt += nf[1] * y1;
        rps += nf[1] * y2;
        delta_p += nf[3] * y0;
        delta_rho_m += nf[0] * y0; rho_m += nf[0];
        rpt_m += nf[1] * y1; rpm += nf[1];
      }
    } else {
      const int stride = L.l_max_ncdm + 1, nqt = P.nq_tot;
      const double* w = s_nw(P);
      for (int s = 0; s < P.N_ncdm; s++) {
        const double rho_n = pvb[P.irho_ncdm1 + s], p_n = pvb[P.ip_ncdm1 + s];
        const double factor = M.C->ncdm_factor[s] * e.drv[DRV_FAC_NCDM];
        double s_rho = 0., s_theta = 0., s_shear = 0., s_p = 0.;
        const int nq = P.ncdm_q_size[s], q0 = P.ncdm_q_off[s];
        const int base = L.psi0_ncdm1 + q0 * stride;
        if (nq <= 8) {  // few bins: every lane sums them all (no shuffles)
          for (int iq = 0; iq < nq; iq++) {
            const int idx = base + iq * stride;
            const double y0 = y[idx], y1 = y[idx + 1], y2 = y[idx + 2];
            s_rho += w[nqt + q0 + iq] * y0;
            s_theta += w[2 * nqt + q0 + iq] * y1;
            s_shear += w[3 * nqt + q0 + iq] * y2;
            s_p += w[3 * nqt + q0 + iq] * y0;
          }
        } else {
#pragma unroll 1
          for (int iq = lane; iq < nq; iq += 32) {
            const int idx = base + iq * stride;
            const double y0 = y[idx], y1 = y[idx + 1], y2 = y[idx + 2];
            s_rho += w[nqt + q0 + iq] * y0;
            s_theta += w[2 * nqt + q0 + iq] * y1;
            s_shear += w[3 * nqt + q0 + iq] * y2;
            s_p += w[3 * nqt + q0 + iq] * y0;
          }
          s_rho = wsum(s_rho); s_theta = wsum(s_theta); s_shear = wsum(s_shear); s_p = wsum(s_p);
        }
        s_rho *= factor;
        s_theta *= k * factor;
        s_shear *= 2.0 / 3.0 * factor;
        s_p *= factor / 3.;
        delta_rho += s_rho; rpt += s_theta; rps += s_shear; delta_p += s_p;
        // delta_ncdm = s_rho / rho_n, theta_ncdm = s_theta / (rho_n + p_n)
        delta_rho_m += s_rho; rho_m += rho_n;
        rpt_m += s_theta; rpm += (rho_n + p_n);
      }
    }
  }
  if (want_matter) m.delta_m = delta_rho_m / rho_m + 3. * aH * (rpt_m / rpm) * ik2;

  // ---- perturb_einstein (synchronous gauge, K = 0)
  const double h_prime = (k2 * eta + 1.5 * a2 * delta_rho) * e.drv[DRV_INV_HALF_AH];
  double rsa_delta_g = 0., rsa_theta_g = 0.;
  if (ap.rsa_on) {
    double rsa_delta_ur = 0., rsa_theta_ur = 0.;
    if (P.rsa_method != CLPP_RSA_NULL) {
      rsa_delta_g = 4. * ik2 * (aH * h_prime - k2 * eta);
      rsa_theta_g = -0.5 * h_prime;
    }
    if (P.rsa_method == CLPP_RSA_MD_WITH_REIO) {
      rsa_delta_g += -4. * ik2 * dkappa * (theta_b + 0.5 * h_prime);
      rsa_theta_g += 3. * ik2 * (pvt[P.iddkappa] * (theta_b + 0.5 * h_prime) +
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
  const double eta_prime = (1.5 * a2 * rpt) * ik2;
  const double alpha = (h_prime + 6. * eta_prime) * 0.5 * ik2;
  if (!ap.tca_off) {
    const double sg = 16. / 45. * e.drv[DRV_TAU_C] * (theta_g + k2 * alpha);
    rps += 4. / 3. * rho_g * sg;
  }
  const double alpha_prime = -2. * aH * alpha + eta - 4.5 * (a2 * ik2) * rps;
  m.h_prime = h_prime; m.eta_prime = eta_prime; m.alpha = alpha; m.alpha_prime = alpha_prime;
  m.rsa_delta_g = rsa_delta_g; m.rsa_theta_g = rsa_theta_g;
  if (dy == nullptr) return;

  // ---- perturb_derivs: hub equations (uniform)
  const double cotKgen = e.drv[DRV_INV_TAU] * M.inv_k;
  const double metric_continuity = h_prime * 0.5;
  const double metric_shear = k2 * alpha;
  const double metric_ufa_class = h_prime * 0.5;
  if (ap.rsa_on) { delta_g = rsa_delta_g; theta_g = rsa_theta_g; }

  double dtheta_b, dtheta_g = 0., dshear_g = 0., dl3_g = 0., dp0 = 0., dp1 = 0., dp2 = 0.;
  if (ap.tca_off) {
    dtheta_b = -aH * theta_b + k2 * delta_p_b_over_rho_b + R * dkappa * (theta_g - theta_b);
    if (full_g) {
      const double P0 = (p0 + p2 + 2. * shear_g) * 0.125;
      dtheta_g = k2 * (delta_g * 0.25 - shear_g) + dkappa * (theta_b - theta_g);
      dshear_g = 0.5 * (8. / 15. * (theta_g + metric_shear) - 3. / 5. * k * g3 - dkappa * (2. * shear_g - 4. / 5. * P0));
      dl3_g = k * (1. / 7.0) * (3. * 2. * shear_g - 4. * g4) - dkappa * g3;
      dp0 = -k * p1 - dkappa * (p0 - 4. * P0);
      dp1 = k * (1. / 3.) * (p0 - 2. * p2) - dkappa * p1;
      dp2 = k * (1. / 5.) * (2. * p1 - 3. * p3) - dkappa * (p2 - 4. / 5. * P0);
    }
  } else {
    // ---- perturb_tca_slip_and_shear
    const double a_primeprime_over_a = e.Hp * e.a + 2. * aH * aH;
    const double tau_c = e.drv[DRV_TAU_C];
    const double dtau_c = -pvt[P.iddkappa] * tau_c * tau_c;
    const double i1pR = e.drv[DRV_INV_1PR];
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
    dtheta_g = -(dtheta_b + aH * theta_b - k2 * delta_p_b_over_rho_b) * e.drv[DRV_INV_R] + k2 * (0.25 * delta_g - sg);
  }
  double ddelta_ur = 0., dtheta_ur = 0., dshear_ur = 0., dl3_ur = 0.;
  if (has_ur) {
    ddelta_ur = -4. / 3. * (theta_ur + metric_continuity) +
                (1. - P.three_ceff2_ur) * aH * (delta_ur + 4. * aH * theta_ur * ik2);
    dtheta_ur = k2 * (P.three_ceff2_ur * delta_ur * 0.25 - shear_ur) - (1. - P.three_ceff2_ur) * aH * theta_ur;
    if (!ap.ufa_on) {
      dshear_ur = 0.5 * (8. / 15. * (theta_ur + metric_shear) - 3. / 5. * k * u3 -
                         (1. - P.three_cvis2_ur) * (8. / 15. * (theta_ur + metric_shear)));
      dl3_ur = k * (1. / 7.) * (3. * 2. * shear_ur - 4. * u4);
    } else {
      if (P.ufa_method == CLPP_UFA_MB) dshear_ur = -3. * e.drv[DRV_INV_TAU] * shear_ur + 2. / 3. * (theta_ur + metric_shear);
      else if (P.ufa_method == CLPP_UFA_HU) dshear_ur = -3. * aH * shear_ur + 2. / 3. * (theta_ur + metric_shear);
      else dshear_ur = -3. * e.drv[DRV_INV_TAU] * shear_ur + 2. / 3. * (theta_ur + metric_ufa_class);
    }
  }
  // ---- stores of the hub equations (lane 0) ...
  if (lane == 0) {
    if (has_g) {
      dy[L.delta_g] = -4. / 3. * (theta_g + metric_continuity);
      dy[L.theta_g] = dtheta_g;
      if (full_g) {
        dy[L.shear_g] = dshear_g;
        dy[L.l3_g] = dl3_g;
        dy[L.pol0_g] = dp0;
        dy[L.pol0_g + 1] = dp1;
        dy[L.pol0_g + 2] = dp2;
      }
    }
    dy[L.delta_b] = -(theta_b + metric_continuity);
    dy[L.theta_b] = dtheta_b;
    dy[L.delta_cdm] = -metric_continuity;
    dy[L.eta] = eta_prime;
    if (has_ur) {
      dy[L.delta_ur] = ddelta_ur;
      dy[L.theta_ur] = dtheta_ur;
      dy[L.shear_ur] = dshear_ur;
      if (!ap.ufa_on) dy[L.l3_ur] = dl3_ur;
    }
  }
  // ---- ... and the multipole chains, one element per lane (1/(2l+1) from a shared-memory table)
  const double* i2l1 = s_i2l1(P);
  if (full_g) {
    const int lg = P.l_max_g, lp = P.l_max_pol_g;
    const double* yg = y + L.delta_g;  // yg[l] = F_l (l>=3), yg[2] = shear
    const double* yp = y + L.pol0_g;
#pragma unroll 1
    for (int l = 4 + lane; l <= lg; l += 32) {
      if (l < lg) dy[L.delta_g + l] = k * i2l1[l] * (l * yg[l - 1] - (l + 1) * yg[l + 1]) - dkappa * yg[l];
      else dy[L.delta_g + l] = k * (yg[l - 1] - (1. + l) * cotKgen * yg[l]) - dkappa * yg[l];
    }
#pragma unroll 1
    for (int l = 3 + lane; l <= lp; l += 32) {
      if (l < lp) dy[L.pol0_g + l] = k * i2l1[l] * (l * yp[l - 1] - (l + 1.) * yp[l + 1]) - dkappa * yp[l];
      else dy[L.pol0_g + l] = k * (yp[l - 1] - (l + 1) * cotKgen * yp[l]) - dkappa * yp[l];
    }
  }
  if (has_ur && !ap.ufa_on) {
    const double* yu = y + L.delta_ur;
    const int lu = P.l_max_ur;
#pragma unroll 1
    for (int l = 4 + lane; l <= lu; l += 32) {
      if (l < lu) dy[L.delta_ur + l] = k * i2l1[l] * (l * yu[l - 1] - (l + 1.) * yu[l + 1]);
      else dy[L.delta_ur + l] = k * (yu[l - 1] - (1. + l) * cotKgen * yu[l]);
    }
  }
  if (P.has_ncdm) {
    if (ap.ncdmfa_on) {
      if (lane < P.N_ncdm) {
        const int s = lane;
        const double* nf = M.nf[s];  // [2] w, [4] ca2, [5] ceff2/(1+w), [6] 8/3 cvis2/(1+w), [7] shear damping rate
        const int idx = L.psi0_ncdm1 + 3 * s;
        const double y0 = y[idx], y1 = y[idx + 1], y2 = y[idx + 2];
        const double w_n = nf[2], ca2 = nf[4];
        dy[idx] = -(1.0 + w_n) * (y1 + metric_continuity) - 3.0 * aH * (ca2 - w_n) * y0;
        dy[idx + 1] = -aH * (1.0 - 3.0 * ca2) * y1 + nf[5] * k2 * y0 - k2 * y2;
        const double ms = (P.ncdmfa_method == CLPP_NCDMFA_CLASS) ? metric_ufa_class : metric_shear;
        dy[idx + 2] = -nf[7] * y2 + nf[6] * (y1 + ms);
      }
    } else {
      const int stride = L.l_max_ncdm + 1, lm = L.l_max_ncdm;
      const int tot = L.eta - L.psi0_ncdm1;  // all species are contiguous, same stride
      const double* w = s_nw(P);
#pragma unroll 1
      for (int e_i = lane; e_i < tot; e_i += 32) {
        const int jq = e_i / stride, l = e_i - jq * stride;  // jq = global momentum-bin index
        const int idx = L.psi0_ncdm1 + jq * stride;
        const double qk = k * w[jq];
        double v;
        if (l == 0) v = -qk * y[idx + 1] + metric_continuity * __ldg(M.C->ncdm_dlnf0 + jq) * (1. / 3.);
        else if (l == 1) v = qk * (1. / 3.0) * (y[idx] - 2 * y[idx + 2]);
        else if (l == 2) v = qk * (1. / 5.0) * (2 * y[idx + 1] - 3. * y[idx + 3]) - metric_shear * 2. / 15. * __ldg(M.C->ncdm_dlnf0 + jq);
        else if (l < lm) v = qk * i2l1[l] * (l * y[idx + (l - 1)] - (l + 1.) * y[idx + (l + 1)]);
        else v = qk * y[idx + l - 1] - (1. + l) * k * cotKgen * y[idx + l];
        dy[idx + l] = v;
      }
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// perturb_sources_member: source functions at sample index_tau from (y, dy)
__device__ __noinline__ void write_sources(const PtParams& P, double tau, int slot_y, int slot_dy, int index_tau) {
  Mode& M = MODE(P);
  const double* y = s_vec(P, slot_y);
  const double* dy = s_vec(P, slot_dy);
  env_at(P, tau, true, 1);
  rhs_apply(P, slot_y, -1, 1);
  if (PT_LANE != 0) return;
  const Layout& L = M.L;
  const Approx& ap = M.ap;
  const Env& e = M.e;
  const double* pvt = s_pvt(P);
  const double e_g = pvt[P.ig], e_dg = pvt[P.idg], e_exp_m_kappa = pvt[P.iexp_m_kappa];
  const Metric& m = M.m;
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
  double* out = M.C->sources + (size_t)M.ik * M.C->tau_size + index_tau;
  if (P.tp_t0 >= 0) {
    int switch_isw = 1;
    if ((P.switch_eisw == 0) && (z >= P.eisw_lisw_split_z)) switch_isw = 0;
    if ((P.switch_lisw == 0) && (z < P.eisw_lisw_split_z)) switch_isw = 0;
    const double eta = y[L.eta], theta_b = y[L.theta_b], dtheta_b = dy[L.theta_b];
    out[P.tp_t0 * stride_tp] =
        P.switch_sw * e_g * (delta_g / 4. + m.alpha_prime) +
        switch_isw * (e_g * (eta - m.alpha_prime - 2 * aH * m.alpha) +
                      e_exp_m_kappa * 2. * (m.eta_prime - aH_prime * m.alpha - aH * m.alpha_prime)) +
        P.switch_dop * (e_g * (dtheta_b / k / k + m.alpha_prime) + e_dg * (theta_b / k / k + m.alpha));
    out[P.tp_t1 * stride_tp] = switch_isw * e_exp_m_kappa * k * (m.alpha_prime + 2. * aH * m.alpha - eta);
    out[P.tp_t2 * stride_tp] = P.switch_pol * e_g * Pi;
  }
  if (P.tp_p >= 0) out[P.tp_p * stride_tp] = sqrt(6.) * e_g * Pi;
  if (P.tp_phi_plus_psi >= 0) out[P.tp_phi_plus_psi * stride_tp] = y[L.eta] + m.alpha_prime;
  if (P.tp_delta_m >= 0) out[P.tp_delta_m * stride_tp] = m.delta_m;
  if (P.tp_delta_cb >= 0) out[P.tp_delta_cb * stride_tp] = m.delta_cb;
}

// ---------------------------------------------------------------------------------------------
// Jacobian J = A(tau): hub columns f(tau, e_j) one by one, chain entries with 3 grouped probes
// (the environment M.e must be set at tau)
__device__ __noinline__ void jacobian(const PtParams& P) {
  Mode& M = MODE(P);
  const int n = M.L.neq, lane = PT_LANE, nh = M.nh, nch = M.nch;
  double* e_j = s_vec(P, V_TMP);
  double* col = s_vec(P, V_DEL);
  double *Jd = M.gJd, *Jl = s_vec(P, V_JL), *Ju = M.gJu;
#pragma unroll 1
  for (int i = lane; i < n; i += 32) { e_j[i] = 0.; Jd[i] = 0.; Jl[i] = 0.; Ju[i] = 0.; }
  __syncwarp();
  for (int j = 0; j < nh; j++) {
    const int hj = s_hub_idx(P)[j];
    if (lane == 0) e_j[hj] = 1.;
    __syncwarp();
    rhs_apply(P, V_TMP, V_DEL, 0);
#pragma unroll 1
    for (int s = lane; s < nh; s += 32) M.Jhh[s + (size_t)j * nh] = col[s_hub_idx(P)[s]];
    if (lane < nch && s_ch_start(P)[lane] - 1 == hj) Jl[hj + 1] = col[hj + 1];  // chain start <- its root
    if (lane == 0) e_j[hj] = 0.;
    __syncwarp();
  }
  if (nch > 0) {
    for (int r = 0; r < 3; r++) {
      if (lane < nch) {
        const int s = s_ch_start(P)[lane], len = s_ch_len(P)[lane];
        for (int p = r; p < len; p += 3) e_j[s + p] = 1.;
      }
      __syncwarp();
      rhs_apply(P, V_TMP, V_DEL, 0);
      if (lane < nch) {
        const int s = s_ch_start(P)[lane], len = s_ch_len(P)[lane];
        for (int p = 0; p < len; p++) {
          const int i = s + p, pm = p % 3;
          if (pm == r) Jd[i] = col[i];
          if ((pm + 1) % 3 == r && p + 1 < len) Ju[i] = col[i];
          if ((pm + 2) % 3 == r && p >= 1) Jl[i] = col[i];
        }
        if (r == 0) Ju[s - 1] = col[s - 1];  // root <- chain start
        for (int p = r; p < len; p += 3) e_j[s + p] = 0.;
      }
      __syncwarp();
    }
  }
  if (PT_LANE == 0) M.st.jacobians++;
  if (PT_LANE == 0) M.st.fevals += nh + (nch > 0 ? 3 : 0);
}

// Gauss-Jordan inversion with partial pivoting of an n x n matrix held one row per lane in registers
// (n <= N <= 32).  On exit G = (P A)^-1 where P are the row exchanges, returned as `perm`
// (A^-1 r = G (P r), (P r)_j = r[perm_j]; equivalently A^-1[i][perm_j] = G[i][j]).
template <int N>
__device__ __forceinline__ void gj_rows(double (&G)[N], int n, int lane, int& perm) {
  perm = lane;
#pragma unroll 1
  for (int j = 0; j < n; j++) {
    double gj = G[0];
#pragma unroll
    for (int q = 1; q < N; q++) gj = (q == j) ? G[q] : gj;
    // pivot: largest |G[.][j]| among rows j..n-1
    const double v = (lane >= j && lane < n) ? fabs(gj) : -1.0;
    const double vmax = wmax_any(v);
    int pj = __ffs(__ballot_sync(PT_FULL, v == vmax)) - 1;
    if (pj < 0) pj = j;
    if (pj != j) {  // exchange rows j and pj (uniform branch)
      const int src = (lane == j) ? pj : (lane == pj) ? j : lane;
#pragma unroll
      for (int q = 0; q < N; q++) G[q] = __shfl_sync(PT_FULL, G[q], src);
      perm = __shfl_sync(PT_FULL, perm, src);
      gj = __shfl_sync(PT_FULL, gj, src);
    }
    double pv = __shfl_sync(PT_FULL, gj, j);
    if (pv == 0.) pv = 1e-50;  // TINY, as ludcmp does for a singular pivot
    const double pinv = 1.0 / pv;
    const double f = (lane == j) ? 0. : gj;
#pragma unroll
    for (int q = 0; q < N; q++) {
      const double pq = __shfl_sync(PT_FULL, G[q], j) * pinv;
      const double upd = (lane == j) ? pq : G[q] - f * pq;
      const double piv = (lane == j) ? pinv : -f * pinv;
      G[q] = (q == j) ? piv : upd;
    }
  }
}

template <int N>
__device__ __forceinline__ void hub_inverse_rows(const PtParams& P, Mode& M, double c, int nh, int lane) {
  double G[N];
  const int li = lane < nh ? lane : 0;
  const double dg = 1.0 + s_hubtmp(P)[li];
#pragma unroll
  for (int q = 0; q < N; q++) {
    const double jv = (q < nh && lane < nh) ? M.Jhh[li + (size_t)q * nh] : 0.;
    G[q] = ((q == lane) ? (lane < nh ? dg : 1.0) : 0.0) - c * jv;
  }
  int perm;
  gj_rows<N>(G, nh, lane, perm);
  double* W = s_sinv(P);
  const int ldh = P.ldh;
#pragma unroll
  for (int q = 0; q < N; q++) {
    const int pq = __shfl_sync(PT_FULL, perm, q);
    if (q < nh && lane < nh) W[lane * ldh + pq] = G[q];
  }
  __syncwarp();
}

// Factorisation of A = I - c J: chains (backward elimination towards their root), Schur
// complement on the hub block, explicit inverse of the hub block (Gauss-Jordan, partial pivoting).
__device__ __noinline__ void factor(const PtParams& P, double c) {
  Mode& M = MODE(P);
  const int lane = PT_LANE, nh = M.nh, nch = M.nch, ldh = P.ldh;
  const double *Jd = M.gJd, *Jl = s_vec(P, V_JL), *Ju = M.gJu;
  double *ip = s_vec(P, V_IP), *mu = s_vec(P, V_MU);
  if (lane == 0) M.fac_c = c;
  double* W = s_sinv(P);
#pragma unroll 1
  for (int i = lane; i < nh; i += 32) s_hubtmp(P)[i] = 0.;
  __syncwarp();
  if (lane < nch) {
    const int s = s_ch_start(P)[lane], last = s + s_ch_len(P)[lane] - 1;
    double ipn = 1.0 / (1.0 - c * Jd[last]);
    ip[last] = ipn;
    double lon = -c * Jl[last];
    for (int i = last - 1; i >= s; i--) {
      const double mui = -c * Ju[i] * ipn;
      const double p = (1.0 - c * Jd[i]) - mui * lon;
      ipn = 1.0 / p;
      lon = -c * Jl[i];
      ip[i] = ipn; mu[i] = mui;
    }
    const double mur = -c * Ju[s - 1] * ipn;
    mu[s - 1] = mur;
    s_hubtmp(P)[s_ch_rootslot(P)[lane]] = -mur * lon;
  }
  __syncwarp();
  // hub block W = I - c Jhh (+ Schur terms on the diagonal of the chain roots); lane i owns row i.
  // Up to 32 hub variables: the rows stay in registers and are exchanged with shuffles.
  if (nh <= 16) { hub_inverse_rows<16>(P, M, c, nh, lane); if (lane == 0) M.st.factorizations++; return; }
  if (nh <= 32) { hub_inverse_rows<32>(P, M, c, nh, lane); if (lane == 0) M.st.factorizations++; return; }
#pragma unroll 1
  for (int i = lane; i < nh; i += 32) {
    for (int j = 0; j < nh; j++) W[i * ldh + j] = (i == j ? 1.0 + s_hubtmp(P)[i] : 0.0) - c * M.Jhh[i + (size_t)j * nh];
  }
  __syncwarp();
  for (int j = 0; j < nh; j++) {
    // pivot search in column j, rows >= j
    double best = -1.;
    int bi = j;
#pragma unroll 1
    for (int i = j + lane; i < nh; i += 32) {
      const double v = fabs(W[i * ldh + j]);
      if (v > best) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(PT_FULL, best, o);
      const int oi = __shfl_xor_sync(PT_FULL, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) s_piv(P)[j] = bi;
    if (bi != j) {
#pragma unroll 1
      for (int cc = lane; cc < nh; cc += 32) {
        const double t = W[j * ldh + cc];
        W[j * ldh + cc] = W[bi * ldh + cc];
        W[bi * ldh + cc] = t;
      }
      __syncwarp();
    }
    double pv = W[j * ldh + j];
    if (pv == 0.) pv = 1e-50;  // TINY, as ludcmp does for a singular pivot
    const double pinv = 1.0 / pv;
    __syncwarp();
#pragma unroll 1
    for (int cc = lane; cc < nh; cc += 32) W[j * ldh + cc] = (cc == j) ? pinv : W[j * ldh + cc] * pinv;
    __syncwarp();
#pragma unroll 1
    for (int i = lane; i < nh; i += 32) {
      if (i != j) {
        const double f = W[i * ldh + j];
        if (f != 0.) {
          for (int cc = 0; cc < nh; cc++) {
            const double wj = W[j * ldh + cc];
            W[i * ldh + cc] = (cc == j) ? -f * wj : W[i * ldh + cc] - f * wj;
          }
        }
      }
    }
    __syncwarp();
  }
  for (int j = nh - 1; j >= 0; j--) {
    const int p = s_piv(P)[j];
    if (p != j) {
#pragma unroll 1
      for (int i = lane; i < nh; i += 32) {
        const double t = W[i * ldh + j];
        W[i * ldh + j] = W[i * ldh + p];
        W[i * ldh + p] = t;
      }
      __syncwarp();
    }
  }
  if (PT_LANE == 0) M.st.factorizations++;
}

// solve A x = b in place (b in shared memory) with the factors of `factor`
__device__ __forceinline__ void solve(const PtParams& P, double* __restrict__ b) {
  Mode& M = MODE(P);
  const int lane = PT_LANE, nh = M.nh, nch = M.nch, ldh = P.ldh;
  const double *ip = s_vec(P, V_IP), *mu = s_vec(P, V_MU), *Jl = s_vec(P, V_JL);
  const double c = M.fac_c;
  if (nch > 0) {
    if (lane < nch) {
      const int s = s_ch_start(P)[lane], last = s + s_ch_len(P)[lane] - 1;
      double r = b[last];
      for (int i = last - 1; i >= s; i--) {
        r = b[i] - mu[i] * r;
        b[i] = r;
      }
      b[s - 1] -= mu[s - 1] * r;
    }
    __syncwarp();
  }
  if (nh <= 32) {
    double x = 0.;
    if (lane < nh) {
      const double* w = s_sinv(P) + lane * ldh;
      double x0 = 0., x1 = 0.;
      int j = 0;
      for (; j + 1 < nh; j += 2) {
        x0 += w[j] * b[s_hub_idx(P)[j]];
        x1 += w[j + 1] * b[s_hub_idx(P)[j + 1]];
      }
      if (j < nh) x0 += w[j] * b[s_hub_idx(P)[j]];
      x = x0 + x1;
    }
    __syncwarp();
    if (lane < nh) b[s_hub_idx(P)[lane]] = x;
  } else {
    double xs[4];  // up to 128 hub variables
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int i = lane + 32 * t;
      double x = 0.;
      if (i < nh) {
        const double* w = s_sinv(P) + i * ldh;
        for (int j = 0; j < nh; j++) x += w[j] * b[s_hub_idx(P)[j]];
      }
      xs[t] = x;
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int i = lane + 32 * t;
      if (i < nh) b[s_hub_idx(P)[i]] = xs[t];
    }
  }
  __syncwarp();
  if (nch > 0) {
    if (lane < nch) {
      const int s = s_ch_start(P)[lane], last = s + s_ch_len(P)[lane] - 1;
      double xp = b[s - 1];
      for (int i = s; i <= last; i++) {
        const double lo = -c * Jl[i];  // T[i,i-1]
        xp = (b[i] - lo * xp) * ip[i];
        b[i] = xp;
      }
    }
    __syncwarp();
  }
  if (PT_LANE == 0) M.st.solves++;
}

// rescale the backward differences when the step changes by the factor r (k = current order)
__device__ __noinline__ void adjust_stepsize(const PtParams& P, double r, int k) {
  Mode& M = MODE(P);
  // RU = R(r) * U; lane t < 25 computes entry (t/5, t%5) and parks it in shared memory
  double* RU = s_hubtmp(P);  // >= 32 doubles
  if (PT_LANE < 25) {
    const int ii = PT_LANE / 5, jj = PT_LANE % 5;
    double s = 0.;
#pragma unroll
    for (int kk = 0; kk < 5; kk++) {
      // R[ii][kk] = prod_{m=1..ii+1} (m - 1 - (kk+1) r) / m
      double Rv = 1.;
      for (int mm = 1; mm <= ii + 1; mm++) Rv *= ((mm - 1) - (kk + 1) * r) / mm;
      s += Rv * c_U[kk][jj];
    }
    RU[PT_LANE] = s;
  }
  __syncwarp();
  const int n = M.L.neq, np = P.np;
  double* dif = s_vec(P, V_DIF0);
  for (int i = PT_LANE; i < n; i += 32) {
    double row[5];
#pragma unroll
    for (int kk = 0; kk < 5; kk++) row[kk] = (kk < k) ? dif[kk * np + i] : 0.;
#pragma unroll
    for (int jj = 0; jj < 5; jj++) {
      if (jj < k) {
        double s = 0.0;
#pragma unroll
        for (int kk = 0; kk < 5; kk++) s += row[kk] * RU[kk * 5 + jj];
        dif[jj * np + i] = s;
      }
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// NDF1-5 over [t0, tfinal] for the current layout; y in V_Y (in/out). `next` = index of the next
// source sample time (carried across intervals). Returns false on failure.
__device__ __noinline__ bool ndf15(const PtParams& P, double t0, double tfinal) {
  Mode& M = MODE(P);
  PROF_DECL;
  const double abstol = 1e-15, eps = 1e-16, threshold = abstol;
  const int maxit = 4, maxk = 5;
  const double rtol = P.rtol;
  const int n = M.L.neq, np = P.np, lane = PT_LANE;
  double *y = s_vec(P, V_Y), *ynew = s_vec(P, V_YNEW), *f0 = s_vec(P, V_F0), *pred = s_vec(P, V_PRED), *psi = s_vec(P, V_PSI),
         *difkp1 = s_vec(P, V_DIFKP1), *del = s_vec(P, V_DEL), *invwt = s_vec(P, V_INVWT), *dif = s_vec(P, V_DIF0);
  const double* t_vec = M.C->tau;
  const int tres = M.C->tau_size;
  int next = M.next;
  while (next < tres && __ldg(t_vec + next) < t0) next++;
  double tnext = (next < tres) ? __ldg(t_vec + next) : 1e300;

  double* const gdif = M.gdif;
#define DIF_SLOT(j) ((j) < 5 ? dif + (j) * np : gdif + ((j) - 5) * np)
  for (int j = 0; j < 7; j++) {
    double* dj = DIF_SLOT(j);
#pragma unroll 1
    for (int i = lane; i < n; i += 32) dj[i] = 0.;
  }
  const double htspan = fabs(tfinal - t0);
  double t = t0, tnew = t0;
  env_at(P, t0, true, 0);
  rhs_apply(P, V_Y, V_F0, 0);
  if (PT_LANE == 0) M.st.fevals++;
  const double hmax = (tfinal - t0) / 10.0;
  jacobian(P);
  bool Jcurrent = true;
  double hmin = 16.0 * eps * fabs(t);
  // initial step from |f0/wt| and the second derivative estimate
  double rh = 0.0;
#pragma unroll 1
  for (int i = lane; i < n; i += 32) {
    const double wt = fmax(fabs(y[i]), threshold);
    rh = fmax(rh, 1.25 / sqrt(rtol) * fabs(f0[i] / wt));
  }
  rh = wmax(rh);
  double absh = fmin(hmax, htspan);
  if (absh * rh > 1.0) absh = 1.0 / rh;
  absh = fmax(absh, hmin);
  double h = absh;
  {
    // J*f0 = f(t0, f0): the system is linear and homogeneous
    rhs_apply(P, V_F0, V_PSI, 0);
    if (PT_LANE == 0) M.st.fevals++;
    const double tdel = (t + fmin(sqrt(eps) * fmax(fabs(t), fabs(t + h)), absh)) - t;
    env_at(P, t + tdel, true, 0);
    rhs_apply(P, V_Y, V_DEL, 0);  // f(t+tdel, y)
    if (PT_LANE == 0) M.st.fevals++;
    rh = 0.0;
#pragma unroll 1
    for (int i = lane; i < n; i += 32) {
      const double wt = fmax(fabs(y[i]), threshold);
      const double s = psi[i] + (del[i] - f0[i]) / tdel;
      rh = fmax(rh, 1.25 * sqrt(0.5 * fabs(s / wt) / rtol));
    }
    rh = wmax(rh);
    absh = fmin(hmax, htspan);
    if (absh * rh > 1.0) absh = 1.0 / rh;
    absh = fmax(absh, hmin);
    h = absh;
  }
  int k = 1, klast = k;
  double abshlast = absh;
#pragma unroll 1
  for (int i = lane; i < n; i += 32) dif[0 * np + i] = h * f0[i];
  __syncwarp();
  double hinvGak = h * c_invGa[k - 1];
  int nconhk = 0;
  factor(P, hinvGak);
  bool havrate = false;
  bool done = false, at_hmin = false;
  double rate = 0., oldnrm = 0., err = 0.;

  int sync_ctr = 0;
  while (!done) {
    PT_COHORT_SYNC(P);
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
      PROF_BEGIN();
      adjust_stepsize(P, absh / abshlast, k);
      PROF_END(PF_ADJUST);
      hinvGak = h * c_invGa[k - 1];
      nconhk = 0;
      PROF_BEGIN();
      factor(P, hinvGak);
      PROF_END(PF_FACTOR);
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
        PROF_BEGIN();
        const double invGak = c_invGa[k - 1];
#pragma unroll 1
        for (int i = lane; i < n; i += 32) {
          double ps = 0.0;
          const double yi = y[i];
          double pr = yi;
          for (int j = 0; j < k; j++) {
            const double d = dif[j * np + i];
            ps += d * c_G[j] * invGak;
            pr += d;
          }
          psi[i] = ps;
          pred[i] = pr;
          ynew[i] = pr;
          difkp1[i] = 0.0;
          const double iw = 1.0 / fmax(fmax(fabs(pr), fabs(yi)), threshold);
          invwt[i] = iw;
          minnrm = fmax(minnrm, 100 * eps * fabs(pr * iw));
        }
        minnrm = wmax(minnrm);
        __syncwarp();
        PROF_END(PF_PREDICT);
        PROF_BEGIN();
        env_at(P, tnew, true, 0);
        PROF_END(PF_ENV);
        bool tooslow = false;
        for (int iter = 1; iter <= maxit; iter++) {
          PROF_BEGIN();
          rhs_apply(P, V_YNEW, V_F0, 0);
          PROF_END(PF_RHS);
          if (!M.ap.tca_off) M.tca_shear_last = M.m.tca_shear_g;
          if (PT_LANE == 0) M.st.fevals++;
          PROF_BEGIN();
#pragma unroll 1
          for (int i = lane; i < n; i += 32) del[i] = hinvGak * f0[i] - (psi[i] + difkp1[i]);
          __syncwarp();
          solve(P, del);
          PROF_END(PF_SOLVE);
          PROF_BEGIN();
          double newnrm = 0.0;
#pragma unroll 1
          for (int i = lane; i < n; i += 32) {
            const double d = del[i];
            newnrm = fmax(newnrm, fabs(d * invwt[i]));
            const double dk = difkp1[i] + d;
            difkp1[i] = dk;
            ynew[i] = pred[i] + dk;
          }
          newnrm = wmax(newnrm);
          __syncwarp();
          PROF_END(PF_UPDATE);
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
          if (PT_LANE == 0) M.st.failed++;
          if (!Jcurrent) {
            env_at(P, t, true, 0);
            rhs_apply(P, V_Y, V_F0, 0);
            if (PT_LANE == 0) M.st.fevals++;
            jacobian(P);
            Jcurrent = true;
          } else if (absh <= hmin) {
            M.status = 2;  // step size too small
            return false;
          } else {
            abshlast = absh;
            absh = fmax(0.3 * absh, hmin);
            h = absh;
            done = false;
            adjust_stepsize(P, absh / abshlast, k);
            hinvGak = h * c_invGa[k - 1];
            nconhk = 0;
          }
          factor(P, hinvGak);
          havrate = false;
        }
      }
      // error estimate
      err = 0.0;
#pragma unroll 1
      for (int i = lane; i < n; i += 32) err = fmax(err, fabs(difkp1[i] * invwt[i]));
      err = wmax(err) * c_erconst[k - 1];
      if (err > rtol) {
        if (PT_LANE == 0) M.st.failed++;
        if (absh <= hmin) {
          M.status = 2;
          return false;
        }
        abshlast = absh;
        if (nofailed) {
          nofailed = false;
          double hopt = absh * fmax(0.1, 0.833 * root_n(rtol / err, k + 1.0));
          if (k > 1) {
            double errkm1 = 0.0;
#pragma unroll 1
            for (int i = lane; i < n; i += 32) errkm1 = fmax(errkm1, fabs((dif[(k - 1) * np + i] + difkp1[i]) * invwt[i]));
            errkm1 = wmax(errkm1) * c_erconst[k - 2];
            const double hkm1 = absh * fmax(0.1, 0.769 * root_n(rtol / errkm1, (double)k));
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
        adjust_stepsize(P, absh / abshlast, k);
        hinvGak = h * c_invGa[k - 1];
        nconhk = 0;
        factor(P, hinvGak);
        havrate = false;
      } else {
        break;
      }
    }
    if (PT_LANE == 0) M.st.steps++;
    PROF_BEGIN();
    // update the difference array
    double* const difk = DIF_SLOT(k);
    double* const difk1 = DIF_SLOT(k + 1);
#pragma unroll 1
    for (int i = lane; i < n; i += 32) {
      const double dk = difkp1[i];
      difk1[i] = dk - difk[i];
      double acc = dk;
      difk[i] = acc;
      for (int j = k - 1; j >= 0; j--) {
        acc += dif[j * np + i];
        dif[j * np + i] = acc;
      }
    }
    __syncwarp();
    PROF_END(PF_DIFUPD);
    PROF_BEGIN();
    // ---- output at the sample times passed by this step
    while ((next < tres) && ((tnew - tnext) >= 0.0)) {
      if (tnew == tnext) {
        write_sources(P, tnext, V_YNEW, V_F0, next);
      } else {
        const double s = (tnext - tnew) / h;
        double* yi = s_vec(P, V_TMP);
        double* ypi = s_vec(P, V_YPI);
#pragma unroll 1
        for (int i = lane; i < n; i += 32) {
          double a1 = 0, a2 = 0;
          double prod = 1.0, sumfrac = 0., fact = 1.0;
          for (int j = 0; j < k; j++) {
            prod *= (s + j);
            fact *= (j + 1);
            sumfrac += 1.0 / (s + j);
            const double d = dif[j * np + i];
            a1 += prod / fact * d;
            a2 += prod * sumfrac / (h * fact) * d;
          }
          yi[i] = ynew[i] + a1;
          ypi[i] = a2;
        }
        __syncwarp();
        write_sources(P, tnext, V_TMP, V_YPI, next);
      }
      next++;
      tnext = (next < tres) ? __ldg(t_vec + next) : 1e300;
    }
    PROF_END(PF_OUTPUT);
    if (done) break;
    PROF_BEGIN();
    klast = k;
    abshlast = absh;
    nconhk = min(nconhk + 1, maxk + 2);
    if (nconhk >= k + 2) {
      // candidate steps at orders k, k-1, k+1: the three norms by warp reductions, the three
      // roots evaluated side by side in lanes 0..2
      double e_km1 = 0., e_kp1 = 0.;
      if (k > 1) {
#pragma unroll 1
        for (int i = lane; i < n; i += 32) e_km1 = fmax(e_km1, fabs(dif[(k - 1) * np + i] * invwt[i]));
        e_km1 = wmax(e_km1) * c_erconst[k - 2];
      }
      if (k < maxk) {
        const double* difk1 = DIF_SLOT(k + 1);
#pragma unroll 1
        for (int i = lane; i < n; i += 32) e_kp1 = fmax(e_kp1, fabs(difk1[i] * invwt[i]));
        e_kp1 = wmax(e_kp1) * c_erconst[k];
      }
      const double my_e = lane == 0 ? err : lane == 1 ? e_km1 : e_kp1;
      const double my_c = lane == 0 ? 1.2 : lane == 1 ? 1.3 : 1.4;
      const double my_n = lane == 0 ? k + 1.0 : lane == 1 ? (double)k : k + 2.0;
      double temp = 0.;
      if (my_e > 0.) temp = my_c * root_n(my_e / rtol, my_n);
      const double my_h = (temp > 0.1) ? absh / temp : 10 * absh;
      double hopt = __shfl_sync(PT_FULL, my_h, 0);
      const double hkm1 = __shfl_sync(PT_FULL, my_h, 1), hkp1 = __shfl_sync(PT_FULL, my_h, 2);
      int kopt = k;
      if (k > 1 && hkm1 > hopt) { hopt = hkm1; kopt = k - 1; }
      if (k < maxk && hkp1 > hopt) { hopt = hkp1; kopt = k + 1; }
      if (hopt > absh) {
        absh = hopt;
        if (k != kopt) k = kopt;
      }
    }
    t = tnew;
#pragma unroll 1
    for (int i = lane; i < n; i += 32) y[i] = ynew[i];
    __syncwarp();
    Jcurrent = false;
    PROF_END(PF_CONTROL);
  }
  // final state: y <- ynew, and a last RHS call so that the environment and the TCA/RSA
  // by-products are current at the end of the interval (evolver_ndf15.cpp:653-662)
#pragma unroll 1
  for (int i = lane; i < n; i += 32) y[i] = ynew[i];
  __syncwarp();
  env_at(P, tnew, true, 0);
  rhs_apply(P, V_Y, V_F0, 0);
  if (!M.ap.tca_off) M.tca_shear_last = M.m.tca_shear_g;
  if (PT_LANE == 0) M.st.fevals++;
  M.next = next;
  return true;
}


// ---------------------------------------------------------------------------------------------
// NDF1-5 for a HUB-ONLY phase (no multipole chains, neq <= 32): radiation streaming approximation
// with the ncdm fluid approximation, i.e. the last and by far the longest interval of every
// high-k mode (10^4..10^5 steps).  Lane i owns equation i: the state, the backward differences and
// all step-control vectors live in REGISTERS; shared memory is only the exchange buffer of the RHS
// and of the linear solve.  Same algorithm and constants as ndf15() above.
__device__ __forceinline__ double pick7(const double (&d)[7], int j) {
  double v = d[0];
#pragma unroll
  for (int q = 1; q < 7; q++) v = (j == q) ? d[q] : v;
  return v;
}

__device__ __noinline__ bool ndf15_hub(const PtParams& P, double t0, double tfinal) {
  Mode& M = MODE(P);
  PROF_DECL;
  const double abstol = 1e-15, eps = 1e-16, threshold = abstol;
  const int maxit = 4, maxk = 5;
  const double rtol = P.rtol;
  const int n = M.L.neq, lane = PT_LANE, ldh = P.ldh;
  const bool act = lane < n;
  const int li = act ? lane : 0;  // safe shared-memory index for idle lanes
  double *YX = s_vec(P, V_YNEW), *F = s_vec(P, V_F0), *B = s_vec(P, V_DEL), *Ys = s_vec(P, V_Y);
  const double* t_vec = M.C->tau;
  const int tres = M.C->tau_size;
  int next = M.next;
  while (next < tres && __ldg(t_vec + next) < t0) next++;
  double tnext = (next < tres) ? __ldg(t_vec + next) : 1e300;

  double y = act ? Ys[li] : 0.;
  double dif[7];
#pragma unroll
  for (int j = 0; j < 7; j++) dif[j] = 0.;
  const double htspan = fabs(tfinal - t0);
  double t = t0, tnew = t0;
  env_at(P, t0, true, 0);
  rhs_apply(P, V_Y, V_F0, 0);
  if (lane == 0) M.st.fevals++;
  const double f0i = act ? F[li] : 0.;
  const double hmax = (tfinal - t0) / 10.0;
  jacobian(P);
  bool Jcurrent = true;
  double hmin = 16.0 * eps * fabs(t);
  const double wt0 = fmax(fabs(y), threshold);
  double rh = wmax(act ? 1.25 / sqrt(rtol) * fabs(f0i / wt0) : 0.);
  double absh = fmin(hmax, htspan);
  if (absh * rh > 1.0) absh = 1.0 / rh;
  absh = fmax(absh, hmin);
  double h = absh;
  {
    rhs_apply(P, V_F0, V_PSI, 0);  // J*f0 = f(t0, f0): the system is linear and homogeneous
    if (lane == 0) M.st.fevals++;
    const double jf = act ? s_vec(P, V_PSI)[li] : 0.;
    const double tdel = (t + fmin(sqrt(eps) * fmax(fabs(t), fabs(t + h)), absh)) - t;
    env_at(P, t + tdel, true, 0);
    rhs_apply(P, V_Y, V_DEL, 0);  // f(t+tdel, y)
    if (lane == 0) M.st.fevals++;
    const double s = act ? jf + (B[li] - f0i) / tdel : 0.;
    rh = wmax(1.25 * sqrt(0.5 * fabs(s / wt0) / rtol));
    absh = fmin(hmax, htspan);
    if (absh * rh > 1.0) absh = 1.0 / rh;
    absh = fmax(absh, hmin);
    h = absh;
  }
  int k = 1, klast = k;
  double abshlast = absh;
  dif[0] = h * f0i;
  double hinvGak = h * c_invGa[k - 1];
  int nconhk = 0;
  factor(P, hinvGak);
  bool havrate = false;
  bool done = false, at_hmin = false;
  double rate = 0., oldnrm = 0., err = 0.;
  double pred = 0., psi = 0., dk1 = 0., iw = 0.;

  // difference-array rescaling in registers (adjust_stepsize)
  auto rescale = [&](double r) {
    double* RU = s_hubtmp(P);
    if (lane < 25) {
      const int ii = lane / 5, jj = lane % 5;
      double s = 0.;
#pragma unroll
      for (int kk = 0; kk < 5; kk++) {
        double Rv = 1.;
        for (int mm = 1; mm <= ii + 1; mm++) Rv *= ((mm - 1) - (kk + 1) * r) / mm;
        s += Rv * c_U[kk][jj];
      }
      RU[lane] = s;
    }
    __syncwarp();
    double nd[5];
#pragma unroll
    for (int jj = 0; jj < 5; jj++) {
      double s = 0.;
#pragma unroll
      for (int kk = 0; kk < 5; kk++) s += ((kk < k) ? dif[kk] : 0.) * RU[kk * 5 + jj];
      nd[jj] = s;
    }
#pragma unroll
    for (int jj = 0; jj < 5; jj++)
      if (jj < k) dif[jj] = nd[jj];
    __syncwarp();
  };

  int sync_ctr = 0;
  while (!done) {
    PT_COHORT_SYNC(P);
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
      PROF_BEGIN();
      rescale(absh / abshlast);
      PROF_END(PF_ADJUST);
      hinvGak = h * c_invGa[k - 1];
      nconhk = 0;
      PROF_BEGIN();
      factor(P, hinvGak);
      PROF_END(PF_FACTOR);
      havrate = false;
    }
    bool nofailed = true;
    for (;;) {  // loop for advancing one step
      bool gotynew = false;
      while (!gotynew) {
        tnew = t + h;
        if (done) tnew = tfinal;
        h = tnew - t;
        PROF_BEGIN();
        const double invGak = c_invGa[k - 1];
        psi = 0.;
        pred = y;
#pragma unroll
        for (int j = 0; j < 5; j++) {
          if (j < k) {
            psi += dif[j] * c_G[j] * invGak;
            pred += dif[j];
          }
        }
        dk1 = 0.;
        iw = 1.0 / fmax(fmax(fabs(pred), fabs(y)), threshold);
        const double minnrm = wmax(act ? 100 * eps * fabs(pred * iw) : 0.);
        if (act) YX[li] = pred;
        __syncwarp();
        PROF_END(PF_PREDICT);
        PROF_BEGIN();
        env_at(P, tnew, true, 0);
        PROF_END(PF_ENV);
        bool tooslow = false;
        for (int iter = 1; iter <= maxit; iter++) {
          PROF_BEGIN();
          rhs_apply(P, V_YNEW, V_F0, 0);
          PROF_END(PF_RHS);
          if (lane == 0) M.st.fevals++;
          PROF_BEGIN();
          // residual into the exchange buffer, then x_i = sum_j Ainv[i][j] r_j (explicit hub inverse)
          if (act) B[li] = hinvGak * F[li] - (psi + dk1);
          __syncwarp();
          double x0 = 0., x1 = 0.;
          {
            const double* w = s_sinv(P) + li * ldh;
            int j = 0;
            for (; j + 1 < n; j += 2) {
              x0 += w[j] * B[j];
              x1 += w[j + 1] * B[j + 1];
            }
            if (j < n) x0 += w[j] * B[j];
          }
          const double d = act ? x0 + x1 : 0.;
          if (lane == 0) M.st.solves++;
          PROF_END(PF_SOLVE);
          PROF_BEGIN();
          const double newnrm = wmax(fabs(d * iw));
          dk1 += d;
          if (act) YX[li] = pred + dk1;
          __syncwarp();
          PROF_END(PF_UPDATE);
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
        if (tooslow) {
          if (lane == 0) M.st.failed++;
          if (!Jcurrent) {
            if (act) Ys[li] = y;
            __syncwarp();
            env_at(P, t, true, 0);
            rhs_apply(P, V_Y, V_F0, 0);
            if (lane == 0) M.st.fevals++;
            jacobian(P);
            Jcurrent = true;
          } else if (absh <= hmin) {
            M.status = 2;  // step size too small
            return false;
          } else {
            abshlast = absh;
            absh = fmax(0.3 * absh, hmin);
            h = absh;
            done = false;
            rescale(absh / abshlast);
            hinvGak = h * c_invGa[k - 1];
            nconhk = 0;
          }
          factor(P, hinvGak);
          havrate = false;
        }
      }
      // error estimate
      err = wmax(fabs(dk1 * iw)) * c_erconst[k - 1];
      if (err > rtol) {
        if (lane == 0) M.st.failed++;
        if (absh <= hmin) {
          M.status = 2;
          return false;
        }
        abshlast = absh;
        if (nofailed) {
          nofailed = false;
          double hopt = absh * fmax(0.1, 0.833 * root_n(rtol / err, k + 1.0));
          if (k > 1) {
            const double errkm1 = wmax(fabs((pick7(dif, k - 1) + dk1) * iw)) * c_erconst[k - 2];
            const double hkm1 = absh * fmax(0.1, 0.769 * root_n(rtol / errkm1, (double)k));
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
        rescale(absh / abshlast);
        hinvGak = h * c_invGa[k - 1];
        nconhk = 0;
        factor(P, hinvGak);
        havrate = false;
      } else {
        break;
      }
    }
    if (lane == 0) M.st.steps++;
    PROF_BEGIN();
    // update the difference array: dif[k+1] = dk1 - dif[k]; dif[k] = dk1; dif[j] += dif[j+1] (j < k)
    {
      const double difk_old = pick7(dif, k);
#pragma unroll
      for (int j = 6; j >= 0; j--) {
        if (j == k + 1) dif[j] = dk1 - difk_old;
        else if (j == k) dif[j] = dk1;
        else if (j < k) dif[j] += dif[j + 1];
      }
    }
    PROF_END(PF_DIFUPD);
    PROF_BEGIN();
    // ---- output at the sample times passed by this step (YX holds ynew, F the last Newton RHS)
    while ((next < tres) && ((tnew - tnext) >= 0.0)) {
      if (tnew == tnext) {
        if (act) YX[li] = pred + dk1;
        __syncwarp();
        write_sources(P, tnext, V_YNEW, V_F0, next);
      } else {
        const double s = (tnext - tnew) / h;
        double a1 = 0, a2 = 0;
        double prod = 1.0, sumfrac = 0., fact = 1.0;
#pragma unroll
        for (int j = 0; j < 5; j++) {
          if (j < k) {
            prod *= (s + j);
            fact *= (j + 1);
            sumfrac += 1.0 / (s + j);
            a1 += prod / fact * dif[j];
            a2 += prod * sumfrac / (h * fact) * dif[j];
          }
        }
        if (act) {
          s_vec(P, V_TMP)[li] = (pred + dk1) + a1;
          s_vec(P, V_YPI)[li] = a2;
        }
        __syncwarp();
        write_sources(P, tnext, V_TMP, V_YPI, next);
      }
      next++;
      tnext = (next < tres) ? __ldg(t_vec + next) : 1e300;
    }
    PROF_END(PF_OUTPUT);
    if (done) break;
    PROF_BEGIN();
    klast = k;
    abshlast = absh;
    nconhk = min(nconhk + 1, maxk + 2);
    if (nconhk >= k + 2) {
      double e_km1 = 0., e_kp1 = 0.;
      if (k > 1) e_km1 = wmax(fabs(pick7(dif, k - 1) * iw)) * c_erconst[k - 2];
      if (k < maxk) e_kp1 = wmax(fabs(pick7(dif, k + 1) * iw)) * c_erconst[k];
      const double my_e = lane == 0 ? err : lane == 1 ? e_km1 : e_kp1;
      const double my_c = lane == 0 ? 1.2 : lane == 1 ? 1.3 : 1.4;
      const double my_n = lane == 0 ? k + 1.0 : lane == 1 ? (double)k : k + 2.0;
      double temp = 0.;
      if (my_e > 0.) temp = my_c * root_n(my_e / rtol, my_n);
      const double my_h = (temp > 0.1) ? absh / temp : 10 * absh;
      double hopt = __shfl_sync(PT_FULL, my_h, 0);
      const double hkm1 = __shfl_sync(PT_FULL, my_h, 1), hkp1 = __shfl_sync(PT_FULL, my_h, 2);
      int kopt = k;
      if (k > 1 && hkm1 > hopt) { hopt = hkm1; kopt = k - 1; }
      if (k < maxk && hkp1 > hopt) { hopt = hkp1; kopt = k + 1; }
      if (hopt > absh) {
        absh = hopt;
        if (k != kopt) k = kopt;
      }
    }
    t = tnew;
    y = pred + dk1;
    Jcurrent = false;
    PROF_END(PF_CONTROL);
  }
  // final state and a last RHS call so that the by-products are current (evolver_ndf15.cpp:653-662)
  y = pred + dk1;
  if (act) Ys[li] = y;
  __syncwarp();
  env_at(P, tnew, true, 0);
  rhs_apply(P, V_Y, V_F0, 0);
  if (lane == 0) M.st.fevals++;
  M.next = next;
  return true;
}


// =============================================================================================
// TAIL: the radiation-streaming interval (RSA on, ncdm fluid on if any): hub-only, neq = 4 + 3 N_ncdm.
// It is the last interval of every mode and, for k >~ 0.1/Mpc, by far the longest one
// (10^4 .. 3x10^5 steps), so it runs in its own kernel (perturb_tail_kernel) whose hot loop is small
// enough for the instruction cache and keeps EVERYTHING in registers: lane i owns equation i, the
// state is exchanged with warp shuffles, the Newton matrix inverse is one row per lane.
// State order (make_layout with rsa_on): delta_b, theta_b, delta_cdm, [delta, theta, shear](ncdm s), eta.
// =============================================================================================
template <int N>
__device__ __forceinline__ double pickN(const double (&d)[N], int j) {
  double v = d[0];
#pragma unroll
  for (int q = 1; q < N; q++) v = (j == q) ? d[q] : v;
  return v;
}

// time-dependent scalars of the RSA right-hand side, loaded once per step (uniform registers)
struct RsaEnv {
  double a2, aH, rho_b, rho_cdm, rho_g, rho_ur, dkappa, ddkappa, cb2, R, inv_half_aH;
};
__device__ __forceinline__ void rsa_env_load(const PtParams& P, const Mode& M, RsaEnv& E) {
  const double* pvb = s_pvb(P);
  const double* pvt = s_pvt(P);
  E.a2 = M.e.a * M.e.a; E.aH = M.e.a * M.e.H;
  E.rho_b = pvb[P.irho_b]; E.rho_cdm = pvb[P.irho_cdm]; E.rho_g = pvb[P.irho_g];
  E.rho_ur = P.has_ur ? pvb[P.irho_ur] : 0.;
  E.dkappa = pvt[P.idkappa]; E.ddkappa = pvt[P.iddkappa]; E.cb2 = pvt[P.icb2];
  E.R = M.e.drv[DRV_R]; E.inv_half_aH = M.e.drv[DRV_INV_HALF_AH];
}

// f(tau, y) for the RSA interval: every lane evaluates all equations from the uniform y[] and returns
// the derivative of ITS equation (perturb_total_stress_energy / perturb_einstein /
// perturb_rsa_delta_and_theta / perturb_derivs_member restricted to this approximation set)
template <int N>
__device__ __forceinline__ double rhs_rsa(const PtParams& P, const Mode& M, const RsaEnv& E, const double (&y)[N],
                                          int lane, int n_ncdm, double k2, double ik2) {
  const double delta_b = y[0], theta_b = y[1], delta_cdm = y[2];
  const double eta = pickN<N>(y, 3 + 3 * n_ncdm);
  double delta_rho = E.rho_b * delta_b + E.rho_cdm * delta_cdm;
  double rpt = E.rho_b * theta_b;
  double rps = 0.;
#pragma unroll
  for (int s = 0; s < PT_MAX_NCDM; s++) {
    if (3 + 3 * s + 2 < N && s < n_ncdm) {
      const double* nf = M.nf[s];
      delta_rho += nf[0] * y[3 + 3 * s];
      rpt += nf[1] * y[4 + 3 * s];
      rps += nf[1] * y[5 + 3 * s];
    }
  }
  const double aH = E.aH;
  const double h_prime = (k2 * eta + 1.5 * E.a2 * delta_rho) * E.inv_half_aH;
  double rsa_delta_g = 0., rsa_theta_g = 0.;
  if (P.rsa_method != CLPP_RSA_NULL) {
    rsa_delta_g = 4. * ik2 * (aH * h_prime - k2 * eta);
    rsa_theta_g = -0.5 * h_prime;
  }
  const double rsa_delta_ur = rsa_delta_g, rsa_theta_ur = rsa_theta_g;  // before the reionisation correction
  if (P.rsa_method == CLPP_RSA_MD_WITH_REIO) {
    rsa_delta_g += -4. * ik2 * E.dkappa * (theta_b + 0.5 * h_prime);
    rsa_theta_g += 3. * ik2 * (E.ddkappa * (theta_b + 0.5 * h_prime) +
                               E.dkappa * (-aH * theta_b + E.cb2 * k2 * delta_b - aH * h_prime + k2 * eta));
  }
  rpt += 4. / 3. * E.rho_g * rsa_theta_g;
  if (P.has_ur) rpt += 4. / 3. * E.rho_ur * rsa_theta_ur;
  (void)rsa_delta_ur;
  const double eta_prime = (1.5 * E.a2 * rpt) * ik2;
  const double alpha = (h_prime + 6. * eta_prime) * 0.5 * ik2;
  const double metric_continuity = 0.5 * h_prime;
  const double metric_shear = k2 * alpha;
  (void)rps;
  double d = 0.;
  if (lane == 0) d = -(theta_b + metric_continuity);
  else if (lane == 1) d = -aH * theta_b + k2 * E.cb2 * delta_b + E.R * E.dkappa * (rsa_theta_g - theta_b);
  else if (lane == 2) d = -metric_continuity;
  else if (lane == 3 + 3 * n_ncdm) d = eta_prime;
#pragma unroll
  for (int s = 0; s < PT_MAX_NCDM; s++) {
    if (3 + 3 * s + 2 < N && s < n_ncdm) {
      const double* nf = M.nf[s];  // [2] w, [4] ca2, [5] ceff2/(1+w), [6] 8/3 cvis2/(1+w), [7] shear damping rate
      const double y0 = y[3 + 3 * s], y1 = y[4 + 3 * s], y2 = y[5 + 3 * s];
      const double w_n = nf[2], ca2 = nf[4];
      const double ms = (P.ncdmfa_method == CLPP_NCDMFA_CLASS) ? metric_continuity : metric_shear;
      if (lane == 3 + 3 * s) d = -(1.0 + w_n) * (y1 + metric_continuity) - 3.0 * aH * (ca2 - w_n) * y0;
      else if (lane == 4 + 3 * s) d = -aH * (1.0 - 3.0 * ca2) * y1 + nf[5] * k2 * y0 - k2 * y2;
      else if (lane == 5 + 3 * s) d = -nf[7] * y2 + nf[6] * (y1 + ms);
    }
  }
  return d;
}

// A = I - c J from the row-major Jacobian Js[n][N] in shared memory, then gj_rows
template <int N>
__device__ __forceinline__ void factor_rows(const double* __restrict__ Js, int n, double c, int lane, double (&G)[N], int& perm) {
  const int li = lane < n ? lane : 0;
#pragma unroll
  for (int q = 0; q < N; q++) G[q] = ((q == lane) ? 1.0 : 0.0) - ((q < n && lane < n) ? c * Js[li * N + q] : 0.0);
  gj_rows<N>(G, n, lane, perm);
}

template <int N>
__device__ __noinline__ bool ndf15_rsa(const PtParams& P, double t0, double tfinal) {
  Mode& M = MODE(P);
  PROF_DECL;
  const double abstol = 1e-15, eps = 1e-16, threshold = abstol;
  const int maxit = 4, maxk = 5;
  const double rtol = P.rtol;
  const int n = M.L.neq, lane = PT_LANE, n_ncdm = P.has_ncdm ? P.N_ncdm : 0;
  const bool act = lane < n;
  const int li = act ? lane : 0;
  const double k2 = M.k2, ik2 = M.inv_k2;
  double* Js = s_sinv(P);  // Jacobian, row-major [n][N]
  double* Ys = s_vec(P, V_Y);
  const double* t_vec = M.C->tau;
  const int tres = M.C->tau_size;
  int next = M.next;
  while (next < tres && __ldg(t_vec + next) < t0) next++;
  double tnext = (next < tres) ? __ldg(t_vec + next) : 1e300;
  RsaEnv E;
  double yv[N];

  auto gather = [&](double mine) {
#pragma unroll
    for (int q = 0; q < N; q++) yv[q] = __shfl_sync(PT_FULL, mine, q);
  };
  auto jacobian_rsa = [&]() {  // J columns = f(e_j) (linear, homogeneous system); environment must be current
#pragma unroll 1
    for (int j = 0; j < n; j++) {
#pragma unroll
      for (int q = 0; q < N; q++) yv[q] = (q == j) ? 1.0 : 0.0;
      const double d = rhs_rsa<N>(P, M, E, yv, lane, n_ncdm, k2, ik2);
      if (act) Js[li * N + j] = d;
    }
    __syncwarp();
    if (lane == 0) { M.st.jacobians++; M.st.fevals += n; }
  };

  double y = act ? Ys[li] : 0.;
  double dif[7];
#pragma unroll
  for (int j = 0; j < 7; j++) dif[j] = 0.;
  const double htspan = fabs(tfinal - t0);
  double t = t0, tnew = t0;
  env_at(P, t0, true, 0);
  rsa_env_load(P, M, E);
  gather(y);
  const double f0i = rhs_rsa<N>(P, M, E, yv, lane, n_ncdm, k2, ik2);
  if (lane == 0) M.st.fevals++;
  const double hmax = (tfinal - t0) / 10.0;
  jacobian_rsa();
  bool Jcurrent = true;
  double hmin = 16.0 * eps * fabs(t);
  const double wt0 = fmax(fabs(y), threshold);
  double rh = wmax(act ? 1.25 / sqrt(rtol) * fabs(f0i / wt0) : 0.);
  double absh = fmin(hmax, htspan);
  if (absh * rh > 1.0) absh = 1.0 / rh;
  absh = fmax(absh, hmin);
  double h = absh;
  {
    gather(f0i);  // J*f0 = f(t0, f0)
    const double jf = rhs_rsa<N>(P, M, E, yv, lane, n_ncdm, k2, ik2);
    const double tdel = (t + fmin(sqrt(eps) * fmax(fabs(t), fabs(t + h)), absh)) - t;
    env_at(P, t + tdel, true, 0);
    rsa_env_load(P, M, E);
    gather(y);
    const double fdel = rhs_rsa<N>(P, M, E, yv, lane, n_ncdm, k2, ik2);
    if (lane == 0) M.st.fevals += 2;
    const double s = act ? jf + (fdel - f0i) / tdel : 0.;
    rh = wmax(1.25 * sqrt(0.5 * fabs(s / wt0) / rtol));
    absh = fmin(hmax, htspan);
    if (absh * rh > 1.0) absh = 1.0 / rh;
    absh = fmax(absh, hmin);
    h = absh;
  }
  int k = 1, klast = k;
  double abshlast = absh;
  dif[0] = h * f0i;
  double hinvGak = h * c_invGa[k - 1];
  int nconhk = 0;
  double G[N];
  int perm;
  factor_rows<N>(Js, n, hinvGak, lane, G, perm);
  if (lane == 0) M.st.factorizations++;
  bool havrate = false;
  bool done = false, at_hmin = false;
  double rate = 0., oldnrm = 0., err = 0.;
  double pred = 0., psi = 0., dk1 = 0., iw = 0., f_last = f0i;

  // difference-array rescaling in registers (adjust_stepsize)
  auto rescale = [&](double r) {
    double* RU = s_hubtmp(P);
    if (lane < 25) {
      const int ii = lane / 5, jj = lane % 5;
      double s = 0.;
#pragma unroll 1
      for (int kk = 0; kk < 5; kk++) {
        double Rv = 1.;
        for (int mm = 1; mm <= ii + 1; mm++) Rv *= ((mm - 1) - (kk + 1) * r) * c_invint[mm];
        s += Rv * c_U[kk][jj];
      }
      RU[lane] = s;
    }
    __syncwarp();
    double nd[5];
#pragma unroll
    for (int jj = 0; jj < 5; jj++) {
      double s = 0.;
#pragma unroll
      for (int kk = 0; kk < 5; kk++) s += ((kk < k) ? dif[kk] : 0.) * RU[kk * 5 + jj];
      nd[jj] = s;
    }
#pragma unroll
    for (int jj = 0; jj < 5; jj++)
      if (jj < k) dif[jj] = nd[jj];
    __syncwarp();
  };

  int sync_ctr = 0;
  while (!done) {
    PT_COHORT_SYNC(P);
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
      PROF_BEGIN();
      rescale(absh / abshlast);
      PROF_END(PF_ADJUST);
      hinvGak = h * c_invGa[k - 1];
      nconhk = 0;
      PROF_BEGIN();
      factor_rows<N>(Js, n, hinvGak, lane, G, perm);
      if (lane == 0) M.st.factorizations++;
      PROF_END(PF_FACTOR);
      havrate = false;
    }
    bool nofailed = true;
    for (;;) {  // loop for advancing one step
      bool gotynew = false;
      while (!gotynew) {
        tnew = t + h;
        if (done) tnew = tfinal;
        h = tnew - t;
        PROF_BEGIN();
        const double invGak = c_invGa[k - 1];
        psi = 0.;
        pred = y;
#pragma unroll
        for (int j = 0; j < 5; j++) {
          if (j < k) {
            psi += dif[j] * c_G[j] * invGak;
            pred += dif[j];
          }
        }
        dk1 = 0.;
        iw = 1.0 / fmax(fmax(fabs(pred), fabs(y)), threshold);
        const double minnrm = wmax(act ? 100 * eps * fabs(pred * iw) : 0.);
        PROF_END(PF_PREDICT);
        PROF_BEGIN();
        env_at(P, tnew, true, 0);
        rsa_env_load(P, M, E);
        PROF_END(PF_ENV);
        bool tooslow = false;
#pragma unroll 1
        for (int iter = 1; iter <= maxit; iter++) {
          PROF_BEGIN();
          gather(pred + dk1);
          f_last = rhs_rsa<N>(P, M, E, yv, lane, n_ncdm, k2, ik2);
          PROF_END(PF_RHS);
          PROF_BEGIN();
          // residual r, then x = A^-1 r = G (P r): lane j fetches r[perm[j]], every lane dots its row of G
          const double r = act ? hinvGak * f_last - (psi + dk1) : 0.;
          const double rp = __shfl_sync(PT_FULL, r, perm);
          double x0 = 0., x1 = 0.;
#pragma unroll
          for (int q = 0; q < N; q += 2) {
            x0 += G[q] * __shfl_sync(PT_FULL, rp, q);
            x1 += G[q + 1] * __shfl_sync(PT_FULL, rp, q + 1);
          }
          const double d = act ? x0 + x1 : 0.;
          PROF_END(PF_SOLVE);
          PROF_BEGIN();
          const double newnrm = wmax(fabs(d * iw));
          dk1 += d;
          if (lane == 0) { M.st.fevals++; M.st.solves++; }
          PROF_END(PF_UPDATE);
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
              double rpw = rate;  // rate^(maxit-iter)
              for (int q = 1; q < maxit - iter; q++) rpw *= rate;
              if (0.5 * rtol < errit * rpw) { tooslow = true; break; }
            }
          }
          oldnrm = newnrm;
        }
        if (tooslow) {
          if (lane == 0) M.st.failed++;
          if (!Jcurrent) {
            env_at(P, t, true, 0);
            rsa_env_load(P, M, E);
            jacobian_rsa();
            if (lane == 0) M.st.fevals++;  // the reference re-evaluates f(t, y) with a new Jacobian
            Jcurrent = true;
          } else if (absh <= hmin) {
            M.status = 2;  // step size too small
            return false;
          } else {
            abshlast = absh;
            absh = fmax(0.3 * absh, hmin);
            h = absh;
            done = false;
            rescale(absh / abshlast);
            hinvGak = h * c_invGa[k - 1];
            nconhk = 0;
          }
          factor_rows<N>(Js, n, hinvGak, lane, G, perm);
          if (lane == 0) M.st.factorizations++;
          havrate = false;
        }
      }
      // error estimate
      err = wmax(fabs(dk1 * iw)) * c_erconst[k - 1];
      if (err > rtol) {
        if (lane == 0) M.st.failed++;
        if (absh <= hmin) {
          M.status = 2;
          return false;
        }
        abshlast = absh;
        if (nofailed) {
          nofailed = false;
          double hopt = absh * fmax(0.1, 0.833 * root_n(rtol / err, k + 1.0));
          if (k > 1) {
            const double errkm1 = wmax(fabs((pick7(dif, k - 1) + dk1) * iw)) * c_erconst[k - 2];
            const double hkm1 = absh * fmax(0.1, 0.769 * root_n(rtol / errkm1, (double)k));
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
        rescale(absh / abshlast);
        hinvGak = h * c_invGa[k - 1];
        nconhk = 0;
        factor_rows<N>(Js, n, hinvGak, lane, G, perm);
        if (lane == 0) M.st.factorizations++;
        havrate = false;
      } else {
        break;
      }
    }
    if (lane == 0) M.st.steps++;
    PROF_BEGIN();
    {
      const double difk_old = pick7(dif, k);
#pragma unroll
      for (int j = 6; j >= 0; j--) {
        if (j == k + 1) dif[j] = dk1 - difk_old;
        else if (j == k) dif[j] = dk1;
        else if (j < k) dif[j] += dif[j + 1];
      }
    }
    PROF_END(PF_DIFUPD);
    PROF_BEGIN();
    // ---- output at the sample times passed by this step (through the generic source routine)
    while ((next < tres) && ((tnew - tnext) >= 0.0)) {
      double* yo = s_vec(P, V_TMP);
      double* dyo = s_vec(P, V_YPI);
      if (tnew == tnext) {
        if (act) { yo[li] = pred + dk1; dyo[li] = f_last; }
      } else {
        const double s = (tnext - tnew) / h;
        double a1 = 0, a2 = 0;
        double prod = 1.0, sumfrac = 0., fact = 1.0;
#pragma unroll
        for (int j = 0; j < 5; j++) {
          if (j < k) {
            prod *= (s + j);
            fact *= (j + 1);
            sumfrac += 1.0 / (s + j);
            a1 += prod / fact * dif[j];
            a2 += prod * sumfrac / (h * fact) * dif[j];
          }
        }
        if (act) { yo[li] = (pred + dk1) + a1; dyo[li] = a2; }
      }
      __syncwarp();
      write_sources(P, tnext, V_TMP, V_YPI, next);
      next++;
      tnext = (next < tres) ? __ldg(t_vec + next) : 1e300;
    }
    PROF_END(PF_OUTPUT);
    if (done) break;
    PROF_BEGIN();
    klast = k;
    abshlast = absh;
    nconhk = min(nconhk + 1, maxk + 2);
    if (nconhk >= k + 2) {
      double e_km1 = 0., e_kp1 = 0.;
      if (k > 1) e_km1 = wmax(fabs(pick7(dif, k - 1) * iw)) * c_erconst[k - 2];
      if (k < maxk) e_kp1 = wmax(fabs(pick7(dif, k + 1) * iw)) * c_erconst[k];
      const double my_e = lane == 0 ? err : lane == 1 ? e_km1 : e_kp1;
      const double my_c = lane == 0 ? 1.2 : lane == 1 ? 1.3 : 1.4;
      const double my_n = lane == 0 ? k + 1.0 : lane == 1 ? (double)k : k + 2.0;
      double temp = 0.;
      if (my_e > 0.) temp = my_c * root_n(my_e / rtol, my_n);
      const double my_h = (temp > 0.1) ? absh / temp : 10 * absh;
      double hopt = __shfl_sync(PT_FULL, my_h, 0);
      const double hkm1 = __shfl_sync(PT_FULL, my_h, 1), hkp1 = __shfl_sync(PT_FULL, my_h, 2);
      int kopt = k;
      if (k > 1 && hkm1 > hopt) { hopt = hkm1; kopt = k - 1; }
      if (k < maxk && hkp1 > hopt) { hopt = hkp1; kopt = k + 1; }
      if (hopt > absh) {
        absh = hopt;
        if (k != kopt) k = kopt;
      }
    }
    t = tnew;
    y = pred + dk1;
    Jcurrent = false;
    PROF_END(PF_CONTROL);
  }
  y = pred + dk1;
  if (act) Ys[li] = y;
  __syncwarp();
  if (lane == 0) M.st.fevals++;  // the reference's final RHS call (nothing follows the last interval)
  M.next = next;
  return true;
}


// ---------------------------------------------------------------------------------------------
// evolver = rk: Cash-Karp Runge-Kutta 4(5) with time-scale limited stepping (the reference's non-default
// evolver: tools/evolver_rkck.c:3-178, tools/dei_rkck.c:79-238, perturb_timescale_member
// perturbations_module.cpp:5691-5826).  Explicit and therefore slow in the stiff phases; provided because it is
// on the path, with the same driver logic: every sample time of the source grid is hit exactly.
__device__ __noinline__ bool rk_interval(const PtParams& P, double t0, double tfinal) {
  Mode& M = MODE(P);
  const int n = M.L.neq, lane = PT_LANE;
  const double eps_tol = P.rtol;
  const double SAFETY = 0.9, PGROW = -0.2, PSHRNK = -0.25, ERRCON = 1.89e-4, TINY = 1.0e-30;
  const int MAXSTP = 100000;
  double *ys = s_vec(P, V_Y), *y = s_vec(P, V_YNEW), *dydx = s_vec(P, V_F0), *ytemp = s_vec(P, V_PRED),
         *ak2 = s_vec(P, V_PSI), *ak3 = s_vec(P, V_DIFKP1), *ak4 = s_vec(P, V_DEL), *ak5 = s_vec(P, V_INVWT);
  double *ak6 = s_vec(P, V_DIF0), *yerr = s_vec(P, V_DIF0 + 1), *yscal = s_vec(P, V_DIF0 + 2);
  const double* t_vec = M.C->tau;
  const int tres = M.C->tau_size;
  int next = M.next;
  while (next < tres && __ldg(t_vec + next) < t0) next++;
  double x1 = t0;
  while ((x1 < tfinal) && (next < tres)) {
    // perturb_timescale: min(tau_h, tau_k [unless rsa without ncdm], tau_c [when tight coupling is off])
    env_at(P, x1, true, 0);
    double timescale = 1. / (M.e.a * M.e.H);
    if (!M.ap.rsa_on || P.has_ncdm) timescale = fmin(1. / M.k, timescale);
    if (M.ap.tca_off) {
      const double dkappa = s_pvt(P)[P.idkappa];
      if (dkappa != 0.) timescale = fmin(1. / dkappa, timescale);
    }
    const double timestep = P.rk_stepsize * timescale;
    if (fabs(timestep / x1) < P.hmin_allowed) { M.status = 2; return false; }
    const double tnext = __ldg(t_vec + next);
    double x2;
    bool call_output = false;
    if (x1 + 2. * timestep < tnext) x2 = x1 + timestep;
    else { x2 = tnext; call_output = true; }
    if (x2 > tfinal) { x2 = tfinal; call_output = false; }
    // ---- generic_integrator(x1 -> x2) with adaptive Cash-Karp steps
    {
      double x = x1, h = x2 - x1;
      const double hmin = x1 * P.hmin_allowed;
#pragma unroll 1
      for (int i = lane; i < n; i += 32) y[i] = ys[i];
      __syncwarp();
      bool reached = false;
#pragma unroll 1
      for (int nstp = 1; nstp <= MAXSTP; nstp++) {
        env_at(P, x, true, 0);
        rhs_apply(P, V_YNEW, V_F0, 0);
        if (lane == 0) M.st.fevals++;
#pragma unroll 1
        for (int i = lane; i < n; i += 32) yscal[i] = fabs(y[i]) + fabs(dydx[i] * h) + TINY;
        if ((x + h - x2) * (x + h - x1) > 0.0) h = x2 - x;
        // rkqs: shrink the step until the embedded error estimate passes
        double errmax, hnext;
        for (;;) {
          // rkck
#pragma unroll 1
          for (int i = lane; i < n; i += 32) ytemp[i] = y[i] + 0.2 * h * dydx[i];
          __syncwarp();
          env_at(P, x + 0.2 * h, true, 0); rhs_apply(P, V_PRED, V_PSI, 0);
#pragma unroll 1
          for (int i = lane; i < n; i += 32) ytemp[i] = y[i] + h * (3.0 / 40.0 * dydx[i] + 9.0 / 40.0 * ak2[i]);
          __syncwarp();
          env_at(P, x + 0.3 * h, true, 0); rhs_apply(P, V_PRED, V_DIFKP1, 0);
#pragma unroll 1
          for (int i = lane; i < n; i += 32) ytemp[i] = y[i] + h * (0.3 * dydx[i] - 0.9 * ak2[i] + 1.2 * ak3[i]);
          __syncwarp();
          env_at(P, x + 0.6 * h, true, 0); rhs_apply(P, V_PRED, V_DEL, 0);
#pragma unroll 1
          for (int i = lane; i < n; i += 32)
            ytemp[i] = y[i] + h * (-11.0 / 54.0 * dydx[i] + 2.5 * ak2[i] - 70.0 / 27.0 * ak3[i] + 35.0 / 27.0 * ak4[i]);
          __syncwarp();
          env_at(P, x + 1.0 * h, true, 0); rhs_apply(P, V_PRED, V_INVWT, 0);
#pragma unroll 1
          for (int i = lane; i < n; i += 32)
            ytemp[i] = y[i] + h * (1631.0 / 55296.0 * dydx[i] + 175.0 / 512.0 * ak2[i] + 575.0 / 13824.0 * ak3[i] +
                                   44275.0 / 110592.0 * ak4[i] + 253.0 / 4096.0 * ak5[i]);
          __syncwarp();
          env_at(P, x + 0.875 * h, true, 0); rhs_apply(P, V_PRED, V_DIF0, 0);
          if (lane == 0) M.st.fevals += 5;
          double em = 0.0;
#pragma unroll 1
          for (int i = lane; i < n; i += 32) {
            ytemp[i] = y[i] + h * (37.0 / 378.0 * dydx[i] + 250.0 / 621.0 * ak3[i] + 125.0 / 594.0 * ak4[i] + 512.0 / 1771.0 * ak6[i]);
            const double ye = h * ((37.0 / 378.0 - 2825.0 / 27648.) * dydx[i] + (250.0 / 621.0 - 18575.0 / 48384.0) * ak3[i] +
                                   (125.0 / 594.0 - 13525.0 / 55296.0) * ak4[i] - 277.00 / 14336.0 * ak5[i] +
                                   (512.0 / 1771.0 - 0.25) * ak6[i]);
            yerr[i] = ye;
            em = fmax(em, fabs(ye / yscal[i]));
          }
          errmax = wmax(em) / eps_tol;
          __syncwarp();
          if (errmax <= 1.0) break;
          if (lane == 0) M.st.failed++;
          const double htemp = SAFETY * h * pow(errmax, PSHRNK);
          h = (h >= 0.0 ? fmax(htemp, 0.1 * h) : fmin(htemp, 0.1 * h));
          if (x + h == x) { M.status = 2; return false; }  // stepsize underflow
        }
        if (errmax > ERRCON) hnext = SAFETY * h * pow(errmax, PGROW);
        else hnext = 5.0 * h;
        x += h;
#pragma unroll 1
        for (int i = lane; i < n; i += 32) y[i] = ytemp[i];
        __syncwarp();
        if (lane == 0) M.st.steps++;
        if ((x - x2) * (x2 - x1) >= 0.0) { reached = true; break; }
        if (fabs(hnext / x1) <= hmin) { M.status = 2; return false; }
        h = hnext;
      }
      if (!reached) { M.status = 9; return false; }  // too many steps within one sub-interval
#pragma unroll 1
      for (int i = lane; i < n; i += 32) ys[i] = y[i];
      __syncwarp();
    }
    if (call_output) {
      env_at(P, x2, true, 0);
      rhs_apply(P, V_Y, V_F0, 0);
      if (!M.ap.tca_off) M.tca_shear_last = M.m.tca_shear_g;
      if (lane == 0) M.st.fevals++;
      __syncwarp();
      write_sources(P, x2, V_Y, V_F0, next);
      next++;
    }
    x1 = x2;
  }
  // last call so that the environment and the TCA/RSA by-products are current at the end of the interval
  env_at(P, x1, true, 0);
  rhs_apply(P, V_Y, V_F0, 0);
  if (!M.ap.tca_off) M.tca_shear_last = M.m.tca_shear_g;
  if (lane == 0) M.st.fevals++;
  __syncwarp();
  M.next = next;
  return true;
}

// ---------------------------------------------------------------------------------------------
// perturb_initial_conditions: adiabatic mode, synchronous gauge, flat space
__device__ __noinline__ void initial_conditions(const PtParams& P, double tau) {
  Mode& M = MODE(P);
  env_at(P, tau, false, 0);
  const Env& e = M.e;
  const Layout& L = M.L;
  const int lane = PT_LANE;
  const double k = M.k, a = e.a;
  double* y = s_vec(P, V_Y);
  const double* pvb0 = s_pvb(P);
  const double e_rho_g = pvb0[P.irho_g], e_rho_b = pvb0[P.irho_b], e_rho_cdm = pvb0[P.irho_cdm];
  const double e_rho_ur = P.has_ur ? pvb0[P.irho_ur] : 0.;
  double rho_r = e_rho_g, rho_m = e_rho_b + e_rho_cdm, rho_nu = 0.;
  if (P.has_ur) { rho_r += e_rho_ur; rho_nu += e_rho_ur; }
  for (int s = 0; s < P.N_ncdm; s++) { rho_r += s_pvb(P)[P.irho_ncdm1 + s]; rho_nu += s_pvb(P)[P.irho_ncdm1 + s]; }
  const double fracnu = rho_nu / rho_r;
  const double fracb = e_rho_b / rho_m;
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
#pragma unroll 1
  for (int i = lane; i < L.neq; i += 32) y[i] = 0.;
  __syncwarp();
  if (lane == 0) {
    y[L.delta_g] = delta_g;
    y[L.theta_g] = theta_g;
    y[L.delta_b] = 3. / 4. * delta_g;
    y[L.theta_b] = theta_g;
    y[L.delta_cdm] = 3. / 4. * delta_g;
    y[L.eta] = eta;
    if (P.has_ur) {
      y[L.delta_ur] = delta_ur;
      y[L.theta_ur] = theta_ur;
      y[L.shear_ur] = shear_ur;
      y[L.l3_ur] = l3_ur;
    }
  }
  if (P.has_ncdm) {
    const int stride = L.l_max_ncdm + 1;
    for (int s = 0; s < P.N_ncdm; s++) {
      const double Ms = M.C->ncdm_M[s];
      const int nq = P.ncdm_q_size[s], off = L.psi0_ncdm1 + P.ncdm_q_off[s] * stride;
#pragma unroll 1
      for (int iq = lane; iq < nq; iq += 32) {
        const int idx = off + iq * stride;
        const double q = M.C->ncdm_q[P.ncdm_q_off[s] + iq];
        const double dlnf0 = M.C->ncdm_dlnf0[P.ncdm_q_off[s] + iq];
        const double eps = sqrt(q * q + a * a * Ms * Ms);
        y[idx + 0] = -0.25 * delta_ur * dlnf0;
        y[idx + 1] = -eps / 3. / q / k * theta_ur * dlnf0;
        y[idx + 2] = -0.5 * shear_ur * dlnf0;
        y[idx + 3] = -0.25 * l3_ur * dlnf0;
      }
    }
  }
  __syncwarp();
}

// perturb_vector_init (switching part): move the state from the old layout (in V_Y) to the new one
__device__ __noinline__ void remap_state(const PtParams& P) {
  Mode& M = MODE(P);
  const Layout& Lo = M.Lprev;
  const Layout& Ln = M.L;
  const Approx& apo = M.apprev;
  const Approx& apn = M.ap;
  const int lane = PT_LANE;
  double* yo = s_vec(P, V_Y);
  double* yn = s_vec(P, V_YNEW);
  const double k = M.k;
#pragma unroll 1
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
      const double sg = M.m.tca_shear_g, dk = s_pvt(P)[P.idkappa];
      yn[Ln.shear_g] = sg;
      yn[Ln.l3_g] = 6. / 7. * k / dk * sg;
      yn[Ln.pol0_g] = 2.5 * sg;
      yn[Ln.pol0_g + 1] = k / dk * (5. - 2.) / 6. * sg;
      yn[Ln.pol0_g + 2] = 0.5 * sg;
      yn[Ln.pol0_g + 3] = k / dk * 3. / 14. * sg;
    }
  }
  if (Ln.shear_g >= 0 && Lo.shear_g >= 0) {
#pragma unroll 1
    for (int l = 2 + lane; l <= P.l_max_g; l += 32) yn[Ln.delta_g + l] = yo[Lo.delta_g + l];
#pragma unroll 1
    for (int l = lane; l <= P.l_max_pol_g; l += 32) yn[Ln.pol0_g + l] = yo[Lo.pol0_g + l];
  }
  if (Ln.l3_ur >= 0 && Lo.l3_ur >= 0)
#pragma unroll 1
    for (int l = 3 + lane; l <= P.l_max_ur; l += 32) yn[Ln.delta_ur + l] = yo[Lo.delta_ur + l];
  if (P.has_ncdm) {
    if (apn.ncdmfa_on == apo.ncdmfa_on) {
      const int tot = Ln.eta - Ln.psi0_ncdm1;
#pragma unroll 1
      for (int i = lane; i < tot; i += 32) yn[Ln.psi0_ncdm1 + i] = yo[Lo.psi0_ncdm1 + i];
    } else {
      // ncdm fluid approximation switched on: integrate the momentum hierarchy (:4478-4518)
      const double a = M.e.a;
      const double a_rel = M.C->a_today / a, a_rel4 = (a_rel * a_rel) * (a_rel * a_rel);
      const int stride = Lo.l_max_ncdm + 1;
      for (int s = 0; s < P.N_ncdm; s++) {
        const double rho_n = s_pvb(P)[P.irho_ncdm1 + s], p_n = s_pvb(P)[P.ip_ncdm1 + s];
        const double factor = M.C->ncdm_factor[s] * a_rel4;
        const double Ms = M.C->ncdm_M[s];
        const int off = Lo.psi0_ncdm1 + P.ncdm_q_off[s] * stride;
        double d = 0., th = 0., sh = 0.;
#pragma unroll 1
        for (int iq = lane; iq < P.ncdm_q_size[s]; iq += 32) {
          const int idx = off + iq * stride;
          const double q = M.C->ncdm_q[P.ncdm_q_off[s] + iq], w0 = M.C->ncdm_w[P.ncdm_q_off[s] + iq];
          const double q2 = q * q, eps = sqrt(q2 + a * a * Ms * Ms);
          d += w0 * q2 * eps * yo[idx];
          th += w0 * q2 * q * yo[idx + 1];
          sh += w0 * q2 * q2 / eps * yo[idx + 2];
        }
        d = wsum(d) * factor / rho_n;
        th = wsum(th) * k * factor / (rho_n + p_n);
        sh = wsum(sh) * 2. / 3. * factor / (rho_n + p_n);
        if (lane == 0) {
          yn[Ln.psi0_ncdm1 + 3 * s] = d;
          yn[Ln.psi0_ncdm1 + 3 * s + 1] = th;
          yn[Ln.psi0_ncdm1 + 3 * s + 2] = sh;
        }
      }
    }
  }
  __syncwarp();
#pragma unroll 1
  for (int i = lane; i < Ln.neq; i += 32) yo[i] = yn[i];
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// common set-up of the shared-memory mode state
__device__ __forceinline__ void mode_init(const PtParams& P, const PtCosmo* C, int ik) {
  Mode& M = MODE(P);
  const int lane = PT_LANE;
  if (lane == 0) {
    M.Jhh = P.hub_jac + (size_t)PT_SLOT(P) * P.scr_stride;
    M.gJd = M.Jhh + (size_t)P.nh_max * P.nh_max; M.gJu = M.gJd + P.np; M.gdif = M.gJu + P.np;
    M.C = C;
    M.ik = ik;
    M.k = C->k[ik];
    M.k2 = M.k * M.k;
    M.inv_k = 1.0 / M.k;
    M.inv_k2 = 1.0 / M.k2;
    M.bg_tau = C->bg_tau; M.bg_y = C->bg_y; M.bg_dd = C->bg_dd;
    M.th_z = C->th_z; M.th_y = C->th_y; M.th_dd = C->th_dd;
    M.bt_size = C->bt_size; M.tt_size = C->tt_size;
    M.z_last = C->th_z[C->tt_size - 1]; M.th_lin = C->th_linear_below_z; M.a_today = C->a_today;
    for (int q = 0; q < 2; q++) {
      M.bgc[q].x0 = 1.; M.bgc[q].x1 = 0.; M.bgc[q].cur = 0;
      M.thc[q].x0 = 1.; M.thc[q].x1 = 0.; M.thc[q].cur = C->tt_size - 2;
    }
    M.need_nw = 0;
    M.nh = 0; M.nch = 0;
    M.status = 0;
    M.next = 0;
    M.tca_shear_last = 0.;
    M.st = Stat{};
    M.m = Metric{};
    M.ap.tca_off = M.ap.rsa_on = M.ap.ufa_on = M.ap.ncdmfa_on = 0;
    for (int q = 0; q < PF_COUNT; q++) M.prof[q] = 0;
  }
#pragma unroll 1
  for (int l = lane; l < P.n_i2l1; l += 32) s_i2l1(P)[l] = 1.0 / (2.0 * l + 1.0);
  __syncwarp();
}

// end of a mode: zero-fill the samples that were not reached (failure only) and publish the counters
__device__ __forceinline__ void mode_finish(const PtParams& P, const PtCosmo* C, int ik, int n_int, int status, double tau_ini) {
  Mode& M = MODE(P);
  if (PT_LANE == 0) {
    const int tau_size = C->tau_size;
    const size_t stride_tp = (size_t)C->k_size * tau_size;
    double* out = C->sources + (size_t)ik * tau_size;
    const int tps[7] = {P.tp_t0, P.tp_t1, P.tp_t2, P.tp_p, P.tp_delta_m, P.tp_delta_cb, P.tp_phi_plus_psi};
    for (int it = M.next; it < tau_size; it++)
      for (int j = 0; j < 7; j++)
        if (tps[j] >= 0) out[tps[j] * stride_tp + it] = 0.;
    clpp_kstat* ks = C->kstat + ik;
    ks->steps = M.st.steps; ks->failed = M.st.failed; ks->fevals = M.st.fevals; ks->jacobians = M.st.jacobians;
    ks->factorizations = M.st.factorizations; ks->solves = M.st.solves;
    ks->intervals = n_int; ks->status = status; ks->tau_ini = tau_ini;
  }
}

// hand-off record of a mode whose last (radiation-streaming) interval runs in perturb_tail_kernel
enum { TL_VALID = 0, TL_T0, TL_TF, TL_NEXT, TL_IV, TL_NINT, TL_TAU_INI, TL_FLAGS, TL_STAT = TL_FLAGS + 4, TL_Y = TL_STAT + 6,
       TL_STRIDE = TL_Y + 16 };

#ifndef PT_MIN_BLOCKS
#define PT_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(32 * PT_MAX_WPC, PT_GEN_MIN_BLOCKS) perturb_kernel(const __grid_constant__ PtParams P) {
  if (PT_SLOT(P) >= P.n_modes) return;
  Mode& M = MODE(P);
  const int lane = PT_LANE;
  const int2 md = P.modes[PT_SLOT(P)];
  const PtCosmo* C = P.cosmo + md.x;
  mode_init(P, C, md.y);
  const double tau_first = C->tau[0];
  const int tau_size = C->tau_size;

  // ---- start time: bisection on (tau_c/tau_h, tau_h/tau_k, ncdm still relativistic)  (:2592-2635)
  double tau_lower = C->bg_tau[0], tau_upper = tau_first;
  int status = 0;
  {
    env_at(P, tau_lower, false, 0);
    if (M.e.a * M.e.H / s_pvt(P)[P.idkappa] > P.start_small_k_at_tau_c_over_tau_h) status = 3;
    if (M.k / M.e.a / M.e.H > P.start_large_k_at_tau_h_over_tau_k) status = 4;
    for (int s = 0; s < P.N_ncdm; s++)
      if (fabs(s_pvb(P)[P.ip_ncdm1 + s] / s_pvb(P)[P.irho_ncdm1 + s] - 1. / 3.) > P.tol_ncdm_initial_w) status = 5;
  }
  double tau_mid = 0.5 * (tau_lower + tau_upper);
  if (status == 0) {
    while ((tau_upper - tau_lower) / tau_lower > P.tol_tau_approx) {
      env_at(P, tau_mid, false, 0);
      bool early = true;
      for (int s = 0; s < P.N_ncdm; s++)
        if (fabs(s_pvb(P)[P.ip_ncdm1 + s] / s_pvb(P)[P.irho_ncdm1 + s] - 1. / 3.) > P.tol_ncdm_initial_w) early = false;
      if (early) {
        if ((M.e.a * M.e.H / s_pvt(P)[P.idkappa] > P.start_small_k_at_tau_c_over_tau_h) ||
            (M.k / M.e.a / M.e.H > P.start_large_k_at_tau_h_over_tau_k))
          early = false;
      }
      if (early) tau_lower = tau_mid; else tau_upper = tau_mid;
      tau_mid = 0.5 * (tau_lower + tau_upper);
    }
  }
  const double tau_ini = tau_mid;
  const double tau_end = C->tau[tau_size - 1];

  // ---- schedule of approximation switches (:2940-3231); every lane computes the same values
  int n_int = 1;
  if (status == 0) {
    const Approx a_ini = approximations_at(P, tau_ini);
    const Approx a_end = approximations_at(P, tau_end);
    int nsw = 0;
    for (int w = 0; w < 4; w++) {
      const int f0 = approx_flag(a_ini, w), f1 = approx_flag(a_end, w);
      if (f1 < f0) { status = 6; break; }
      if (f1 > f0) {
        double lo = tau_ini, hi = tau_end, mid = 0.5 * (lo + hi);
        while (hi - lo > P.tol_tau_approx) {
          const Approx am = approximations_at(P, mid);
          if (approx_flag(am, w) > f0) hi = mid; else lo = mid;
          mid = 0.5 * (lo + hi);
        }
        M.sw[nsw++] = mid;
      }
    }
    __syncwarp();
    n_int = nsw + 1;
    M.limit[0] = tau_ini;
    for (int i = 1; i < n_int; i++) {
      double nxt = tau_end;
      for (int j = 0; j < nsw; j++)
        if ((M.sw[j] > M.limit[i - 1]) && (M.sw[j] < nxt)) nxt = M.sw[j];
      M.limit[i] = nxt;
    }
    M.limit[n_int] = tau_end;
    M.sched[0] = a_ini;
    for (int i = 1; i < n_int && status == 0; i++) {
      const Approx ai = approximations_at(P, 0.5 * (M.limit[i] + M.limit[i + 1]));
      const Approx ap = M.sched[i - 1];
      M.sched[i] = ai;
      int nchange = 0;
      for (int w = 0; w < 4; w++) {
        if (approx_flag(ai, w) < approx_flag(ap, w)) status = 6;
        if (approx_flag(ai, w) != approx_flag(ap, w)) nchange++;
      }
      if (nchange != 1) status = 7;
    }
    if (a_ini.tca_off || a_ini.rsa_on || a_ini.ufa_on || a_ini.ncdmfa_on) status = 8;
  }
  M.status = status;
  __syncwarp();
  clpp_kstat* ks = C->kstat + md.y;

  // the last interval goes to the tail kernel when it is the hub-only radiation-streaming phase
  bool tail = false;
  if (status == 0 && P.tail != nullptr && n_int >= 2) {
    const Approx al = M.sched[n_int - 1];
    tail = al.rsa_on && (!P.has_ncdm || al.ncdmfa_on) && (4 + 3 * (P.has_ncdm ? P.N_ncdm : 0) <= 16);
  }
  const int n_here = tail ? n_int - 1 : n_int;

  // ---- integrate interval by interval
  for (int iv = 0; iv < n_int && status == 0; iv++) {
    const Approx apn = M.sched[iv];
    __syncwarp();
    M.Lprev = M.L;
    M.apprev = M.ap;
    __syncwarp();
    M.ap = apn;
    make_layout(P, apn, M.L);
    __syncwarp();
    if (iv == 0) initial_conditions(P, M.limit[0]);
    else remap_state(P);
    if (iv >= n_here) break;  // state is laid out for the tail interval: hand it off below
    make_structure(P);
    M.need_nw = P.has_ncdm && !apn.ncdmfa_on;
    __syncwarp();
    const long long c0 = clock64();
    const int s0 = M.st.steps;
    bool ok;
    if (P.evolver == 0) ok = rk_interval(P, M.limit[iv], M.limit[iv + 1]);
    else if (P.force_generic) ok = ndf15(P, M.limit[iv], M.limit[iv + 1]);
    else if (M.nch == 0 && M.L.neq <= 32) ok = ndf15_hub(P, M.limit[iv], M.limit[iv + 1]);  // (the register-only RSA integrator lives in the tail kernel)
    else ok = ndf15(P, M.limit[iv], M.limit[iv + 1]);
    __syncwarp();
    if (lane == 0) {
      ks->iv_neq[iv] = M.L.neq;
      ks->iv_steps[iv] = M.st.steps - s0;
      ks->iv_cycles[iv] = clock64() - c0;
#ifdef PT_PROF
      for (int q = 0; q < PF_COUNT; q++) { ks->prof[iv * 12 + q] = M.prof[q]; M.prof[q] = 0; }
#endif
    }
    __syncwarp();
    status = M.status;
    if (!ok) break;
  }
  if (tail && status == 0) {
    double* T = P.tail + (size_t)PT_SLOT(P) * TL_STRIDE;
    const int n = M.L.neq;
    if (lane < n) T[TL_Y + lane] = s_vec(P, V_Y)[lane];
    if (lane == 0) {
      const Approx al = M.sched[n_int - 1];
      T[TL_VALID] = (M.k < P.tail_lane_kmax) ? 2. : 1.;  // 2: perturb_tail_lane_kernel takes it
      T[TL_T0] = M.limit[n_int - 1]; T[TL_TF] = M.limit[n_int]; T[TL_NEXT] = M.next;
      T[TL_IV] = n_int - 1; T[TL_NINT] = n_int; T[TL_TAU_INI] = tau_ini;
      T[TL_FLAGS] = al.tca_off; T[TL_FLAGS + 1] = al.rsa_on; T[TL_FLAGS + 2] = al.ufa_on; T[TL_FLAGS + 3] = al.ncdmfa_on;
      T[TL_STAT] = M.st.steps; T[TL_STAT + 1] = M.st.failed; T[TL_STAT + 2] = M.st.fevals; T[TL_STAT + 3] = M.st.jacobians;
      T[TL_STAT + 4] = M.st.factorizations; T[TL_STAT + 5] = M.st.solves;
    }
    return;
  }
  mode_finish(P, C, md.y, n_int, status, tau_ini);
}

// Tail kernel: the radiation-streaming interval of every mode that perturb_kernel handed off.  Small
// shared-memory footprint and register budget (16+ warps per SM), compact hot loop (ndf15_rsa).
#define PT_TAIL_MAX_WPC 6
#ifndef PT_TAIL_MIN_BLOCKS
#define PT_TAIL_MIN_BLOCKS 12  // 168 registers: no spills in ndf15_rsa (16 -> 128 registers spills and is 25 % slower)
#endif
__global__ void __launch_bounds__(32 * PT_TAIL_MAX_WPC, PT_TAIL_MIN_BLOCKS / PT_TAIL_MAX_WPC) perturb_tail_kernel(const __grid_constant__ PtParams P) {
  if (PT_SLOT(P) >= P.n_modes) return;
  const double* T = P.tail + (size_t)PT_SLOT(P) * TL_STRIDE;
  if (T[TL_VALID] != 1.) return;
  Mode& M = MODE(P);
  const int lane = PT_LANE;
  const int2 md = P.modes[PT_SLOT(P)];
  const PtCosmo* C = P.cosmo + md.x;
  mode_init(P, C, md.y);
  const int iv = (int)T[TL_IV], n_int = (int)T[TL_NINT];
  Approx ap;
  ap.tca_off = (int)T[TL_FLAGS]; ap.rsa_on = (int)T[TL_FLAGS + 1]; ap.ufa_on = (int)T[TL_FLAGS + 2]; ap.ncdmfa_on = (int)T[TL_FLAGS + 3];
  if (lane == 0) {
    M.ap = ap;
    M.next = (int)T[TL_NEXT];
    M.st.steps = (int)T[TL_STAT]; M.st.failed = (int)T[TL_STAT + 1]; M.st.fevals = (int)T[TL_STAT + 2];
    M.st.jacobians = (int)T[TL_STAT + 3]; M.st.factorizations = (int)T[TL_STAT + 4]; M.st.solves = (int)T[TL_STAT + 5];
  }
  make_layout(P, ap, M.L);
  __syncwarp();
  const int n = M.L.neq;
  if (lane < n) s_vec(P, V_Y)[lane] = T[TL_Y + lane];
  if (lane == 0) { M.nh = n; M.nch = 0; }
  __syncwarp();
  const long long c0 = clock64();
  const int s0 = M.st.steps;
  bool ok;
  if (n <= 8) ok = ndf15_rsa<8>(P, T[TL_T0], T[TL_TF]);
  else ok = ndf15_rsa<16>(P, T[TL_T0], T[TL_TF]);
  __syncwarp();
  clpp_kstat* ks = C->kstat + md.y;
  if (lane == 0) {
    ks->iv_neq[iv] = n;
    ks->iv_steps[iv] = M.st.steps - s0;
    ks->iv_cycles[iv] = clock64() - c0;
#ifdef PT_PROF
    for (int q = 0; q < PF_COUNT; q++) ks->prof[iv * 12 + q] = M.prof[q];
#endif
  }
  __syncwarp();
  (void)ok;
  mode_finish(P, C, md.y, n_int, M.status, T[TL_TAU_INI]);
}

// =============================================================================================
// LANE KERNEL: one THREAD per mode (lane.cuh).  The default path for evolver = ndf15.
// =============================================================================================
#include "lane.cuh"

// W = doubles of per-thread state (compile-time size of the local-memory slab); W = 0: slab in global memory
// (fallback for state vectors larger than the largest instantiation; correct but slow)
#ifndef LN_MIN_CTAS
#define LN_MIN_CTAS 16  // register budget of the lane kernel (CTA = one warp): 16 CTAs per SM -> 128 registers per thread, so
                        // that 8 lane warps still fit beside a resident CTA of the warp-per-mode kernel (252 registers x 128 threads)
#endif
template <int W>
__global__ void __launch_bounds__(LN_CTA, LN_MIN_CTAS) perturb_lane_kernel(const __grid_constant__ PtParams P) {
  const int slot = blockIdx.x * LN_CTA + threadIdx.x;
  if (slot >= P.n_modes) return;
  const int2 md = P.modes[slot];
  if constexpr (W > 0) {
    double slab[W];
    ln_mode(P, slab, P.cosmo + md.x, md.y);
  } else {
    ln_mode(P, P.lane_scratch + (size_t)slot * P.ln_words, P.cosmo + md.x, md.y);
  }
}

// LANE TAIL: the radiation-streaming interval of the handed-off modes with k < P.tail_lane_kmax, one THREAD per mode
// (ln_ndf15_tail: everything in registers).  A tail step attempt keeps 7 of a warp's 32 lanes busy in perturb_tail_kernel and
// costs it 15-25 k cycles; a thread needs ~38 k cycles for the same attempt but a warp then advances 32 modes, so the tails
// of the bulk (70 % of all step attempts of a launch, half of them in modes below k = 8/Mpc) leave the warp kernels' SM
// slots to the generic phases.  Neighbouring slots hold neighbouring k of the sorted mode list (the same k of the
// cosmologies of a sweep): the lanes of a warp take nearly the same number of steps.  Only the few longest chains stay in
// perturb_tail_kernel: a lone thread walks them 2.6x slower than a lone warp, and they are the critical path of a launch.
// P here is the generic kernel's parameter block with the slab geometry of a <= 16-equation system (np = 16, lo_vec = 0).
#ifndef LN_TAIL_MIN_CTAS
#define LN_TAIL_MIN_CTAS 8  // CTA = one warp; 8 per SM -> up to 255 registers per thread (the state of 7 equations is ~100 doubles)
#endif
#define LN_TAIL_NP 16
__global__ void __launch_bounds__(LN_CTA, LN_TAIL_MIN_CTAS) perturb_tail_lane_kernel(const __grid_constant__ PtParams P) {
  const int slot = blockIdx.x * LN_CTA + threadIdx.x;
  if (slot >= P.n_modes) return;
  const double* T = P.tail + (size_t)slot * TL_STRIDE;
  if (T[TL_VALID] != 2.) return;
  const int2 md = P.modes[slot];
  const PtCosmo* C = P.cosmo + md.x;
  const int ik = md.y;
  double slab[5 * LN_TAIL_NP];  // LV_Y, and LV_TMP / LV_YPI (slots 4 and 3) of the source output
  double* mem = slab;
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
  M.bg_cur = -1000000; M.th_cur = -1000000;  // first lookup by bisection
  M.need_nw = 0; M.status = 0; M.tca_shear_last = 0.; M.fac_c = 0.;
  M.m.h_prime = M.m.eta_prime = M.m.alpha = M.m.alpha_prime = M.m.rsa_delta_g = M.m.rsa_theta_g = 0.;
  M.m.delta_m = M.m.delta_cb = M.m.tca_shear_g = 0.;
  M.ap.tca_off = (int)T[TL_FLAGS]; M.ap.rsa_on = (int)T[TL_FLAGS + 1]; M.ap.ufa_on = (int)T[TL_FLAGS + 2];
  M.ap.ncdmfa_on = (int)T[TL_FLAGS + 3];
  M.apprev = M.ap;
  ln_make_layout(P, M.ap, M.L);
  M.Lprev = M.L;
  M.next = (int)T[TL_NEXT];
  M.st.steps = (int)T[TL_STAT]; M.st.failed = (int)T[TL_STAT + 1]; M.st.fevals = (int)T[TL_STAT + 2];
  M.st.jacobians = (int)T[TL_STAT + 3]; M.st.factorizations = (int)T[TL_STAT + 4]; M.st.solves = (int)T[TL_STAT + 5];
  const int n = M.L.neq;
  for (int i = 0; i < n; i++) LVP(LV_Y)[i] = T[TL_Y + i];
  const int iv = (int)T[TL_IV], n_int = (int)T[TL_NINT];
  const long long c0 = clock64();
  const int s0 = M.st.steps;
  ln_run_tail(P, M, mem, T[TL_T0], T[TL_TF]);
  clpp_kstat* ks = C->kstat + ik;
  ks->iv_neq[iv] = n;
  ks->iv_steps[iv] = M.st.steps - s0;
  ks->iv_cycles[iv] = clock64() - c0;
  // zero-fill the samples that were not reached (failure only) and publish the counters (mode_finish)
  const int tau_size = C->tau_size;
  const size_t stride_tp = (size_t)C->k_size * tau_size;
  double* out = C->sources + (size_t)ik * tau_size;
  const int tps[7] = {P.tp_t0, P.tp_t1, P.tp_t2, P.tp_p, P.tp_delta_m, P.tp_delta_cb, P.tp_phi_plus_psi};
  for (int it = M.next; it < tau_size; it++)
    for (int j = 0; j < 7; j++)
      if (tps[j] >= 0) out[tps[j] * stride_tp + it] = 0.;
  ks->steps = M.st.steps; ks->failed = M.st.failed; ks->fevals = M.st.fevals; ks->jacobians = M.st.jacobians;
  ks->factorizations = M.st.factorizations; ks->solves = M.st.solves;
  ks->intervals = n_int; ks->status = M.status; ks->tau_ini = T[TL_TAU_INI];
}

// =============================================================================================
// host side
// =============================================================================================

static void set_geometry(PtParams& P, int neq_max, int nh_max, const clpp_perturb_desc& pd);

// settings shared by every cosmology of a batch (pointers and geometry are filled by the caller)
static int fill_common(const clpp_ctx* c, PtParams& P, char* err) {
  const clpp_perturb_desc& pd = c->pd;
  const clpp_background_desc& bg = c->bg;
  const clpp_thermo_desc& th = c->th;
  const clpp_perturb_info& I = c->pinfo;
  memset(&P, 0, sizeof(P));
  CLPP_CHECK(c->N_ncdm <= PT_MAX_NCDM, err, "at most %d ncdm species are supported on the device", PT_MAX_NCDM);
  CLPP_CHECK(bg.bg_size_normal <= 32 && th.th_size <= 32, err, "background/thermo vectors wider than a warp");
  P.bg_size = bg.bg_size; P.bg_size_normal = bg.bg_size_normal; P.th_size = th.th_size;
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
  // the reference leaves the indices of absent columns unset: normalise them so that equal settings compare equal
  if (!P.compute_damping_scale) P.ir_d = 0;
  if (!P.compute_cb2_derivatives) { P.idcb2 = 0; P.iddcb2 = 0; }
  if (!bg.has_ur) P.irho_ur = 0;
  if (!bg.has_ncdm) { P.irho_ncdm1 = 0; P.ip_ncdm1 = 0; P.ipseudo_p_ncdm1 = 0; }
  P.has_ur = bg.has_ur; P.has_ncdm = bg.has_ncdm; P.N_ncdm = bg.has_ncdm ? bg.N_ncdm : 0;
  int off = 0;
  for (int s = 0; s < P.N_ncdm; s++) {
    P.ncdm_q_size[s] = c->ncdm_q_size[s];
    P.ncdm_q_off[s] = off;
    off += c->ncdm_q_size[s];
  }
  P.nq_tot = off;
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
  P.evolver = pd.evolver; P.rk_stepsize = pd.perturb_integration_stepsize;
  P.curvature_ini = pd.curvature_ini; P.three_ceff2_ur = pd.three_ceff2_ur; P.three_cvis2_ur = pd.three_cvis2_ur;
  P.switch_sw = pd.switch_sw; P.switch_eisw = pd.switch_eisw; P.switch_lisw = pd.switch_lisw;
  P.switch_dop = pd.switch_dop; P.switch_pol = pd.switch_pol; P.eisw_lisw_split_z = pd.eisw_lisw_split_z;
  P.tp_t0 = I.index_tp_t0; P.tp_t1 = I.index_tp_t1; P.tp_t2 = I.index_tp_t2; P.tp_p = I.index_tp_p;
  P.tp_delta_m = I.index_tp_delta_m; P.tp_delta_cb = I.index_tp_delta_cb; P.tp_phi_plus_psi = I.index_tp_phi_plus_psi;

  // largest state vector over the approximation phases (full hierarchy, everything off) and its hub block
  int neq_max = 2 + 1 + (pd.l_max_g - 2) + (pd.l_max_pol_g + 1) + 3 + 1;
  int nh_max = 3 + 3 + 3 + 1;
  int n_chains = 2;
  if (bg.has_ur) { neq_max += 3 + (pd.l_max_ur - 2); nh_max += 3; n_chains++; }
  neq_max += P.nq_tot * (pd.l_max_ncdm + 1);
  nh_max += 3 * P.nq_tot;
  n_chains += P.nq_tot;
  CLPP_CHECK(n_chains <= PT_MAX_CHAINS, err,
             "%d multipole hierarchies (photons, ur, ncdm momentum bins) exceed the %d chains a warp handles: reduce "
             "the number of ncdm momentum bins", n_chains, PT_MAX_CHAINS);
  CLPP_CHECK(nh_max <= 128, err, "%d hub variables exceed the 128 a warp handles: reduce the number of ncdm momentum bins",
             nh_max);
  set_geometry(P, neq_max, nh_max, pd);
  return CLPP_SUCCESS;
}

// shared-memory layout of one CTA for state vectors of up to neq_max equations / nh_max hub variables
static void set_geometry(PtParams& P, int neq_max, int nh_max, const clpp_perturb_desc& pd) {
  P.neq_max = neq_max;
  P.np = (neq_max + 1) & ~1;
  P.nh_max = nh_max;
  P.ldh = nh_max | 1;
  P.o_mode = 64;
  P.o_hubtmp = P.o_mode + (int)((sizeof(Mode) + 7) / 8);
  P.o_nw = P.o_hubtmp + std::max(nh_max, 32);
  P.o_i2l1 = P.o_nw + 4 * ((P.nq_tot + 1) & ~1);
  P.n_i2l1 = (std::max(std::max(pd.l_max_g, pd.l_max_pol_g), std::max(pd.l_max_ur, pd.l_max_ncdm)) + 2 + 1) & ~1;
  P.ncol = (std::max(P.bg_size_normal, P.th_size) <= 16) ? 16 : 32;
  P.o_tabc = P.o_i2l1 + P.n_i2l1;
  P.o_vec = P.o_tabc + 8 * P.ncol;
  P.o_sinv = P.o_vec + V_COUNT * P.np;
  P.o_int = P.o_sinv + ((P.nh_max * P.ldh + 1) & ~1);
  P.wpc = 1;
  P.scr_stride = P.nh_max * P.nh_max + 4 * P.np;
  P.wstride = (int)((((size_t)P.o_int * sizeof(double) + (size_t)(2 * P.nh_max + 3 * PT_MAX_CHAINS) * sizeof(int)) + 15) / 16 * 2);
  // per-thread slab of the lane kernels (lane.cuh)
  P.lo_vec = 0;
  P.lo_nw = LV_COUNT * P.np;
  P.lo_jhh = P.lo_nw + 4 * P.nq_tot;
  P.lo_lu = P.lo_jhh + P.nh_max * P.nh_max;
  P.lo_piv = P.lo_lu + P.nh_max * P.nh_max;
  // hub region: dense (J, LU, pivots) or structured (ln_jacobian_s: 29 nh + 20 doubles), whichever is larger
  P.lo_ch = P.lo_jhh + std::max(2 * P.nh_max * P.nh_max + P.nh_max, 29 * P.nh_max + 20);
  P.ln_structured = getenv("CLPP_LANE_DENSE") ? 0 : 1;
  P.ln_words = P.lo_ch + 2 * PT_MAX_CHAINS;
}

// the common block of a context, for the developer harness tests/hostsim (CPU execution of the lane program)
extern "C" int clpp_pt_fill_common(const clpp_ctx* c, PtParams* P, char* err) { return fill_common(c, *P, err); }

// dynamic shared memory of one warp (one k mode); a CTA of P.wpc warps takes wpc times this
static size_t perturb_smem_bytes(const PtParams& P) { return (size_t)P.wstride * sizeof(double); }

// Integrates modes [k_begin[i], k_end[i]) of every context in ONE kernel launch on the stream of
// the first context (all contexts must live on the same device).
// `k_list` (optional): explicit mode indices of context 0 instead of the range (cost-balanced multi-GPU partitions).
int clpp_dev_perturb_solve_batch(clpp_ctx** cs, int n_ctx, const int* k_begin, const int* k_end, const int* k_list,
                                 int n_list, char* err) {
  CLPP_CHECK(n_ctx >= 1, err, "empty batch");
  CLPP_CHECK(k_list == nullptr || n_ctx == 1, err, "an explicit mode list is only supported for a single context");
  // the modes of context b that this call integrates
  auto n_sel = [&](int b) { return (k_list && b == 0) ? n_list : k_end[b] - k_begin[b]; };
  auto sel = [&](int b, int i) { return (k_list && b == 0) ? k_list[i] : k_begin[b] + i; };
  clpp_ctx* c0 = cs[0];
  clpp_ctx::Dev* d0 = c0->dev;
  cudaStream_t st = d0->stream;
  PtParams P;
  if (fill_common(c0, P, err)) return CLPP_FAILURE;
  for (int b = 1; b < n_ctx; b++) {
    CLPP_CHECK(cs[b]->dev && cs[b]->device == c0->device, err, "all contexts of a batch must live on the same CUDA device");
    PtParams Q;
    if (fill_common(cs[b], Q, err)) return CLPP_FAILURE;
    if (memcmp(&P, &Q, sizeof(P)) != 0) {
      size_t off = 0;
      while (off < sizeof(P) && ((const unsigned char*)&P)[off] == ((const unsigned char*)&Q)[off]) off++;
      return clpp_fail(err, "cosmology %d of the batch differs from cosmology 0 in the precision settings / species content / "
                       "requested sources (first difference at byte %zu of the common block): such cosmologies must be solved "
                       "in separate batches", b, off);
    }
  }

  std::vector<PtCosmo> cosmo(n_ctx);
  std::vector<int2> modes;
  std::vector<double> cost;
  for (int b = 0; b < n_ctx; b++) {
    clpp_ctx* c = cs[b];
    clpp_ctx::Dev* d = c->dev;
    const clpp_perturb_info& I = c->pinfo;
    const int nk = I.k_size, nt = I.tau_size, ntp = I.tp_size;
    const size_t nsrc = (size_t)ntp * nk * nt;
    cudaStream_t sb = d->stream;
    if (clpp_dev_reserve(d, &d->sources, nsrc, err)) return CLPP_FAILURE;
    if (d->sources_count != nsrc) CLPP_CUDA(cudaMemsetAsync(d->sources, 0, nsrc * sizeof(double), sb), err);
    d->sources_count = nsrc;
    if (clpp_dev_reserve(d, &d->k, nk, err) || clpp_dev_reserve(d, &d->tau, nt, err) ||
        clpp_dev_reserve(d, &d->kstat, nk, err))
      return CLPP_FAILURE;
    CLPP_CUDA(cudaMemcpyAsync(d->k, c->k.data(), nk * sizeof(double), cudaMemcpyHostToDevice, sb), err);
    CLPP_CUDA(cudaMemcpyAsync(d->tau, c->tau.data(), nt * sizeof(double), cudaMemcpyHostToDevice, sb), err);
    CLPP_CUDA(cudaMemsetAsync(d->kstat, 0, nk * sizeof(clpp_kstat), sb), err);
    const size_t tot = c->ncdm_q.size();
    if (c->N_ncdm > 0) {
      if (clpp_dev_reserve(d, &d->ncdm, 3 * tot, err)) return CLPP_FAILURE;
      CLPP_CUDA(cudaMemcpyAsync(d->ncdm, c->ncdm_q.data(), tot * sizeof(double), cudaMemcpyHostToDevice, sb), err);
      CLPP_CUDA(cudaMemcpyAsync(d->ncdm + tot, c->ncdm_w.data(), tot * sizeof(double), cudaMemcpyHostToDevice, sb), err);
      CLPP_CUDA(cudaMemcpyAsync(d->ncdm + 2 * tot, c->ncdm_dlnf0.data(), tot * sizeof(double), cudaMemcpyHostToDevice, sb), err);
    }
    PtCosmo& Q = cosmo[b];
    memset(&Q, 0, sizeof(Q));
    Q.bg_tau = d->bg_tau; Q.bg_y = d->bg_y; Q.bg_dd = d->bg_dd;
    Q.th_z = d->th_z; Q.th_y = d->th_y; Q.th_dd = d->th_dd;
    Q.ncdm_q = d->ncdm; Q.ncdm_w = d->ncdm ? d->ncdm + tot : nullptr; Q.ncdm_dlnf0 = d->ncdm ? d->ncdm + 2 * tot : nullptr;
    Q.k = d->k; Q.tau = d->tau; Q.sources = d->sources; Q.kstat = d->kstat;
    Q.bt_size = c->bg.bt_size; Q.tt_size = c->th.tt_size; Q.k_size = nk; Q.tau_size = nt;
    Q.th_linear_below_z = -1.;
    if (c->th.reio_parametrization == CLPP_REIO_HALF_TANH) Q.th_linear_below_z = 2 * c->th.z_reionization;
    if (c->th.reio_parametrization == CLPP_REIO_INTER) Q.th_linear_below_z = 50.;
    Q.n_e = c->th.n_e; Q.YHe = c->th.YHe; Q.T_cmb = c->bg.T_cmb; Q.tau_free_streaming = c->th.tau_free_streaming;
    Q.a_today = c->bg.a_today;
    for (int s = 0; s < P.N_ncdm; s++) { Q.ncdm_M[s] = c->ncdm_M[s]; Q.ncdm_factor[s] = c->ncdm_factor[s]; }
    for (int i = 0; i < n_sel(b); i++) {
      const int ik = sel(b, i);
      CLPP_CHECK(ik >= 0 && ik < nk, err, "mode index %d outside [0,%d)", ik, nk);
      modes.push_back(make_int2(b, ik));
      cost.push_back(c->k[ik]);  // the number of steps of a mode grows with k tau_0
    }
    if (b > 0) CLPP_CUDA(cudaStreamSynchronize(sb), err);  // uploads of the other contexts' streams
  }
  // issue order: decreasing expected cost (longest chains first), across the whole batch
  const int n_modes = (int)modes.size();
  std::vector<int> perm(n_modes);
  for (int i = 0; i < n_modes; i++) perm[i] = i;
  // (developer knob CLPP_SORT_BLOCK = n: cohorts of n ADJACENT k of the same cosmology instead of the same k of n neighbouring
  //  cosmologies -- the warps of a cohort then walk through the same background / thermodynamics table rows)
  const int sort_block = getenv("CLPP_SORT_BLOCK") ? std::max(1, atoi(getenv("CLPP_SORT_BLOCK"))) : 1;
  if (sort_block > 1 && k_list == nullptr) {
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) {
      const int ka = modes[a].y / sort_block, kb = modes[b].y / sort_block;
      if (ka != kb) return ka > kb;
      if (modes[a].x != modes[b].x) return modes[a].x < modes[b].x;
      return modes[a].y > modes[b].y;
    });
  } else {
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  }
  // ---- two families of kernels.
  // WARPS (one warp per mode, below; the default): 2.2 s for one Planck-18 cosmology, 66 ms per cosmology in a batch of 128
  // different ones.
  // LANES (lane.cuh, one THREAD per mode; opt-in: context option "lane_path" or CLPP_LANE = 1): every instruction serves 32
  // modes and tens of thousands of modes run at once, but a GPU thread walks a serial chain ~10x slower than a warp does
  // (a step attempt of the 7-equation tail: 31 k cycles in a lane, 15 k in a warp; of the 136-equation system: 3-5 M cycles,
  // its 35 KB of state stream from L2), so a launch lasts >= 14 s (identical cosmologies) .. 43 s (different ones: some lane
  // of a warp refactorises / outputs at almost every attempt) whatever the batch; beyond ~500 cosmologies per launch the
  // cost per cosmology (9 ms identical) undercuts the warp kernels.
  // HYBRID (lane_path / CLPP_LANE = 2): modes with k >= kcut in the warp kernels, the bulk in the lane kernel on its own
  // stream.  Measured on 128 different cosmologies (profiles/r02_lane_vs_warp.txt): the halves take 4.4 s + 5.6 s alone and
  // 9.7 s together -- both fill the register file, so they run one after the other -- against 8.5 s for the warp kernels alone.
  int lane_mode = 0;  // 0: warps only, 1: lanes only, 2: hybrid
  if (c0->pd.evolver == 1 && !getenv("CLPP_WARP_PATH") && !getenv("CLPP_GENERIC_ONLY") && !getenv("CLPP_NO_TAIL") &&
      !getenv("CLPP_COHORT")) {
    const char* e = getenv("CLPP_LANE");
    const int forced = c0->lane_path >= 0 ? c0->lane_path : (e ? atoi(e) : -1);
    lane_mode = forced >= 0 ? forced : 0;
  }
  const double lane_kcut = getenv("CLPP_LANE_KCUT") ? atof(getenv("CLPP_LANE_KCUT")) : 0.6;  // 1/Mpc
  int n_lane = 0;  // the last n_lane modes of the cost-sorted order run in the lane kernel
  if (lane_mode == 1) n_lane = n_modes;
  else if (lane_mode == 2)
    while (n_lane < n_modes && cost[perm[n_modes - 1 - n_lane]] < lane_kcut) n_lane++;
  int n_warp_modes = n_modes - n_lane;
  // developer knob (timing the two halves of a hybrid launch separately; the skipped modes are simply not integrated)
  if (getenv("CLPP_HYBRID_SKIP")) {
    if (!strcmp(getenv("CLPP_HYBRID_SKIP"), "warp")) n_warp_modes = 0;
    else n_lane = 0;
  }
  if (clpp_dev_reserve(d0, &d0->pt_cosmo, n_ctx * sizeof(PtCosmo), err)) return CLPP_FAILURE;
  CLPP_CUDA(cudaMemcpyAsync(d0->pt_cosmo, cosmo.data(), n_ctx * sizeof(PtCosmo), cudaMemcpyHostToDevice, st), err);
  P.cosmo = (const PtCosmo*)d0->pt_cosmo;
  cudaEventRecord(d0->ev[0], st);
  if (n_lane > 0) {
    std::vector<int2> lsorted(n_lane);
    for (int i = 0; i < n_lane; i++) lsorted[i] = modes[perm[n_modes - n_lane + i]];
    if (clpp_dev_reserve(d0, &d0->ln_modes, (size_t)n_lane * sizeof(int2), err)) return CLPP_FAILURE;
    CLPP_CUDA(cudaMemcpyAsync(d0->ln_modes, lsorted.data(), n_lane * sizeof(int2), cudaMemcpyHostToDevice, st), err);
    std::vector<double> i2l1(P.n_i2l1);
    for (int l = 0; l < P.n_i2l1; l++) i2l1[l] = 1.0 / (2.0 * l + 1.0);
    if (clpp_dev_reserve(d0, &d0->i2l1, (size_t)P.n_i2l1, err)) return CLPP_FAILURE;
    CLPP_CUDA(cudaMemcpyAsync(d0->i2l1, i2l1.data(), P.n_i2l1 * sizeof(double), cudaMemcpyHostToDevice, st), err);
    const int n_cta = (n_lane + LN_CTA - 1) / LN_CTA;
    static const int slab_sizes[] = {1536, 4608, 14336};
    int slab = 0;
    for (int w : slab_sizes)
      if (slab == 0 && P.ln_words <= w) slab = w;
    if (slab == 0 && clpp_dev_reserve(d0, &d0->lane_scratch, (size_t)n_lane * P.ln_words, err)) return CLPP_FAILURE;
    PtParams PL = P;
    PL.modes = (const int2*)d0->ln_modes;
    PL.n_modes = n_lane;
    PL.lane_scratch = d0->lane_scratch;
    PL.i2l1 = d0->i2l1;
    if (getenv("CLPP_VERBOSE"))
      fprintf(stderr, "[clpp] perturb (lane kernel): %d of %d modes (k < %g/Mpc), %d CTAs of %d threads, %d doubles of state per mode "
              "(slab %d), neq_max %d, hub %d\n", n_lane, n_modes, lane_mode == 2 ? lane_kcut : 1e30, n_cta, LN_CTA, P.ln_words, slab,
              P.neq_max, P.nh_max);
    if (!d0->lane_stream) {
      int prio_lo = 0, prio_hi = 0;
      cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
      CLPP_CUDA(cudaStreamCreateWithPriority(&d0->lane_stream, cudaStreamNonBlocking, prio_lo), err);
      CLPP_CUDA(cudaEventCreateWithFlags(&d0->lane_done, cudaEventDisableTiming), err);
      CLPP_CUDA(cudaEventCreateWithFlags(&d0->lane_go, cudaEventDisableTiming), err);
    }
    cudaStream_t sl = n_warp_modes > 0 ? d0->lane_stream : st;  // lanes only: the context stream itself
    if (sl != st) {
      cudaEventRecord(d0->lane_go, st);  // uploads on st are complete
      cudaStreamWaitEvent(sl, d0->lane_go, 0);
    }
    // Hybrid: the lane warps must leave room for the warp-per-mode kernels launched next (one 128-thread CTA of 252
    // registers and 115 KB of shared memory per SM): an (unused) dynamic shared-memory request of 14 KB caps the lane kernel
    // at 8 of its 16 possible warps per SM, i.e. half of the register file and 112 KB of shared memory.
    size_t lane_smem = (n_warp_modes > 0) ? 14 * 1024 : 0;
    if (getenv("CLPP_LANE_SMEM_KB")) lane_smem = (size_t)atoi(getenv("CLPP_LANE_SMEM_KB")) * 1024;  // developer knob
    switch (slab) {
      case 1536: perturb_lane_kernel<1536><<<n_cta, LN_CTA, lane_smem, sl>>>(PL); break;
      case 4608: perturb_lane_kernel<4608><<<n_cta, LN_CTA, lane_smem, sl>>>(PL); break;
      case 14336: perturb_lane_kernel<14336><<<n_cta, LN_CTA, lane_smem, sl>>>(PL); break;
      default: perturb_lane_kernel<0><<<n_cta, LN_CTA, lane_smem, sl>>>(PL); break;
    }
    c0->launches++;
    if (sl != st) cudaEventRecord(d0->lane_done, sl);
  }
  if (n_warp_modes > 0) {
    const int n_modes = n_warp_modes;  // (shadows the total: this block launches the first n_warp_modes of the sorted order)
    // Launch groups.  Group L: the modes with long radiation-streaming tails (k >= 3 % of k_max: measured optimum), high-priority
    // stream, issued first: they finish their early phases quickly and run their tails -- the serial critical
    // path -- while the bulk is still in the generic kernel.  The bulk is dealt round-robin into chunks, one
    // low-priority stream each (generic kernel -> tail kernel), so that the tails of a chunk overlap the
    // generic phases of the next ones instead of all waiting for the last generic CTA.
    const bool force_generic = getenv("CLPP_GENERIC_ONLY") != nullptr;
    const bool use_tail = getenv("CLPP_NO_TAIL") == nullptr && !force_generic && c0->pd.evolver == 1;
    int n_long = 0;
    if (use_tail && n_modes > 0) {
      const double kcut_frac = getenv("CLPP_KCUT") ? atof(getenv("CLPP_KCUT")) : 0.03;  // developer knob
      const double kcut = kcut_frac * cost[perm[0]];
      while (n_long < n_modes && cost[perm[n_long]] >= kcut) n_long++;
      if (n_long == n_modes) n_long = 0;  // nothing to overlap with
    }
    const int n_bulk = n_modes - n_long;
    // cohort width: modes per CTA. Batches of several cosmologies put the same k of neighbouring cosmologies side by side in
    // the sorted order (near-identical step sequences); a single cosmology is a latency problem and keeps one mode per CTA.
    const size_t smem1 = perturb_smem_bytes(P);
    // (measured, scripts/time_varied.py on 128 different cosmologies: cohorts of 8 = one 256-thread CTA per SM 7.8 s, of 4 = two
    //  CTAs per SM 8.7 s, of 6 8.7 s: gpurun_out/r2_time_varied_v5.log in profiles/r02_lane_vs_warp.txt)
    int wpc = getenv("CLPP_COHORT") ? atoi(getenv("CLPP_COHORT")) : (n_ctx >= 64 ? 8 : n_ctx >= 4 ? 4 : 1);
    wpc = std::max(1, std::min(wpc, PT_MAX_WPC));
    while (wpc > 1 && smem1 * wpc > 227 * 1024) wpc--;
    int wpc_tail = getenv("CLPP_COHORT_TAIL") ? atoi(getenv("CLPP_COHORT_TAIL")) : std::min(wpc, PT_TAIL_MAX_WPC);
    wpc_tail = std::max(1, std::min(wpc_tail, PT_TAIL_MAX_WPC));
    // the long-tail group is the latency-critical path and keeps one mode per CTA: lockstep with neighbours only delays it
    // (measured, scripts/sweep_varied.py: 32 different cosmologies lose 15 % with cohorts there, 96 gain 4 % with cohorts of 4;
    //  scripts/time_varied.py with bulk cohorts of 8 on 128 different cosmologies: long group in cohorts of 8 7.84 s, of 4
    //  7.86 s, of 1 7.65 s -- profiles/r02_time_varied_v7.log)
    int wpc_long = getenv("CLPP_COHORT_LONG") ? atoi(getenv("CLPP_COHORT_LONG")) : 1;
    wpc_long = std::max(1, std::min(wpc_long, wpc));
    const int chunk_modes = getenv("CLPP_CHUNK_MODES") ? atoi(getenv("CLPP_CHUNK_MODES")) : 4000;  // developer knob
    const int n_chunks = use_tail ? std::max(1, std::min(PT_MAX_CHUNKS, n_bulk / std::max(chunk_modes, 1))) : 1;
    std::vector<int2> sorted(n_modes);
    std::vector<int> chunk_first(n_chunks + 1, n_long);
    for (int i = 0; i < n_long; i++) sorted[i] = modes[perm[i]];
    {
      int pos = n_long;
      for (int c = 0; c < n_chunks; c++) {
        chunk_first[c] = pos;
        for (int i = n_long; i < n_modes; i++)  // cohorts (wpc consecutive modes of the sorted order) stay together
          if (((i - n_long) / wpc) % n_chunks == c) sorted[pos++] = modes[perm[i]];
      }
      chunk_first[n_chunks] = pos;
    }
  
    if (clpp_dev_reserve(d0, &d0->pt_modes, (size_t)std::max(n_modes, 1) * sizeof(int2), err)) return CLPP_FAILURE;
    if (clpp_dev_reserve(d0, &d0->jac_scratch, (size_t)std::max(n_modes, 1) * P.scr_stride, err))
      return CLPP_FAILURE;
    CLPP_CUDA(cudaMemcpyAsync(d0->pt_modes, sorted.data(), n_modes * sizeof(int2), cudaMemcpyHostToDevice, st), err);
    P.modes = (const int2*)d0->pt_modes;
    P.n_modes = n_modes;
    P.hub_jac = d0->jac_scratch;
    // hand-off records of the tail (radiation-streaming) kernel
    P.force_generic = force_generic;
    if (use_tail) {
      if (clpp_dev_reserve(d0, &d0->pt_tail, (size_t)std::max(n_modes, 1) * TL_STRIDE, err)) return CLPP_FAILURE;
      CLPP_CUDA(cudaMemsetAsync(d0->pt_tail, 0, (size_t)std::max(n_modes, 1) * TL_STRIDE * sizeof(double), st), err);
      P.tail = d0->pt_tail;
    }
  
    const size_t smem = perturb_smem_bytes(P);
    CLPP_CHECK(smem <= 227 * 1024, err,
               "state vector of %d equations needs %zu bytes of shared memory per k-mode (> 227 KB): reduce l_max_ncdm / "
               "the number of ncdm momentum bins", P.neq_max, smem);
    // developer knob: extra dynamic shared memory per CTA of the generic kernel (limits the CTAs per SM)
    // Two unsynchronised cohorts on one SM share the instruction cache again: measured (bench.py --batch 64) the large
    // Planck-18 system (136 equations) runs 4 % faster with ONE 4-mode cohort per SM, the small LCDM one (46 equations,
    // smaller hot code) 1.7x faster with two. Large systems therefore claim more than half of the SM's shared memory.
    // (measured, scripts/time_varied.py: for 128 DIFFERENT cosmologies two cohorts per SM are 13 % faster again -- their
    // warps are out of step anyway -- so the padding only applies to small batches, where identical settings are likely)
    size_t smem_pad = 0;
    if (wpc > 1 && n_ctx < 8 && P.neq_max >= 100 && smem * wpc <= 114 * 1024) smem_pad = 115 * 1024 - smem * wpc;
    if (getenv("CLPP_SMEM_PAD_KB")) smem_pad = (size_t)atoi(getenv("CLPP_SMEM_PAD_KB")) * 1024;  // developer knob
    { static const cudaError_t once = clpp_allow_max_dynamic_smem(perturb_kernel); CLPP_CUDA(once, err); }
    P.wpc = wpc;
    P.sync_every = std::max(1, getenv("CLPP_SYNC_EVERY") ? atoi(getenv("CLPP_SYNC_EVERY")) : 1);  // developer knob
    PtParams Pt = P;  // geometry of the tail kernel: at most 16 equations, all hub
    set_geometry(Pt, 16, 16, c0->pd);
    const size_t smem_tail = perturb_smem_bytes(Pt);
    { static const cudaError_t once = clpp_allow_max_dynamic_smem(perturb_tail_kernel); CLPP_CUDA(once, err); }
    Pt.wpc = wpc_tail;
    if (getenv("CLPP_VERBOSE"))
      fprintf(stderr, "[clpp] perturb: %d modes, shared memory per mode %zu B (tail %zu B), sizeof(Mode) %zu, neq_max %d, hub %d, "
              "modes per CTA %d (tail %d)\n", n_modes, smem, smem_tail, sizeof(Mode), P.neq_max, P.nh_max, wpc, wpc_tail);
    for (int i = 0; i < 6; i++)
      if (!d0->ev2[i]) cudaEventCreate(&d0->ev2[i]);
    if (!d0->stream2) {
      int prio_lo = 0, prio_hi = 0;
      cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // numerically lower = higher priority
      CLPP_CUDA(cudaStreamCreateWithPriority(&d0->stream2, cudaStreamNonBlocking, prio_lo), err);
      CLPP_CUDA(cudaStreamCreateWithPriority(&d0->stream_hi, cudaStreamNonBlocking, prio_hi), err);
      for (int c = 0; c < PT_MAX_CHUNKS; c++) {
        CLPP_CUDA(cudaStreamCreateWithPriority(&d0->chunk_stream[c], cudaStreamNonBlocking, prio_lo), err);
        CLPP_CUDA(cudaEventCreateWithFlags(&d0->chunk_done[c], cudaEventDisableTiming), err);
      }
    }
    cudaStream_t sth = d0->stream_hi;
    // Tails.  Handed-off modes with k >= tail_lane_kmax run in perturb_tail_kernel (a warp per mode), the others in
    // perturb_tail_lane_kernel (a thread per mode).  OPT-IN (developer knob CLPP_TAIL_LANE_KMAX, 1/Mpc; default 0 = warps only):
    // measured on 128 DIFFERENT Planck-18 cosmologies (profiles/r02_lane_tail.txt) the launch takes 8.5 s with warp tails,
    // 9.6 s with lane tails below k = 3/Mpc, 10.7-16.5 s below 8/Mpc, 42-45 s with every tail in lanes: the 32 lanes of a warp
    // belong to 32 different cosmologies, some lane refactorises or writes a source sample at almost every attempt, and the
    // warp pays for the union of the branches (~340 k cycles per attempt against 38 k for identical cosmologies).
    double tail_lane_kmax = 0.;
    if (getenv("CLPP_TAIL_LANE_KMAX") && use_tail && n_ctx >= 8) tail_lane_kmax = atof(getenv("CLPP_TAIL_LANE_KMAX"));
    if (4 + 3 * P.N_ncdm > LN_TAIL_NP) tail_lane_kmax = 0.;
    P.tail_lane_kmax = tail_lane_kmax;
    PtParams Pl = P;  // slab geometry of the lane tail: vectors of LN_TAIL_NP doubles
    Pl.np = LN_TAIL_NP; Pl.lo_vec = 0; Pl.lo_nw = 5 * LN_TAIL_NP;
    if (tail_lane_kmax > 0. && !d0->tlane_stream) {
      int prio_lo = 0, prio_hi = 0;
      cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
      CLPP_CUDA(cudaStreamCreateWithPriority(&d0->tlane_stream, cudaStreamNonBlocking, prio_hi), err);
      CLPP_CUDA(cudaEventCreateWithFlags(&d0->tlane_go, cudaEventDisableTiming), err);
      CLPP_CUDA(cudaEventCreateWithFlags(&d0->tlane_done, cudaEventDisableTiming), err);
    }
    bool tlane_used = false;
    auto launch_group = [&](cudaStream_t s, int first, int count, int wpc, int wpc_tail) {
      if (count <= 0) return;
      PtParams G = P, Gt = Pt;
      G.wpc = wpc; Gt.wpc = wpc_tail;
      G.modes = P.modes + first; G.n_modes = count;
      G.hub_jac = P.hub_jac + (size_t)first * P.scr_stride;
      G.tail = P.tail ? P.tail + (size_t)first * TL_STRIDE : nullptr;
      perturb_kernel<<<(count + wpc - 1) / wpc, 32 * wpc, smem * wpc + smem_pad, s>>>(G);
      c0->launches++;
      if (use_tail) {
        bool any_warp = false, any_lane = false;
        for (int i = first; i < first + count; i++) {
          if (cs[sorted[i].x]->k[sorted[i].y] < tail_lane_kmax) any_lane = true; else any_warp = true;
        }
        const bool side = any_warp && any_lane && !tlane_used;  // lane tails beside the warp tails, on their own stream
        if (side) cudaEventRecord(d0->tlane_go, s);
        if (any_warp) {
          Gt.modes = G.modes; Gt.n_modes = count; Gt.tail = G.tail; Gt.hub_jac = G.hub_jac;
          perturb_tail_kernel<<<(count + wpc_tail - 1) / wpc_tail, 32 * wpc_tail, smem_tail * wpc_tail, s>>>(Gt);
          c0->launches++;
        }
        if (any_lane) {
          PtParams Gl = Pl;
          Gl.modes = G.modes; Gl.n_modes = count; Gl.tail = G.tail;
          cudaStream_t sl = side ? d0->tlane_stream : s;
          if (side) cudaStreamWaitEvent(sl, d0->tlane_go, 0);
          perturb_tail_lane_kernel<<<(count + LN_CTA - 1) / LN_CTA, LN_CTA, 0, sl>>>(Gl);
          c0->launches++;
          if (side) {
            cudaEventRecord(d0->tlane_done, sl);
            cudaStreamWaitEvent(s, d0->tlane_done, 0);
            tlane_used = true;
          }
        }
      }
    };
    if (n_modes > 0) {
      cudaEventRecord(d0->ev2[4], st);        // uploads on st are complete before the other streams start
      if (n_long > 0) {
        cudaStreamWaitEvent(sth, d0->ev2[4], 0);
        launch_group(sth, 0, n_long, wpc_long, std::min(wpc_long, wpc_tail));  // high priority: its tail CTAs take the slots as they free up
        cudaEventRecord(d0->ev2[5], sth);
        cudaStreamWaitEvent(st, d0->ev2[5], 0);
      }
      for (int c = 0; c < n_chunks; c++) {
        cudaStream_t sc = d0->chunk_stream[c];
        cudaStreamWaitEvent(sc, d0->ev2[4], 0);
        launch_group(sc, chunk_first[c], chunk_first[c + 1] - chunk_first[c], wpc, wpc_tail);
        cudaEventRecord(d0->chunk_done[c], sc);
        cudaStreamWaitEvent(st, d0->chunk_done[c], 0);
      }
    }
}
  if (n_lane > 0 && n_warp_modes > 0) cudaStreamWaitEvent(st, d0->lane_done, 0);
  cudaEventRecord(d0->ev[1], st);
  CLPP_CUDA(cudaGetLastError(), err);
  for (int b = 0; b < n_ctx; b++) {
    clpp_ctx* c = cs[b];
    const int nk = c->pinfo.k_size;
    c->kstat.assign(nk, clpp_kstat{});
    CLPP_CUDA(cudaMemcpyAsync(c->kstat.data(), c->dev->kstat, nk * sizeof(clpp_kstat), cudaMemcpyDeviceToHost, st), err);
  }
  CLPP_CUDA(cudaStreamSynchronize(st), err);
  {
    float ms = 0;
    cudaEventElapsedTime(&ms, d0->ev[0], d0->ev[1]);
    for (int b = 0; b < n_ctx; b++) { cs[b]->dev->t_perturb_ms = 0.; cs[b]->dev->t_perturb_tail_ms = 0.; }
    d0->t_perturb_ms = ms;
  }
  // per-cosmology outcome: a failing mode only invalidates ITS cosmology (the reference fails only the offending
  // Cosmology object); the call reports the first failure and how many cosmologies of the batch failed
  int n_failed = 0;
  char first[CLPP_ERRLEN];
  first[0] = 0;
  for (int b = 0; b < n_ctx; b++) {
    clpp_ctx* c = cs[b];
    bool ok = true;
    for (int i = 0; i < n_sel(b) && ok; i++) {
      const int ik = sel(b, i);
      const int s = c->kstat[ik].status;
      if (s != 0) {
        const char* what = s == 2 ? "Step size too small in the NDF15 evolver"
                         : s == 3 ? "your choice of initial time for integrating wavenumbers is inappropriate: it corresponds to a time before that at which the background has been integrated. You should increase 'start_small_k_at_tau_c_over_tau_h'"
                         : s == 4 ? "your choice of initial time for integrating wavenumbers is inappropriate: it corresponds to a time before that at which the background has been integrated. You should increase 'start_large_k_at_tau_h_over_tau_k'"
                         : s == 5 ? "your choice of initial time for integrating wavenumbers is inappropriate: ncdm species not ultra-relativistic"
                         : s == 6 ? "an approximation flag goes backward in time, this cannot be handled"
                         : s == 7 ? "you switch several approximations at the same time, this cannot be handled"
                         : s == 9 ? "Too many integration steps needed within one interval (rk evolver), the system of equations is probably buggy or featuring a discontinuity"
                         : s == 8 ? "scalar initial conditions assume tight coupling on and all other approximations off"
                                  : "unknown device error";
        if (n_failed == 0)
          snprintf(first, sizeof(first), "perturb_solve failed for k=%e (index %d, cosmology %d of the batch): %s", c->k[ik], ik, b, what);
        ok = false;
      }
    }
    c->has_sources = ok;
    c->nl_dev_valid = false;
    if (!ok) n_failed++;
  }
  if (n_failed > 0)
    return clpp_fail(err, "%s%s", first, n_failed > 1 ? " (and further cosmologies of the batch failed: check clpp_perturb_get_kstat)" : "");
  return CLPP_SUCCESS;
}

int clpp_dev_perturb_solve(clpp_ctx* c, int k_begin, int k_end, char* err) {
  return clpp_dev_perturb_solve_batch(&c, 1, &k_begin, &k_end, nullptr, 0, err);
}

int clpp_dev_perturb_solve_list(clpp_ctx* c, const int* k_list, int n, char* err) {
  const int zero = 0;
  return clpp_dev_perturb_solve_batch(&c, 1, &zero, &zero, k_list, n, err);
}
