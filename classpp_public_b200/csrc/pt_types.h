// Types shared by the stage-1 kernels (perturb.cu: one warp per mode; lane.cuh: one thread per mode).
#ifndef CLPP_PT_TYPES_H
#define CLPP_PT_TYPES_H

#include "clpp.h"
#ifndef __CUDACC__
struct int2 { int x, y; };  // host-only builds (tests/hostsim): the CUDA vector type
#endif

#define PT_MAX_NCDM 3
#define PT_MAX_INTERVALS 6
#define PT_MAX_CHAINS 32
#define PT_MAX_CHUNKS CLPP_PT_MAX_CHUNKS
#define PT_FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------
// per-cosmology inputs (global memory, one entry per context of the batch)
struct PtCosmo {
  const double *bg_tau, *bg_y, *bg_dd;
  const double *th_z, *th_y, *th_dd;
  const double *ncdm_q, *ncdm_w, *ncdm_dlnf0;
  const double *k, *tau;
  double* sources;  // [tp][k][tau]
  clpp_kstat* kstat;
  int bt_size, tt_size, k_size, tau_size;
  double th_linear_below_z;  // < 0: never use linear interpolation
  double n_e, YHe, T_cmb, tau_free_streaming, a_today;
  double ncdm_M[PT_MAX_NCDM], ncdm_factor[PT_MAX_NCDM];
};

// settings common to every cosmology of a batch (kernel parameter -> constant bank)
struct PtParams {
  const PtCosmo* cosmo;
  const int2* modes;  // (cosmology index, k index), decreasing expected cost
  int n_modes;
  double* hub_jac;  // per-CTA global scratch [nh_max*nh_max]: hub block of the Jacobian
  double* tail;     // per-mode hand-off records for perturb_tail_kernel (nullptr: no tail kernel)
  double tail_lane_kmax;  // handed-off modes with k below this run their tail in perturb_tail_lane_kernel (one thread per mode)
  int bg_size, bg_size_normal, th_size;
  // background column indices
  int ia, iH, iHp, irho_g, irho_b, irho_cdm, irho_ur, irho_ncdm1, ip_ncdm1, ipseudo_p_ncdm1;
  // thermo column indices
  int ixe, idkappa, iddkappa, idddkappa, iexp_m_kappa, ig, idg, iddg, icb2, iwb, iTb, itau_d, irate, ir_d, idcb2, iddcb2;
  int compute_cb2_derivatives, compute_damping_scale;
  int has_ur, has_ncdm, N_ncdm;
  int ncdm_q_size[PT_MAX_NCDM], ncdm_q_off[PT_MAX_NCDM], nq_tot;
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
  int tp_t0, tp_t1, tp_t2, tp_p, tp_delta_m, tp_delta_cb, tp_phi_plus_psi;
  // shared-memory geometry (offsets in doubles into the CTA's dynamic shared memory)
  int evolver;        // 0 = rk (Cash-Karp), 1 = ndf15
  double rk_stepsize; // perturb_integration_stepsize (rk only)
  int force_generic;  // developer/test switch: integrate every interval with the generic shared-memory NDF
  int wpc, wstride;   // warps (k modes) per CTA; doubles of shared memory per warp
  int sync_every;     // the cohort barrier is taken every sync_every-th step attempt of a warp (1: every attempt)
  int scr_stride;     // doubles of global scratch per mode (hub Jacobian + 4 vectors)
  int neq_max, np, nh_max, ldh;
  int o_mode, o_hubtmp, o_nw, o_i2l1, n_i2l1, o_tabc, ncol, o_vec, o_sinv, o_int;
  // ---- lane kernels (lane.cuh: one THREAD per mode): per-thread scratch in global memory, element e of thread t of CTA b
  // at lane_scratch[(b * ln_words + e) * LN_CTA + t] (interleaved: a warp touches 256 contiguous bytes per element)
  double* lane_scratch;
  const double* i2l1;  // 1/(2l+1), l = 0..n_i2l1-1
  int ln_words;        // doubles of scratch per thread
  int lo_vec, lo_nw, lo_jhh, lo_lu, lo_piv, lo_ch;  // offsets (doubles) of the vector slots, ncdm weights, hub J, hub LU, pivots, per-chain scalars
  int ln_structured;   // 1: hub solve by block elimination + low-rank metric coupling; 0: dense LU of the hub block
};

struct Approx {
  int tca_off, rsa_on, ufa_on, ncdmfa_on;  // monotone flags (0 -> 1 in time)
};

#endif
