/* clpp.h -- C ABI of the B200-native CLASS++ hot path (libclpp.so)
 *
 * Drop-in boundary for the one data-parallel path of CLASS++:
 *
 *   PerturbationsModule (per-k Einstein-Boltzmann integration, NDF15)  -> S(k,tau)
 *   TransferModule      (line-of-sight integrals over Bessel functions) -> Delta_l(q)
 *   SpectraModule       (C_l k-quadrature)                              -> C_l
 *
 * The reference has no FFI table: its boundary is three C++ constructors that do all the
 * work and expose public data members (reference: source/perturbations_module.h:7-178,
 * source/transfer_module.h:7-54, source/spectra_module.h:11-81).  This header is what a
 * thin replacement of those three constructors binds: plain pointers
 * and sizes, no C++/torch types, status return 0 = _SUCCESS_ / 1 = _FAILURE_ with the
 * message in a caller-supplied 2048-char buffer (the reference's ErrorMsg convention,
 * include/common.h:140-300).  Nothing throws across this ABI; all state lives in clpp_ctx.
 *
 * All arrays are row-major FP64 host pointers unless stated otherwise.
 * There is NO CPU fallback: every compute entry point fails if no CUDA device is usable.
 */
#ifndef CLPP_H
#define CLPP_H

#ifdef __cplusplus
extern "C" {
#endif

#define CLPP_SUCCESS 0
#define CLPP_FAILURE 1
#define CLPP_ERRLEN 2048
#define CLPP_MAX_NCDM 8

typedef struct clpp_ctx clpp_ctx;

/* ---- context -------------------------------------------------------------------------- */
/* replaces: nothing (the reference is a single address space); one ctx = one cosmology
 * being pushed through the three stages on one CUDA device. */
int clpp_ctx_create(int device, clpp_ctx** ctx, char* err);
void clpp_ctx_destroy(clpp_ctx* ctx);
/* number of kernels launched by this ctx so far (bench.py's gpu_launches) */
long clpp_ctx_launch_count(const clpp_ctx* ctx);
const char* clpp_version(void);
/* the CUDA stream (cudaStream_t) every kernel of this ctx is launched on, for event timing */
int clpp_ctx_get_stream(clpp_ctx* ctx, void** stream);
/* device time (ms, CUDA events on the ctx stream) of the kernels of the last stage calls:
 * out[0] perturb_kernel + perturb_tail_kernel, [1] k_spline_kernel, [2] bessel_table_kernel, [3] los_kernel,
 * [4] spectra kernels, [5] perturb_tail_kernel alone (included in [0]; 0 when the groups overlap), [6] halofit_kernel,
 * [7] lensing kernels */
int clpp_ctx_get_kernel_ms(const clpp_ctx* ctx, double out[8]);
/* options of a context. "lean_scratch" = 1: the transfer stage takes its work buffers (copy of the sources, their k-spline,
 * the Bessel table: ~90 MB) from the stream-ordered memory pool of the device and returns them when the stage has run,
 * so that a sweep can keep thousands of contexts resident (45 MB each: sources + transfer functions + tables);
 * the Bessel accessor clpp_transfer_get_bessel is then unavailable.  "lane_path" = 0 / 1: force the warp-per-mode /
 * thread-per-mode perturbation kernels (default -1: by batch size). */
int clpp_ctx_set_option(clpp_ctx* ctx, const char* name, double value, char* err);
/* FP64 vector-pipe peak of this device measured with a dependent-free DFMA loop (TFLOP/s);
 * the denominator of the stage-1/2 roofline (MEASURED_PEAKS.json has no FP64 entry) */
int clpp_measure_fp64_peak(clpp_ctx* ctx, double* tflops, char* err);

/* ---- upstream tables (inputs of the path) ------------------------------------------------ */
/* replaces the reads of BackgroundModule public members (source/background_module.h:23-85):
 * tau_table_[bt_size], background_table_[bt_size*bg_size], index_bg_*_, bg_size_{short,normal}_,
 * conformal_age_, and struct background fields (source/background.h).  The second-derivative
 * table (private in the reference) is rebuilt with the same natural-order spline recurrences
 * as array_spline_table_lines(..., _SPLINE_EST_DERIV_) (tools/arrays.c:514-660). */
typedef struct clpp_background_desc {
  int bt_size, bg_size, bg_size_short, bg_size_normal;
  int index_bg_a, index_bg_H, index_bg_H_prime;
  int index_bg_rho_g, index_bg_rho_b, index_bg_rho_cdm, index_bg_rho_ur;
  int index_bg_rho_ncdm1, index_bg_p_ncdm1, index_bg_pseudo_p_ncdm1;
  int has_cdm, has_ur, has_ncdm, N_ncdm, sgnK;
  /* species of the reference that this path does not integrate: must all be 0 */
  int has_fld, has_scf, has_dcdm, has_dr, has_idr, has_idm_dr, has_curvature;
  double conformal_age, a_today, H0, K, h, Omega0_b, T_cmb;
} clpp_background_desc;

int clpp_set_background(clpp_ctx* ctx, const clpp_background_desc* desc,
                        const double* tau_table, const double* background_table, char* err);

/* replaces the reads of ThermodynamicsModule (source/thermodynamics_module.h:13-80) and of
 * its interpolator thermodynamics_at_z (source/thermodynamics_module.cpp:114-285).
 * z_table[tt_size] (growing), thermodynamics_table[tt_size*th_size]. */
typedef struct clpp_thermo_desc {
  int tt_size, th_size;
  int index_th_xe, index_th_rate, index_th_tau_d, index_th_dkappa, index_th_ddkappa, index_th_dddkappa;
  int index_th_exp_m_kappa, index_th_g, index_th_dg, index_th_ddg;
  int index_th_Tb, index_th_wb, index_th_cb2, index_th_dcb2, index_th_ddcb2, index_th_r_d;
  int compute_cb2_derivatives, compute_damping_scale;
  int reio_parametrization; /* enum reionization_parametrization of the reference */
  double z_reionization, YHe, n_e;
  double tau_ini, tau_rec, rs_rec, angular_rescaling, tau_free_streaming, tau_cut;
} clpp_thermo_desc;

int clpp_set_thermo(clpp_ctx* ctx, const clpp_thermo_desc* desc,
                    const double* z_table, const double* thermodynamics_table, char* err);

/* replaces the reads of the NonColdDarkMatter helper (tools/non_cold_dark_matter.h:65-76).
 * Arrays q,w,dlnf0_dlnq are the concatenation over species (q_size[n] entries each). */
int clpp_set_ncdm(clpp_ctx* ctx, int N_ncdm, const int* q_size, const double* q, const double* w,
                  const double* dlnf0_dlnq, const double* M, const double* factor, char* err);

/* ---- stage 1: PerturbationsModule -------------------------------------------------------- */
/* precision/perturbs fields read by the path (include/precisions.h, source/perturbations.h) */
typedef struct clpp_perturb_desc {
  /* struct perturbs */
  int has_cl_cmb_temperature, has_cl_cmb_polarization, has_cl_cmb_lensing_potential;
  int has_pk_matter, has_nl_corrections_based_on_delta_m;
  int gauge; /* 1 = synchronous (only one supported), 0 = newtonian */
  int l_scalar_max;
  double k_max_for_pk, z_max_pk;
  int switch_sw, switch_eisw, switch_lisw, switch_dop, switch_pol;
  double eisw_lisw_split_z;
  double three_ceff2_ur, three_cvis2_ur;
  /* struct precision */
  double k_min_tau0, k_max_tau0_over_l_max, k_step_sub, k_step_super, k_step_transition;
  double k_step_super_reduction, k_per_decade_for_pk, k_per_decade_for_bao, k_bao_center, k_bao_width;
  double start_small_k_at_tau_c_over_tau_h, start_large_k_at_tau_h_over_tau_k;
  double tight_coupling_trigger_tau_c_over_tau_h, tight_coupling_trigger_tau_c_over_tau_k;
  double start_sources_at_tau_c_over_tau_h;
  int tight_coupling_approximation; /* enum tca_method; 5 = compromise_CLASS */
  int l_max_g, l_max_pol_g, l_max_ur, l_max_ncdm;
  double tol_ncdm_initial_w, tol_tau_approx, tol_perturb_integration, perturb_sampling_stepsize;
  double smallest_allowed_variation;
  int radiation_streaming_approximation; /* enum rsa_method; 2 = rsa_MD_with_reio */
  double radiation_streaming_trigger_tau_over_tau_k;
  int ur_fluid_approximation; /* enum ufa_method; 2 = ufa_CLASS */
  double ur_fluid_trigger_tau_over_tau_k;
  int ncdm_fluid_approximation; /* enum ncdmfa_method; 2 = ncdmfa_CLASS */
  double ncdm_fluid_trigger_tau_over_tau_k;
  int evolver; /* enum evolver_type: 0 = rk (Cash-Karp RK45, tools/evolver_rkck.c), 1 = ndf15 (default) */
  double curvature_ini;
  double perturb_integration_stepsize; /* rk only: step = this * min(tau_h, tau_k, tau_c) (precisions.h:226) */
} clpp_perturb_desc;

/* sizes and indices that the reference publishes as PerturbationsModule members */
typedef struct clpp_perturb_info {
  int k_size, k_size_cl, k_size_cmb, tau_size, tp_size, ln_tau_size;
  int index_tp_t0, index_tp_t1, index_tp_t2, index_tp_p, index_tp_delta_m, index_tp_delta_cb,
      index_tp_phi_plus_psi; /* -1 when absent */
  double k_min, k_max;
} clpp_perturb_info;

/* per-k work counters, same meaning as evolver_ndf15's stepstat[6]
 * (tools/evolver_ndf15.cpp:112): steps, failed steps, RHS evaluations, Jacobians,
 * LU factorisations, triangular solves; summed over the approximation intervals. */
typedef struct clpp_kstat {
  int steps, failed, fevals, jacobians, factorizations, solves;
  int intervals, status;
  double tau_ini;
  /* per approximation interval (at most 6): size of the ODE system, accepted steps and the SM
   * clock cycles the mode spent in it (device profile; no reference counterpart) */
  int iv_neq[6], iv_steps[6];
  long long iv_cycles[6];
  long long prof[72]; /* developer profile (builds with -DPT_PROF): cycles per (interval, code section), [6][12] */
} clpp_kstat;

/* replaces perturb_indices_of_perturbs + perturb_get_k_list + perturb_timesampling_for_sources
 * (perturbations_module.cpp:843-2238): host-side, bit-exact grids. */
int clpp_perturb_grids(clpp_ctx* ctx, const clpp_perturb_desc* desc, clpp_perturb_info* info, char* err);
/* replaces the k loop of perturb_init (perturbations_module.cpp:668-717): integrates every
 * k mode on the device (modes [k_begin,k_end) only; pass 0,k_size for all). Sources stay
 * resident on the device for stage 2. */
int clpp_perturb_solve(clpp_ctx* ctx, int k_begin, int k_end, char* err);
/* batched form for parameter sweeps (BASELINE config 5): every k mode of n_ctx cosmologies (all on
 * the same device, same precision settings and species content) in ONE kernel launch, modes
 * issued longest-first across the whole batch. The reference has no counterpart (it runs one
 * Cosmology object at a time); per cosmology the result is identical to clpp_perturb_solve. */
int clpp_perturb_solve_batch(clpp_ctx** ctxs, int n_ctx, char* err);
/* (multi-GPU, one cosmology over several GPUs) integrates an explicit list of modes: cost-balanced partitions
 * of the k grid are interleaved, not contiguous. Same result per mode as clpp_perturb_solve. */
int clpp_perturb_solve_list(clpp_ctx* ctx, const int* k_indices, int n, char* err);
int clpp_perturb_get_k(const clpp_ctx* ctx, double* k /*[k_size]*/);
int clpp_perturb_get_tau(const clpp_ctx* ctx, double* tau /*[tau_size]*/);
/* sources in the reference's layout sources_[index_tp][index_tau*k_size + index_k]
 * (source/perturbations.h:20), concatenated over tp: out[tp][tau][k] */
int clpp_perturb_get_sources(clpp_ctx* ctx, double* out, char* err);
int clpp_perturb_get_kstat(const clpp_ctx* ctx, clpp_kstat* out /*[k_size]*/);
/* test / pipeline hook: inject externally computed grids and sources (same layout) so that
 * stage 2 can be run on the reference's S(k,tau) */
int clpp_perturb_set_sources(clpp_ctx* ctx, const clpp_perturb_info* info, const double* k,
                             const double* tau, const double* sources, char* err);
/* (multi-GPU) raw device pointer and element count of the device-resident source table
 * laid out [tp][k][tau] (tau fastest); ranks all-gather k-slices of it in place. */
int clpp_perturb_device_sources(clpp_ctx* ctx, void** dptr, long* count, char* err);

/* ---- stage 2: TransferModule ------------------------------------------------------------- */
typedef struct clpp_transfer_desc {
  int has_cl_cmb_temperature, has_cl_cmb_polarization, has_cl_cmb_lensing_potential;
  int l_scalar_max;
  double l_logstep, l_linstep;
  double hyper_x_min, hyper_sampling_flat, hyper_phi_min_abs;
  double q_linstep, q_logstep_spline, q_logstep_open;
  double transfer_neglect_delta_k_S_t0, transfer_neglect_delta_k_S_t1, transfer_neglect_delta_k_S_t2,
      transfer_neglect_delta_k_S_e;
  double transfer_neglect_late_source;
  double l_switch_limber;
  double lcmb_rescale, lcmb_tilt, lcmb_pivot;
} clpp_transfer_desc;

typedef struct clpp_transfer_info {
  int tt_size, l_size, l_size_max, q_size;
  int index_tt_t0, index_tt_t1, index_tt_t2, index_tt_e, index_tt_lcmb; /* -1 when absent */
  int x_size; /* rows of the flat Bessel table */
  long n_integrals, n_points; /* non-neglected, non-Limber LOS integrals and their integrand points */
} clpp_transfer_info;

/* replaces transfer_indices_of_transfers + transfer_get_{l,q,k}_list
 * (transfer_module.cpp:402-1167): host-side, bit-exact grids. */
int clpp_transfer_grids(clpp_ctx* ctx, const clpp_transfer_desc* desc, clpp_transfer_info* info, char* err);
/* replaces the rest of transfer_init (transfer_module.cpp:127-345).
 * nl_corr_density: NULL, or NonlinearModule::nl_corr_density_[index_pk_m][tau*k_size+k]
 * (applied to phi+psi as in transfer_module.cpp:559-590).
 * q range [q_begin,q_end) for multi-GPU partitioning; 0,q_size for all. */
int clpp_transfer_compute(clpp_ctx* ctx, const double* nl_corr_density, int q_begin, int q_end, char* err);
int clpp_transfer_get_l(const clpp_ctx* ctx, int* l /*[l_size_max]*/, int* l_size_tt /*[tt_size]*/);
int clpp_transfer_get_q(const clpp_ctx* ctx, double* q /*[q_size]*/, double* k /*[q_size]*/);
/* transfer_[((index_tt*l_size)+index_l)*q_size + index_q]  (source/transfer_module.h:52) */
int clpp_transfer_get_transfer(clpp_ctx* ctx, double* out, char* err);
int clpp_transfer_set_transfer(clpp_ctx* ctx, const double* transfer, char* err);
int clpp_transfer_device_transfer(clpp_ctx* ctx, void** dptr, long* count, char* err);
/* Bessel table accessors for unit tests: phi/dphi[l_size_max][x_size], chi_at_phimin[l_size_max] */
int clpp_transfer_get_bessel(clpp_ctx* ctx, double* x, double* phi, double* dphi, double* chi_at_phimin, char* err);

/* ---- stage 3: SpectraModule -------------------------------------------------------------- */
typedef struct clpp_spectra_info {
  int ct_size, l_size;
  int index_ct_tt, index_ct_ee, index_ct_te, index_ct_bb, index_ct_pp, index_ct_tp, index_ct_ep; /* -1 absent */
} clpp_spectra_info;

/* replaces spectra_indices + spectra_cls + spectra_compute_cl (spectra_module.cpp:527-1353).
 * primordial_pk[q_size] = P_R(k_q) for the adiabatic mode as returned by
 * PrimordialModule::primordial_spectrum_at_k(index_md, linear, k, .) (spectra_module.cpp:996).
 * cl_out[l_size*ct_size] in the reference layout cl_[(index_l*ic_ic_size+0)*ct_size + index_ct]. */
int clpp_spectra_compute(clpp_ctx* ctx, const double* primordial_pk, clpp_spectra_info* info,
                         double* cl_out, char* err);
/* partial sums over q in [q_begin,q_end) only (multi-GPU: ranks all-reduce cl_out) */
int clpp_spectra_compute_range(clpp_ctx* ctx, const double* primordial_pk, int q_begin, int q_end,
                               clpp_spectra_info* info, double* cl_out, char* err);

/* replaces PerturbationsModule::perturb_sources_at_tau (perturbations_module.cpp:79-131; scalars, adiabatic mode) for the
 * default z_max_pk = 0 (ln_tau_size_ <= 1): psource[k_size] = S^{index_tp}(k, tau), linear in tau between the sampling
 * times; fails like array_interpolate_two_bis outside [tau_sampling_[0], tau_sampling_[tau_size-1]]. */
int clpp_perturb_sources_at_tau(clpp_ctx* ctx, int index_tp, double tau, double* psource, char* err);

/* ---- P(k): first "next" row of SURVEY 8f ---------------------------------------------------- */
/* replaces NonlinearModule::nonlinear_pk_linear (nonlinear_module.cpp:1886-2024, adiabatic mode):
 * pk_out[i] = 2 pi^2 / k_i^3 * primordial_pk[i] * delta(k_i, tau)^2 on the perturbation k grid, read from the
 * device-resident sources. primordial_pk[k_size] = P_R(k_i) (primordial_spectrum_at_k, linear);
 * index_tau < 0: today (last sample); cb = 0: total matter (delta_m), 1: cdm+baryons (delta_cb). */
int clpp_pk_linear(clpp_ctx* ctx, const double* primordial_pk, int index_tau, int cb, double* pk_out, char* err);

/* replaces the halofit branch of NonlinearModule::nonlinear_init (nonlinear_module.cpp:1228-1420) with
 * nonlinear_halofit (:2291-2726): R_NL(k,tau) = sqrt(P_NL/P_L) from the device-resident delta_m / delta_cb sources.
 * The result stays on the device and is applied by the next clpp_transfer_compute called with nl_corr_density = NULL
 * (it is invalidated by a new clpp_perturb_solve).  nl_corr_out: NULL, or [tau_size*k_size] in the reference layout
 * nl_corr_density_[index_pk_m][tau*k_size+k]; index_tau_min_nl: NULL or the reference's index_tau_min_nl_. */
typedef struct clpp_halofit_desc {
  double halofit_min_k_nonlinear, halofit_k_per_decade, halofit_sigma_precision, halofit_tol_sigma; /* precisions.h:432-449 */
} clpp_halofit_desc;
int clpp_nonlinear_halofit(clpp_ctx* ctx, const clpp_halofit_desc* desc, const double* primordial_pk /*[k_size]*/,
                           double* nl_corr_out, int* index_tau_min_nl, char* err);

/* ---- lensed C_l: second "next" row of SURVEY 8f ---------------------------------------------- */
/* replaces LensingModule::lensing_init (lensing_module.cpp:149-860) with lensing_indices (:886-1092): lensed
 * TT, TE, EE, BB by the full-sky correlation-function method from the table of clpp_spectra_compute (which must hold
 * tt and pp). Fast mode (accurate_lensing = 0, the default of precisions.h:492) or Gauss-Legendre mode.
 * l_out[info.l_size] (NULL allowed): multipoles l_; cl_lens_out[info.l_size*info.lt_size] (NULL allowed): cl_lens_
 * in the reference layout [index_l*lt_size + index_lt]; types other than tt/te/ee/bb are copied unlensed. */
typedef struct clpp_lensing_desc {
  int accurate_lensing, delta_l_max, num_mu_minus_lmax; /* precisions.h:492-494 */
  double tol_gauss_legendre;                              /* precisions.h:495 */
} clpp_lensing_desc;
typedef struct clpp_lensing_info {
  int lt_size, l_size, l_unlensed_max, l_lensed_max;
  int index_lt_tt, index_lt_ee, index_lt_te, index_lt_bb, index_lt_pp, index_lt_tp, index_lt_ep; /* -1 absent */
} clpp_lensing_info;
int clpp_lensing_compute(clpp_ctx* ctx, const clpp_lensing_desc* desc, clpp_lensing_info* info, double* l_out,
                         double* cl_lens_out, char* err);
/* replaces LensingModule::lensing_cl_at_l (lensing_module.cpp:111-140): spline in l through cl_lens_, fails above
 * l_lensed_max with the reference's message; cl_lensed[lt_size]. */
int clpp_lensing_cl_at_l(const clpp_ctx* ctx, int l, double* cl_lensed, char* err);

/* replaces SpectraModule::spectra_cl_at_l (spectra_module.cpp:220-264, one mode / one initial condition):
 * cubic spline in l through the table of clpp_spectra_compute, zero above l_scalar_max. cl_tot[ct_size]. */
int clpp_spectra_cl_at_l(const clpp_ctx* ctx, double l, double* cl_tot, char* err);
/* replaces SpectraModule::cl_output (spectra_module.cpp:146-198): out[(lmax+1)*ct_size], row l = C_l^{ct}
 * (dimensionless), rows l = 0, 1 are zero; fails like the reference when lmax is outside [0, l_max_tot]. */
int clpp_spectra_cl_output(const clpp_ctx* ctx, int lmax, double* out, char* err);

#ifdef __cplusplus
}
#endif
#endif /* CLPP_H */
