// Halofit on the device: the non-linear correction R_NL(k,tau) = sqrt(P_NL/P_L) that multiplies the phi+psi source
// before the transfer stage (transfer_module.cpp:559-590).  First "next" row of SURVEY 8f: with `non linear =
// halofit` (BASELINE config 2) the reference's NonlinearModule sits BETWEEN stage 1 and stage 2 of the hot path
// (0.25 s of serial CPU time per cosmology); here it runs on the device-resident delta_m sources.
//
// Reference restated: NonlinearModule::nonlinear_init, halofit branch (source/nonlinear_module.cpp:1228-1420),
// nonlinear_pk_linear (:1886-2024), nonlinear_halofit (:2291-2726: Takahashi et al. 2012 + Bird et al. 2011),
// nonlinear_halofit_integrate (:2740-2810), array_spline / array_spline_table_columns (natural),
// array_interpolate_spline, array_integrate_all_spline (tools/arrays.c:315-420, 1354-1380 -- including its
// "+ h^3/24" sign).
//
// B200 mapping: one warp per (tau, spectrum); the k-grid spline, the 80-points-per-decade integrand grid and the
// sigma(R) integrals live in shared memory; lanes fill integrands / reduce integrals, the two tridiagonal sweeps of
// every spline are sequential (x-only coefficients are precomputed once per warp).  704 x 2 warps, ~1 ms.
#include <algorithm>
#include <cmath>
#include <vector>

#include "device.h"

struct HfParams {
  const double *k, *tau, *sources, *primordial;  // sources [tp][k][tau]
  const double *bg_tau, *bg_y, *bg_dd;
  int bt_size, bg_size;
  int ia, iH, irho_g, irho_b, irho_cdm, irho_ur, irho_ncdm1, ip_ncdm1, has_ur, N_ncdm;
  int nk, nt, ni;  // k grid, tau grid, integrand grid
  int tp_m, tp_cb, n_spec;
  double min_k_nonlinear, k_per_decade, sigma_precision, tol_sigma;
  double fnu_m, Omega0_m, h;
  double* corr;  // [spec][k][tau]
  int* fail;     // [spec][tau]
  int* status;   // != 0: error (1: sigma(R_max) > 1, 2: bisection did not converge)
};

__device__ __forceinline__ double hf_wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// natural cubic spline second derivatives of y(x), sequential (lane 0); x-only coefficients in sg / ip
__device__ void hf_spline_natural(int n, const double* x, const double* y, double* dd, double* u) {
  dd[0] = 0.;
  u[0] = 0.;
  for (int i = 1; i < n - 1; i++) {
    const double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
    const double p = sig * dd[i - 1] + 2.0;
    dd[i] = (sig - 1.0) / p;
    double t = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
    u[i] = (6.0 * t / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
  }
  dd[n - 1] = 0.;
  for (int k = n - 2; k >= 0; k--) dd[k] = dd[k] * dd[k + 1] + u[k];
}

__global__ void __launch_bounds__(32) halofit_kernel(const HfParams P) {
  extern __shared__ double hs[];
  const int lane = threadIdx.x;
  const int it = blockIdx.x % P.nt, spec = blockIdx.x / P.nt;
  const int nk = P.nk, ni = P.ni, nt = P.nt;
  double* lnk = hs;             // [nk]
  double* lnpk = lnk + nk;      // [nk]
  double* ddlnpk = lnpk + nk;   // [nk]
  double* ki = ddlnpk + nk;     // [ni] integrand grid
  double* pki = ki + ni;        // [ni]
  double* f = pki + ni;         // [ni]
  double* dd = f + ni;          // [ni]
  double* sgi = dd + ni;        // [ni] x-only spline coefficients of the integrand grid: sig,
  double* pin = sgi + ni;       // [ni]   1/p,
  double* cdd = pin + ni;       // [ni]   (sig-1)/p
  double* u = cdd + ni;         // [max(ni,nk)]
  const int tp = spec == 0 ? P.tp_m : P.tp_cb;
  const double fnu = spec == 0 ? P.fnu_m : 0.;
  const double* src = P.sources + (size_t)tp * nk * nt;
  const double tau = P.tau[it];
  const double PI = CLPP_PI;

  // ---- linear spectrum at this time and its spline along ln k (nonlinear_pk_linear + array_spline_table_columns)
  for (int i = lane; i < nk; i += 32) {
    const double kk = P.k[i], s = src[(size_t)i * nt + it];
    lnk[i] = log(kk);
    lnpk[i] = log(2. * PI * PI / (kk * kk * kk) * s * s * P.primordial[i]);
  }
  __syncwarp();
  if (lane == 0) hf_spline_natural(nk, lnk, lnpk, ddlnpk, u);
  __syncwarp();
  // ---- Omega_m(tau), Omega_v(tau) from the background table (cubic spline, bisection): rho_m / H^2 etc.
  double Omega_m, Omega_v;
  {
    int lo = 0, hi = P.bt_size - 1;
    while (hi - lo > 1) {
      const int mid = (int)(0.5 * (lo + hi));
      if (tau < P.bg_tau[mid]) hi = mid; else lo = mid;
    }
    const double x0 = P.bg_tau[lo], x1 = P.bg_tau[hi], hh = x1 - x0, b = (tau - x0) / hh, a = 1 - b;
    auto col = [&](int c) {
      const size_t r0 = (size_t)lo * P.bg_size + c, r1 = (size_t)hi * P.bg_size + c;
      return a * P.bg_y[r0] + b * P.bg_y[r1] + ((a * a * a - a) * P.bg_dd[r0] + (b * b * b - b) * P.bg_dd[r1]) * hh * hh / 6.;
    };
    const double H = col(P.iH);
    double rho_m = col(P.irho_b) + col(P.irho_cdm), rho_r = col(P.irho_g);
    if (P.has_ur) rho_r += col(P.irho_ur);
    for (int s = 0; s < P.N_ncdm; s++) {
      const double rn = col(P.irho_ncdm1 + s), pn = col(P.ip_ncdm1 + s);
      rho_r += 3. * pn;
      rho_m += rn - 3. * pn;
    }
    Omega_m = rho_m / (H * H);
    Omega_v = 1. - Omega_m - rho_r / (H * H);
  }
  const double w0 = -1.;  // no fluid dark energy on this path (has_fld is rejected): w_fld(today) = -1
  // ---- integrand grid: k_n = k_0 10^(n / k_per_decade), P(k_n) by spline interpolation in ln k
  for (int i = lane; i < ni; i += 32) {
    const double kk = P.k[0] * pow(10., i / P.k_per_decade);
    double lp;
    if (i == 0) lp = lnpk[0];
    else {
      const double x = log(kk);
      int lo = 0, hi = nk - 1;
      while (hi - lo > 1) {
        const int mid = (int)(0.5 * (lo + hi));
        if (x < lnk[mid]) hi = mid; else lo = mid;
      }
      const double hh = lnk[hi] - lnk[lo], b = (x - lnk[lo]) / hh, a = 1 - b;
      lp = a * lnpk[lo] + b * lnpk[hi] + ((a * a * a - a) * ddlnpk[lo] + (b * b * b - b) * ddlnpk[hi]) * hh * hh / 6.;
    }
    ki[i] = kk;
    pki[i] = exp(lp);
  }
  __syncwarp();
  // the abscissae of the sigma(R) integrals never change: factor the tridiagonal matrix of their natural spline once
  if (lane == 0) {
    cdd[0] = 0.;
    for (int i = 1; i < ni - 1; i++) {
      const double sig = (ki[i] - ki[i - 1]) / (ki[i + 1] - ki[i - 1]);
      const double p = sig * cdd[i - 1] + 2.0;
      sgi[i] = sig;
      pin[i] = 1.0 / p;
      cdd[i] = (sig - 1.0) / p;
    }
  }
  __syncwarp();
  const double anorm = 1. / (2 * PI * PI);
  // sigma^2-type integral at radius R (nonlinear_halofit_integrate): type 1, 2, 3
  auto integrate = [&](double R, int type) {
    for (int i = lane; i < ni; i += 32) {
      const double kk = ki[i], x2 = kk * kk * R * R;
      double v = pki[i] * kk * kk * anorm * exp(-x2);
      if (type == 2) v *= 2. * x2;
      if (type == 3) v *= 4. * x2 * (1. - x2);
      f[i] = v;
    }
    __syncwarp();
    // natural spline of f on the integrand grid: right-hand sides in parallel, the two sweeps on lane 0
    for (int i = 1 + lane; i < ni - 1; i += 32) {
      const double t = (f[i + 1] - f[i]) / (ki[i + 1] - ki[i]) - (f[i] - f[i - 1]) / (ki[i] - ki[i - 1]);
      u[i] = 6.0 * t / (ki[i + 1] - ki[i - 1]);
    }
    __syncwarp();
    if (lane == 0) {
      double up = 0.;
      for (int i = 1; i < ni - 1; i++) {
        up = (u[i] - sgi[i] * up) * pin[i];
        u[i] = up;
      }
      double dn = 0.;
      dd[ni - 1] = 0.;
      for (int k = ni - 2; k >= 1; k--) {
        dn = cdd[k] * dn + u[k];
        dd[k] = dn;
      }
      dd[0] = 0.;
    }
    __syncwarp();
    double s = 0.;
    for (int i = lane; i < ni - 1; i += 32) {
      const double hh = ki[i + 1] - ki[i];
      s += (f[i] + f[i + 1]) * hh / 2. + (dd[i] + dd[i + 1]) * hh * hh * hh / 24.;
    }
    s = hf_wsum(s);
    __syncwarp();
    return s;
  };
  double* out = P.corr + ((size_t)spec * nk) * nt + it;  // element (k, tau) at out[k*nt]
  int* failp = P.fail + spec * nt + it;
  double R = sqrt(-log(P.sigma_precision)) / ki[ni - 1];
  double sum1 = integrate(R, 1);
  double sigma = sqrt(sum1);
  if (sigma < 1.) {  // k_max too small to find the non-linear scale at this redshift: R_NL = 1
    for (int i = lane; i < nk; i += 32) out[(size_t)i * nt] = 1.;
    if (lane == 0) *failp = 1;
    return;
  }
  if (lane == 0) *failp = 0;
  double xlogr1 = log(R) / log(10.);
  R = 1. / P.min_k_nonlinear;
  sum1 = integrate(R, 1);
  sigma = sqrt(sum1);
  if (sigma > 1.) { if (lane == 0) atomicMax(P.status, 1); return; }
  double xlogr2 = log(R) / log(10.);
  int counter = 0;
  double rmid, diff;
  do {
    rmid = pow(10, (xlogr2 + xlogr1) / 2.0);
    counter++;
    sum1 = integrate(rmid, 1);
    sigma = sqrt(sum1);
    diff = sigma - 1.0;
    if (diff > P.tol_sigma) xlogr1 = log10(rmid);
    else if (diff < -P.tol_sigma) xlogr2 = log10(rmid);
    if (counter > 10000) { if (lane == 0) atomicMax(P.status, 2); return; }
  } while (fabs(diff) > P.tol_sigma);
  const double sum2 = integrate(rmid, 2);
  const double sum3 = integrate(rmid, 3);
  const double d1 = -sum2 / sum1;
  const double d2 = -sum2 * sum2 / sum1 / sum1 - sum3 / sum1;
  const double rknl = 1. / rmid, rneff = -3. - d1, rncur = -d2;
  // ---- fitting formula (identical for every k: hoist the k-independent coefficients)
  const double gam = 0.1971 - 0.0843 * rneff + 0.8460 * rncur;
  double a = 1.5222 + 2.8553 * rneff + 2.3706 * rneff * rneff + 0.9903 * rneff * rneff * rneff +
             0.2250 * rneff * rneff * rneff * rneff - 0.6038 * rncur + 0.1749 * Omega_v * (1. + w0);
  a = pow(10, a);
  const double b = pow(10, (-0.5642 + 0.5864 * rneff + 0.5716 * rneff * rneff - 1.5474 * rncur + 0.2279 * Omega_v * (1. + w0)));
  const double c = pow(10, 0.3698 + 2.0404 * rneff + 0.8161 * rneff * rneff + 0.5869 * rncur);
  const double xmu = 0.;
  const double xnu = pow(10, 5.2105 + 3.6902 * rneff);
  const double alpha = fabs(6.0835 + 1.3373 * rneff - 0.1959 * rneff * rneff - 5.5274 * rncur);
  const double beta = 2.0379 - 0.7354 * rneff + 0.3157 * pow(rneff, 2) + 1.2490 * pow(rneff, 3) + 0.3980 * pow(rneff, 4) -
                      0.1682 * rncur + fnu * (1.081 + 0.395 * pow(rneff, 2));
  double f1, f2, f3;
  if (fabs(1 - Omega_m) > 0.01) {
    const double f1a = pow(Omega_m, (-0.0732)), f2a = pow(Omega_m, (-0.1423)), f3a = pow(Omega_m, (0.0725));
    const double f1b = pow(Omega_m, (-0.0307)), f2b = pow(Omega_m, (-0.0585)), f3b = pow(Omega_m, (0.0743));
    const double frac = Omega_v / (1. - Omega_m);
    f1 = frac * f1b + (1 - frac) * f1a;
    f2 = frac * f2b + (1 - frac) * f2a;
    f3 = frac * f3b + (1 - frac) * f3a;
  } else {
    f1 = 1.; f2 = 1.; f3 = 1.;
  }
  for (int i = lane; i < nk; i += 32) {
    const double rk = P.k[i];
    double pk_nl;
    const double pl = exp(lnpk[i]);
    if (rk > P.min_k_nonlinear) {
      const double pk_lin = pl * rk * rk * rk * anorm;
      const double y = rk / rknl;
      double pk_halo = a * pow(y, f1 * 3.) / (1. + b * pow(y, f2) + pow(f3 * c * y, 3. - gam));
      pk_halo = pk_halo / (1 + xmu * pow(y, -1) + xnu * pow(y, -2)) * (1 + fnu * (0.977 - 18.015 * (P.Omega0_m - 0.3)));
      const double rkh2 = (rk / P.h) * (rk / P.h);
      const double pk_linaa = pk_lin * (1 + fnu * 47.48 * rkh2 / (1 + 1.5 * rkh2));
      const double pk_quasi = pk_lin * pow((1 + pk_linaa), beta) / (1 + pk_linaa * alpha) * exp(-y / 4.0 - y * y / 8.0);
      pk_nl = (pk_halo + pk_quasi) / (rk * rk * rk) / anorm;
    } else {
      pk_nl = pl;
    }
    out[(size_t)i * nt] = sqrt(pk_nl / pl);
  }
}

// R_NL = 1 at and before the first time (going backwards) at which either spectrum could not be computed
__global__ void halofit_mask_kernel(int nk, int nt, int n_spec, int i_fail, double* corr) {
  const size_t n = (size_t)n_spec * nk * nt;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (int)(i % nt) <= i_fail) corr[i] = 1.;
}

int clpp_dev_halofit(clpp_ctx* c, const clpp_halofit_desc* hd, const double* primordial_pk, double* nl_corr_out,
                     int* index_tau_min_nl, char* err) {
  clpp_ctx::Dev* d = c->dev;
  const clpp_perturb_info& I = c->pinfo;
  const clpp_background_desc& bg = c->bg;
  cudaStream_t st = d->stream;
  const int nk = I.k_size, nt = I.tau_size;
  CLPP_CHECK(I.index_tp_delta_m >= 0, err, "halofit needs the delta_m source (has_nl_corrections_based_on_delta_m)");
  CLPP_CHECK(!bg.has_fld, err, "halofit on the B200 path assumes w = -1 (no fluid dark energy)");
  HfParams P;
  memset(&P, 0, sizeof(P));
  P.nk = nk; P.nt = nt;
  P.ni = (int)(log(c->k[nk - 1] / c->k[0]) / log(10.) * hd->halofit_k_per_decade) + 1;
  CLPP_CHECK(P.ni >= 3, err, "halofit integrand grid too small");
  P.tp_m = I.index_tp_delta_m; P.tp_cb = I.index_tp_delta_cb;
  P.n_spec = (I.index_tp_delta_cb >= 0) ? 2 : 1;
  P.min_k_nonlinear = hd->halofit_min_k_nonlinear; P.k_per_decade = hd->halofit_k_per_decade;
  P.sigma_precision = hd->halofit_sigma_precision; P.tol_sigma = hd->halofit_tol_sigma;
  P.h = bg.h;
  // Omega0_m and Omega0_ncdm_tot from the last row of the background table (background_module.cpp:1315)
  {
    const HostTable& t = c->bgt;
    const double* row = &t.y[(size_t)(t.n_lines - 1) * t.n_cols];
    const double H0sq = row[bg.index_bg_H] * row[bg.index_bg_H];
    double rho_m = row[bg.index_bg_rho_b] + row[bg.index_bg_rho_cdm], rho_ncdm = 0.;
    for (int s = 0; s < (bg.has_ncdm ? bg.N_ncdm : 0); s++) {
      const double rn = row[bg.index_bg_rho_ncdm1 + s], pn = row[bg.index_bg_p_ncdm1 + s];
      rho_m += rn - 3. * pn;
      rho_ncdm += rn;
    }
    P.Omega0_m = rho_m / H0sq;
    P.fnu_m = (rho_ncdm / H0sq) / P.Omega0_m;
  }
  P.bg_tau = d->bg_tau; P.bg_y = d->bg_y; P.bg_dd = d->bg_dd; P.bt_size = bg.bt_size; P.bg_size = bg.bg_size;
  P.ia = bg.index_bg_a; P.iH = bg.index_bg_H; P.irho_g = bg.index_bg_rho_g; P.irho_b = bg.index_bg_rho_b;
  P.irho_cdm = bg.index_bg_rho_cdm; P.irho_ur = bg.index_bg_rho_ur; P.irho_ncdm1 = bg.index_bg_rho_ncdm1;
  P.ip_ncdm1 = bg.index_bg_p_ncdm1; P.has_ur = bg.has_ur; P.N_ncdm = bg.has_ncdm ? bg.N_ncdm : 0;
  if (clpp_dev_reserve(d, &d->pk, (size_t)2 * nk, err)) return CLPP_FAILURE;
  CLPP_CUDA(cudaMemcpyAsync(d->pk, primordial_pk, nk * sizeof(double), cudaMemcpyHostToDevice, st), err);
  if (clpp_dev_reserve(d, &d->nl_corr2, (size_t)P.n_spec * nk * nt, err)) return CLPP_FAILURE;
  if (clpp_dev_reserve(d, &d->hf_flags, (size_t)2 * nt + 1, err)) return CLPP_FAILURE;
  CLPP_CUDA(cudaMemsetAsync(d->hf_flags, 0, ((size_t)2 * nt + 1) * sizeof(int), st), err);
  P.k = d->k; P.tau = d->tau; P.sources = d->sources; P.primordial = d->pk;
  P.corr = d->nl_corr2; P.fail = d->hf_flags; P.status = d->hf_flags + 2 * nt;
  const size_t smem = (size_t)(3 * nk + 7 * P.ni + std::max(P.ni, nk)) * sizeof(double);
  CLPP_CHECK(smem <= 200 * 1024, err, "k grid too large for the shared-memory staging of halofit");
  { static const cudaError_t once = clpp_allow_max_dynamic_smem(halofit_kernel); CLPP_CUDA(once, err); }
  cudaEventRecord(d->ev[0], st);
  halofit_kernel<<<P.n_spec * nt, 32, smem, st>>>(P);
  c->launches++;
  CLPP_CUDA(cudaGetLastError(), err);
  std::vector<int> flags(2 * nt + 1);
  CLPP_CUDA(cudaMemcpyAsync(flags.data(), d->hf_flags, flags.size() * sizeof(int), cudaMemcpyDeviceToHost, st), err);
  CLPP_CUDA(cudaStreamSynchronize(st), err);
  CLPP_CHECK(flags[2 * nt] != 1, err,
             "Your input value for the precision parameter halofit_min_k_nonlinear=%e is too large, such that "
             "sigma(R=1/halofit_min_k_nonlinear) > 1. For self-consistency, it should have been <1. Decrease "
             "halofit_min_k_nonlinear", hd->halofit_min_k_nonlinear);
  CLPP_CHECK(flags[2 * nt] != 2, err, "could not converge within maximum allowed number of iterations");
  int i_fail = -1;
  for (int s = 0; s < P.n_spec; s++)
    for (int it = 0; it < nt; it++)
      if (flags[s * nt + it]) i_fail = std::max(i_fail, it);
  if (i_fail >= 0) {
    const size_t n = (size_t)P.n_spec * nk * nt;
    halofit_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(nk, nt, P.n_spec, i_fail, d->nl_corr2);
    c->launches++;
  }
  cudaEventRecord(d->ev[1], st);
  if (index_tau_min_nl) *index_tau_min_nl = std::min(nt - 1, i_fail + 1);
  if (nl_corr_out) {  // reference layout [tau][k] of the total-matter correction
    std::vector<double> tmp((size_t)nk * nt);
    CLPP_CUDA(cudaMemcpyAsync(tmp.data(), d->nl_corr2, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, st), err);
    CLPP_CUDA(cudaStreamSynchronize(st), err);
    for (int ik = 0; ik < nk; ik++)
      for (int it = 0; it < nt; it++) nl_corr_out[(size_t)it * nk + ik] = tmp[(size_t)ik * nt + it];
  }
  CLPP_CUDA(cudaStreamSynchronize(st), err);
  { float ms = 0; cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]); d->t_halofit_ms = ms; }
  c->nl_dev_valid = true;
  c->nl_dev_nk = nk; c->nl_dev_nt = nt;
  return CLPP_SUCCESS;
}
