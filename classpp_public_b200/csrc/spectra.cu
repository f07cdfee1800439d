// Stage 3: C_l k-quadrature on the device.
//
// Reference: SpectraModule::spectra_compute_cl (source/spectra_module.cpp:958-1353) builds, for
// every l, the integrand f_ct(q) = 4 pi / k * P_R(k) * Delta^X_l(q) Delta^Y_l(q), splines it in
// k (array_spline, tools/arrays.c:315-420, _SPLINE_EST_DERIV_) and integrates it with
// array_integrate_all_trapzd_or_spline (tools/arrays.c:1382-1427):
//     C = sum_i (f_i + f_{i+1}) h_i / 2 + (f''_i + f''_{i+1}) h_i^3 / 24 .
// Both steps are linear in f on a fixed grid, so the whole quadrature is one fixed linear
// functional  C = sum_q w_q f_q.  We compute w once per cosmology on the host (adjoint
// tridiagonal solve, O(q_size)) and the device does a batched contraction that streams
// Delta_l(q) exactly once:  C_l^{ct} = sum_q W_q * pair_ct(Delta_l(q)),  W_q = w_q 4pi P(k_q)/k_q.
// This is HBM-bound: algorithmic bytes = tt_size*l_size*q_size*8 (the transfer table).
#include <cmath>

#include "device.h"

// ---- host: quadrature weights --------------------------------------------------------------
// w such that sum_q w_q f_q reproduces spline(EST_DERIV) + integrate of the reference.
static void spline_integration_weights(const std::vector<double>& xin, std::vector<double>& wout) {
  typedef long double R;
  const int n = (int)xin.size();
  std::vector<R> x(xin.begin(), xin.end()), h(n - 1);
  for (int i = 0; i < n - 1; i++) h[i] = x[i + 1] - x[i];
  // trapezoid part t and the coefficient c of the second derivatives m_i
  std::vector<R> w(n, 0), c(n, 0);
  for (int i = 0; i < n - 1; i++) {
    w[i] += h[i] / 2;
    w[i + 1] += h[i] / 2;
    c[i] += h[i] * h[i] * h[i] / 24;
    c[i + 1] += h[i] * h[i] * h[i] / 24;
  }
  // spline system T m = B f.  Rows of T: (lo, di, up)
  std::vector<R> lo(n, 0), di(n, 2), up(n, 0);
  up[0] = 1;
  for (int i = 1; i < n - 1; i++) {
    R sig = h[i - 1] / (x[i + 1] - x[i - 1]);
    lo[i] = sig;
    up[i] = 1 - sig;
  }
  lo[n - 1] = 1;
  // adjoint solve T^T z = c.  T^T has sub-diagonal up[i-1], diagonal di[i], super-diagonal lo[i+1].
  std::vector<R> z(n), cp(n), dp(n);
  {
    cp[0] = lo[1] / di[0];
    dp[0] = c[0] / di[0];
    for (int i = 1; i < n; i++) {
      R sub = up[i - 1];
      R sup = (i < n - 1) ? lo[i + 1] : 0;
      R den = di[i] - sub * cp[i - 1];
      cp[i] = sup / den;
      dp[i] = (c[i] - sub * dp[i - 1]) / den;
    }
    z[n - 1] = dp[n - 1];
    for (int i = n - 2; i >= 0; i--) z[i] = dp[i] - cp[i] * z[i + 1];
  }
  // w += B^T z
  {
    // first row: (6/h0) [ (f1-f0)/h0 - yp0 ],  yp0 = A1 (f1-f0) - A2 (f2-f0)
    R D0 = (x[2] - x[0]) * (x[1] - x[0]) * (x[2] - x[1]);
    R A1 = (x[2] - x[0]) * (x[2] - x[0]) / D0, A2 = (x[1] - x[0]) * (x[1] - x[0]) / D0;
    R g = 6 / h[0];
    w[0] += z[0] * g * (-1 / h[0] + A1 - A2);
    w[1] += z[0] * g * (1 / h[0] - A1);
    w[2] += z[0] * g * A2;
    for (int i = 1; i < n - 1; i++) {
      R gi = 6 / (x[i + 1] - x[i - 1]);
      w[i + 1] += z[i] * gi / h[i];
      w[i] += z[i] * (-gi / h[i] - gi / h[i - 1]);
      w[i - 1] += z[i] * gi / h[i - 1];
    }
    // last row: (6/h) [ ypn - (f_{n-1}-f_{n-2})/h ],  ypn = C1 (f_{n-2}-f_{n-1}) - C2 (f_{n-3}-f_{n-1})
    R hn = h[n - 2];
    R Dn = (x[n - 3] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (x[n - 3] - x[n - 2]);
    R C1 = (x[n - 3] - x[n - 1]) * (x[n - 3] - x[n - 1]) / Dn, C2 = (x[n - 2] - x[n - 1]) * (x[n - 2] - x[n - 1]) / Dn;
    R gn = 6 / hn;
    w[n - 2] += z[n - 1] * gn * (C1 + 1 / hn);
    w[n - 1] += z[n - 1] * gn * (-C1 + C2 - 1 / hn);
    w[n - 3] += z[n - 1] * gn * (-C2);
  }
  wout.resize(n);
  for (int i = 0; i < n; i++) wout[i] = (double)w[i];
}

// ---- device ----------------------------------------------------------------------------------
struct SpectraParams {
  int q_size, l_size, tt_size, ct_size;
  int q_begin, q_end, n_chunk;
  int tt_t0, tt_t1, tt_t2, tt_e, tt_lcmb;
  int ct_tt, ct_ee, ct_te, ct_bb, ct_pp, ct_tp, ct_ep;
};

#define SPECTRA_MAX_CT 7
#define SPECTRA_THREADS 256

// grid (l_size, n_chunk): each block contracts one q-chunk of one multipole; coalesced, vectorisable
// FP64 loads along q (the fastest index of transfer_).  partial[l][chunk][ct].
__global__ void __launch_bounds__(SPECTRA_THREADS)
spectra_partial_kernel(SpectraParams P, const double* __restrict__ transfer, const double* __restrict__ W,
                       double* __restrict__ partial) {
  const int il = blockIdx.x, chunk = blockIdx.y;
  const int span = (P.q_end - P.q_begin + P.n_chunk - 1) / P.n_chunk;
  const int q0 = P.q_begin + chunk * span;
  const int q1 = min(q0 + span, P.q_end);
  double acc[SPECTRA_MAX_CT];
#pragma unroll
  for (int i = 0; i < SPECTRA_MAX_CT; i++) acc[i] = 0.;
  const size_t stride_tt = (size_t)P.l_size * P.q_size;
  const double* base = transfer + (size_t)il * P.q_size;
  for (int iq = q0 + threadIdx.x; iq < q1; iq += SPECTRA_THREADS) {
    const double w = W[iq];
    double T = 0., E = 0., Pp = 0.;
    if (P.tt_t0 >= 0)
      T = __ldg(base + P.tt_t0 * stride_tt + iq) + __ldg(base + P.tt_t1 * stride_tt + iq) +
          __ldg(base + P.tt_t2 * stride_tt + iq);
    if (P.tt_e >= 0) E = __ldg(base + P.tt_e * stride_tt + iq);
    if (P.tt_lcmb >= 0) Pp = __ldg(base + P.tt_lcmb * stride_tt + iq);
    acc[0] += w * T * T;
    acc[1] += w * E * E;
    acc[2] += w * T * E;
    acc[3] += w * Pp * Pp;
    acc[4] += w * T * Pp;
    acc[5] += w * E * Pp;
  }
  // warp-shuffle tree then one smem stage
  __shared__ double red[SPECTRA_THREADS / 32][6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0.;
    for (int wdx = 0; wdx < SPECTRA_THREADS / 32; wdx++) v += red[wdx][threadIdx.x];
    partial[((size_t)il * P.n_chunk + chunk) * 6 + threadIdx.x] = v;
  }
}

// fixed-order reduction over chunks and scatter into the reference's ct order
__global__ void spectra_final_kernel(SpectraParams P, const double* __restrict__ partial, double* __restrict__ cl) {
  const int il = blockIdx.x * blockDim.x + threadIdx.x;
  if (il >= P.l_size) return;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int c = 0; c < P.n_chunk; c++)
    for (int i = 0; i < 6; i++) s[i] += partial[((size_t)il * P.n_chunk + c) * 6 + i];
  double* out = cl + (size_t)il * P.ct_size;
  for (int i = 0; i < P.ct_size; i++) out[i] = 0.;  // C_l^BB of scalars is identically zero
  if (P.ct_tt >= 0) out[P.ct_tt] = s[0];
  if (P.ct_ee >= 0) out[P.ct_ee] = s[1];
  if (P.ct_te >= 0) out[P.ct_te] = s[2];
  if (P.ct_pp >= 0) out[P.ct_pp] = s[3];
  if (P.ct_tp >= 0) out[P.ct_tp] = s[4];
  if (P.ct_ep >= 0) out[P.ct_ep] = s[5];
}

int clpp_dev_spectra(clpp_ctx* c, const double* primordial_pk, int q_begin, int q_end, clpp_spectra_info* info,
                     double* cl_out, char* err) {
  clpp_ctx::Dev* d = c->dev;
  const clpp_transfer_info& T = c->tinfo;
  const bool has_t = T.index_tt_t0 >= 0, has_e = T.index_tt_e >= 0, has_p = T.index_tt_lcmb >= 0;
  clpp_spectra_info I;
  int ct = 0;
  I.index_ct_tt = has_t ? ct++ : -1;
  I.index_ct_ee = has_e ? ct++ : -1;
  I.index_ct_te = (has_t && has_e) ? ct++ : -1;
  I.index_ct_bb = has_e ? ct++ : -1;
  I.index_ct_pp = has_p ? ct++ : -1;
  I.index_ct_tp = (has_t && has_p) ? ct++ : -1;
  I.index_ct_ep = (has_e && has_p) ? ct++ : -1;
  I.ct_size = ct;
  I.l_size = T.l_size;
  if (info) *info = I;

  // W_q = w_q * 4 pi / k_q * P(k_q)   (spectra_module.cpp:1136: factor = 4 pi / k)
  std::vector<double> w;
  spline_integration_weights(c->kq, w);
  std::vector<double> W(T.q_size);
  for (int i = 0; i < T.q_size; i++) W[i] = w[i] * (4. * CLPP_PI / c->kq[i]) * primordial_pk[i];
  if (clpp_dev_reserve(d, &d->wq, (size_t)T.q_size, err)) return CLPP_FAILURE;
  CLPP_CUDA(cudaMemcpyAsync(d->wq, W.data(), T.q_size * sizeof(double), cudaMemcpyHostToDevice, d->stream), err);

  SpectraParams P;
  P.q_size = T.q_size; P.l_size = T.l_size; P.tt_size = T.tt_size; P.ct_size = I.ct_size;
  P.q_begin = q_begin; P.q_end = q_end;
  P.n_chunk = 4;
  P.tt_t0 = T.index_tt_t0; P.tt_t1 = T.index_tt_t1; P.tt_t2 = T.index_tt_t2; P.tt_e = T.index_tt_e;
  P.tt_lcmb = T.index_tt_lcmb;
  P.ct_tt = I.index_ct_tt; P.ct_ee = I.index_ct_ee; P.ct_te = I.index_ct_te; P.ct_bb = I.index_ct_bb;
  P.ct_pp = I.index_ct_pp; P.ct_tp = I.index_ct_tp; P.ct_ep = I.index_ct_ep;

  if (clpp_dev_reserve(d, &d->cl, (size_t)T.l_size * I.ct_size + (size_t)T.l_size * P.n_chunk * 6, err)) return CLPP_FAILURE;
  double* partial = d->cl + (size_t)T.l_size * I.ct_size;
  dim3 grid(T.l_size, P.n_chunk);
  cudaEventRecord(d->ev[0], d->stream);
  spectra_partial_kernel<<<grid, SPECTRA_THREADS, 0, d->stream>>>(P, d->transfer, d->wq, partial);
  spectra_final_kernel<<<(T.l_size + 127) / 128, 128, 0, d->stream>>>(P, partial, d->cl);
  cudaEventRecord(d->ev[1], d->stream);
  c->launches += 2;
  CLPP_CUDA(cudaGetLastError(), err);
  CLPP_CUDA(cudaMemcpyAsync(cl_out, d->cl, (size_t)T.l_size * I.ct_size * sizeof(double), cudaMemcpyDeviceToHost,
                            d->stream), err);
  CLPP_CUDA(cudaStreamSynchronize(d->stream), err);
  { float ms = 0; cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]); d->t_spectra_ms = ms; }
  // table on the l grid + second derivatives along l for spectra_cl_at_l (spectra_module.cpp:905-927); only
  // meaningful for the full q range (partial sums of a multi-GPU partition are splined after the all-reduce)
  if (q_begin == 0 && q_end == T.q_size) {
    HostTable& t = c->clt;
    t.n_lines = T.l_size;
    t.n_cols = I.ct_size;
    t.x.resize(T.l_size);
    for (int i = 0; i < T.l_size; i++) t.x[i] = (double)c->l[i];
    t.y.assign(cl_out, cl_out + (size_t)T.l_size * I.ct_size);
    t.ddy.resize(t.y.size());
    clpp_spline_table_lines(t.x.data(), t.n_lines, t.y.data(), t.n_cols, t.ddy.data());
    c->cl_l_max = c->td.l_scalar_max;
    c->sinfo = I;
    c->has_cl_lens = false;
    c->has_cl = true;
  }
  return CLPP_SUCCESS;
}

// spectra_cl_at_l, case (a): one mode, one initial condition (spectra_module.cpp:236-264)
int clpp_host_cl_at_l(const clpp_ctx* c, double l, double* cl_tot, char* err) {
  CLPP_CHECK(c && c->has_cl, err, "no C_l table: run clpp_spectra_compute first");
  const HostTable& t = c->clt;
  if ((int)l <= (int)t.x[t.n_lines - 1]) {
    int last = 0;
    if (clpp_interp_spline(t, l, &last, cl_tot, t.n_cols, err)) return CLPP_FAILURE;
    if ((int)l > c->cl_l_max)
      for (int ct = 0; ct < t.n_cols; ct++) cl_tot[ct] = 0.;
  } else {
    for (int ct = 0; ct < t.n_cols; ct++) cl_tot[ct] = 0.;
  }
  return CLPP_SUCCESS;
}


// =============================================================================================
// Linear matter power spectrum P(k) = 2 pi^2 / k^3 P_R(k) delta_m(k, tau)^2 on the perturbation k grid
// (NonlinearModule::nonlinear_pk_linear, nonlinear_module.cpp:1886-2024, adiabatic mode): SURVEY 8f row 1.
// Reads delta_m (or delta_cb) straight from the device-resident source table [tp][k][tau].
// =============================================================================================
__global__ void pk_linear_kernel(int nk, int nt, int index_tau, const double* __restrict__ k, const double* __restrict__ src_tp,
                                 const double* __restrict__ primordial, double* __restrict__ pk) {
  const int ik = blockIdx.x * blockDim.x + threadIdx.x;
  if (ik >= nk) return;
  const double s = src_tp[(size_t)ik * nt + index_tau];
  const double kk = k[ik];
  pk[ik] = 2. * CLPP_PI * CLPP_PI / (kk * kk * kk) * s * s * primordial[ik];
}

int clpp_dev_pk_linear(clpp_ctx* c, const double* primordial_pk, int index_tau, int cb, double* pk_out, char* err) {
  clpp_ctx::Dev* d = c->dev;
  const clpp_perturb_info& I = c->pinfo;
  const int nk = I.k_size, nt = I.tau_size;
  const int tp = cb ? I.index_tp_delta_cb : I.index_tp_delta_m;
  CLPP_CHECK(tp >= 0, err, "P(k) is set neither to total matter nor to cold dark matter + baryons: no %s source", cb ? "delta_cb" : "delta_m");
  if (index_tau < 0) index_tau = nt - 1;
  CLPP_CHECK(index_tau < nt, err, "index_tau=%d out of range [0,%d)", index_tau, nt);
  if (clpp_dev_reserve(d, &d->pk, (size_t)2 * nk, err)) return CLPP_FAILURE;
  CLPP_CUDA(cudaMemcpyAsync(d->pk, primordial_pk, nk * sizeof(double), cudaMemcpyHostToDevice, d->stream), err);
  pk_linear_kernel<<<(nk + 127) / 128, 128, 0, d->stream>>>(nk, nt, index_tau, d->k, d->sources + (size_t)tp * nk * nt, d->pk,
                                                           d->pk + nk);
  c->launches++;
  CLPP_CUDA(cudaGetLastError(), err);
  CLPP_CUDA(cudaMemcpyAsync(pk_out, d->pk + nk, nk * sizeof(double), cudaMemcpyDeviceToHost, d->stream), err);
  CLPP_CUDA(cudaStreamSynchronize(d->stream), err);
  return CLPP_SUCCESS;
}


// =============================================================================================
// PerturbationsModule::perturb_sources_at_tau (perturbations_module.cpp:79-131), the branch every default run takes
// (z_max_pk = 0: ln_tau_size_ <= 1): linear interpolation in tau of one source type at all k (array_interpolate_two_bis,
// tools/arrays.c:2380-2436), read from the device-resident table [tp][k][tau].
// =============================================================================================
__global__ void sources_at_tau_kernel(int nk, int nt, int inf, int sup, double weight, const double* __restrict__ src_tp,
                                      double* __restrict__ out) {
  const int ik = blockIdx.x * blockDim.x + threadIdx.x;
  if (ik >= nk) return;
  const double* row = src_tp + (size_t)ik * nt;
  // no FMA contraction: same rounding as the reference's CPU expression
  out[ik] = __dadd_rn(__dmul_rn(row[inf], 1. - weight), __dmul_rn(weight, row[sup]));
}

int clpp_dev_sources_at_tau(clpp_ctx* c, int index_tp, double tau, double* psource, char* err) {
  clpp_ctx::Dev* d = c->dev;
  const clpp_perturb_info& I = c->pinfo;
  const int nk = I.k_size, nt = I.tau_size;
  CLPP_CHECK(index_tp >= 0 && index_tp < I.tp_size, err, "source type %d outside [0,%d)", index_tp, I.tp_size);
  const std::vector<double>& x = c->tau;
  int inf = 0, sup = nt - 1;
  CLPP_CHECK(tau >= x[inf], err, "x=%e < x_min=%e", tau, x[inf]);
  CLPP_CHECK(tau <= x[sup], err, "x=%e > x_max=%e", tau, x[sup]);
  while (sup - inf > 1) {
    const int mid = (int)(0.5 * (inf + sup));
    if (tau < x[mid]) sup = mid;
    else inf = mid;
  }
  const double weight = (tau - x[inf]) / (x[sup] - x[inf]);
  if (clpp_dev_reserve(d, &d->pk, (size_t)2 * nk, err)) return CLPP_FAILURE;
  sources_at_tau_kernel<<<(nk + 127) / 128, 128, 0, d->stream>>>(nk, nt, inf, sup, weight, d->sources + (size_t)index_tp * nk * nt, d->pk);
  c->launches++;
  CLPP_CUDA(cudaGetLastError(), err);
  CLPP_CUDA(cudaMemcpyAsync(psource, d->pk, nk * sizeof(double), cudaMemcpyDeviceToHost, d->stream), err);
  CLPP_CUDA(cudaStreamSynchronize(d->stream), err);
  return CLPP_SUCCESS;
}
