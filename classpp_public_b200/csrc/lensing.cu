// Lensed C_l (TT, TE, EE, BB) from the unlensed spectra and C_l^{phi phi}: SURVEY 8f row 2.
// Replaces LensingModule::lensing_init (lensing_module.cpp:149-860): full-sky correlation-function method of
// Challinor & Lewis 2005. Fast mode (accurate_lensing = 0): Riemann sum over theta in (0, pi/16] of the lensed MINUS
// unlensed correlation functions, unlensed C_l added back (:249-259, :683-748, :1152-1260); accurate mode: Gauss-Legendre
// nodes on [-1,1] (:237-248).
//
// Device layout: one thread per angle mu. Each thread runs the twelve reduced Wigner d^l_{mn}(mu) three-term recurrences
// (:1261-1935) side by side in registers for l = 2..l_max and accumulates the four correlation functions on the fly; the
// reference's d[mu][l] tables (12 x num_mu x l_max doubles) never exist. Only the four d's needed for the back-transform
// are stored, and only at the ~60 multipoles of the output grid. The l-dependent recurrence coefficients are the same for
// every angle: built once per l_max by a small kernel and read as warp-wide broadcasts.
#include <cuda_pipeline.h>

#include <cmath>
#include <vector>

#include "device.h"

int clpp_host_cl_at_l(const clpp_ctx* c, double l, double* cl_tot, char* err);

namespace {

enum { F_00 = 0, F_11, F_1M1, F_2M2, F_22, F_20, F_31, F_3M1, F_3M3, F_40, F_4M2, F_4M4, NFAM };
__constant__ int c_fam_m[NFAM] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4};
__constant__ int c_fam_n[NFAM] = {0, 1, -1, -2, 2, 0, 1, -1, -3, 0, -2, -4};

// One 512-byte row of l-only quantities per multipole, shared by every angle and every cosmology with this l_max:
//   row[f*4 + {0,1,2}] = a, b, c of  N_{l+1} = a (mu - b) N_l - c N_{l-1},  N_l = sqrt((2l+1)/2) d^l_mn  (:1261-1935);
//                        zero below the first multipole of the family, so the recurrence runs unconditionally on N = 0
//   row[48..55]        = l(l+1)/4, (2l+1)/4pi, 8/(l(l+1)), sqrt1/4, sqrt4/4, -sqrt2/2, -sqrt3/2, 2/sqrt5   (:622-681):
//                        the angle loop has no square root or division left
//   row[56]            = sqrt(2/(2l+1)),  row[57] = (2l+1) l (l+1)
constexpr int LROW = 64;  // doubles per row
constexpr int LCH = 32;   // rows per shared-memory stage

__global__ void lens_table_kernel(int lmax, double* __restrict__ tab) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  const double ll = (double)l;
  double* row = tab + (size_t)l * LROW;
  for (int f = 0; f < NFAM; f++) {
    const int m = c_fam_m[f], n = c_fam_n[f];
    const int l0 = max(max(abs(m), abs(n)), 1);
    double a = 0., b = 0., c = 0.;
    if (l >= l0) {
      const double m2 = (double)(m * m), n2 = (double)(n * n), l1 = ll + 1.;
      const double den = sqrt((l1 * l1 - m2) * (l1 * l1 - n2));
      a = sqrt((2. * ll + 3.) * (2. * ll + 1.)) * l1 / den;
      b = (double)(m * n) / (ll * l1);
      c = sqrt((2. * ll + 3.) / (2. * ll - 1.)) * sqrt((ll * ll - m2) * (ll * ll - n2)) / den * l1 / ll;
    }
    row[f * 4 + 0] = a; row[f * 4 + 1] = b; row[f * 4 + 2] = c; row[f * 4 + 3] = 0.;
  }
  row[48] = ll * (ll + 1.) / 4.;
  row[49] = (2. * ll + 1.) / (4. * CLPP_PI);
  row[50] = l > 0 ? 8. / (ll * (ll + 1.)) : 0.;
  row[51] = l >= 1 ? 0.25 * sqrt((ll + 2.) * (ll + 1.) * ll * (ll - 1.)) : 0.;
  row[52] = l >= 3 ? 0.25 * sqrt((ll + 4.) * (ll + 3.) * (ll - 2.) * (ll - 3.)) : 0.;
  row[53] = l >= 1 ? -0.5 * sqrt((ll + 2.) * (ll - 1.)) : 0.;
  row[54] = l >= 2 ? -0.5 * sqrt((ll + 3.) * (ll - 2.)) : 0.;
  row[55] = l > 0 ? 2. / sqrt(ll * (ll + 1.)) : 0.;
  row[56] = sqrt(2. / (2. * ll + 1.));
  row[57] = (2. * ll + 1.) * ll * (ll + 1.);
  for (int j = 58; j < LROW; j++) row[j] = 0.;
}

struct LensParams {
  int lmax, num_mu, n_int;  // n_int: number of integration nodes (num_mu - 1 in both modes; the last mu is 1)
  int l_size;
  int has_te, has_pol, subtract_unlensed;
  const double *mu, *w8, *tab;
  const double* cl;  // [5][lmax+1]: tt, te, ee, bb, pp
  const int* lgrid;
  double *cgl, *cgl2;  // [num_mu]
  double *ksi;         // [4][n_int]: ksi, ksiX, ksip, ksim
  double *dgrid;       // [4][l_size][n_int]: d00, d20, d22, d2m2 at the output multipoles
  double *out;         // [l_size][4]: tt, te, ee, bb integrals
};

// The angle kernels are one warp per CTA, every lane walking l = 1..l_max in lock step, so the l-only rows and the C_l are
// staged through shared memory: asynchronous 16-byte copies of the next LCH rows (double buffered) while the current
// ones are consumed as warp-wide broadcasts. Without the staging every iteration waits an L2 round trip.
struct LensStage {
  double tab[2][LCH * LROW];
  double cl[2][5][LCH];
};

__device__ __forceinline__ void lens_prefetch(LensStage& S, int buf, const LensParams& P, int l_begin) {
  const int lane = threadIdx.x;
  const int nrow = min(LCH, P.lmax + 1 - l_begin);
  const double* src = P.tab + (size_t)l_begin * LROW;
  for (int i = lane; i < nrow * (LROW / 2); i += 32) __pipeline_memcpy_async(&S.tab[buf][2 * i], src + 2 * i, 16);
  if (lane < nrow) {
#pragma unroll
    for (int j = 0; j < 5; j++)
      __pipeline_memcpy_async(&S.cl[buf][j][lane], P.cl + (size_t)j * (P.lmax + 1) + l_begin + lane, 8);
  }
  __pipeline_commit();
}

__device__ __forceinline__ void lens_step(const double* __restrict__ row, int f, double mu, double& cur, double& prev) {
  const double2 ab = *reinterpret_cast<const double2*>(row + f * 4);
  const double c = row[f * 4 + 2];
  const double nxt = ab.x * (mu - ab.y) * cur - c * prev;
  prev = cur;
  cur = nxt;
}

// Cgl(mu), Cgl2(mu) (lensing_module.cpp:560-575): sums over l of (2l+1) l (l+1) C_l^pp d^l_{11}, d^l_{1-1} / 4 pi
__global__ void __launch_bounds__(32) lens_cgl_kernel(LensParams P) {
  __shared__ __align__(16) LensStage S;
  const int i = min((int)(blockIdx.x * 32 + threadIdx.x), P.num_mu - 1);
  const double mu = P.mu[i];
  double c11 = (1. + mu) / 2. * sqrt(3. / 2.), p11 = 0.;
  double c1m1 = (1. - mu) / 2. * sqrt(3. / 2.), p1m1 = 0.;
  double s1 = 0., s2 = 0.;
  const int nch = P.lmax / LCH + 1;
  lens_prefetch(S, 0, P, 0);
  for (int c = 0; c < nch; c++) {
    if (c + 1 < nch) lens_prefetch(S, (c + 1) & 1, P, (c + 1) * LCH);
    else __pipeline_commit();
    __pipeline_wait_prior(1);
    __syncwarp();
    const int l_end = min((c + 1) * LCH - 1, P.lmax);
    for (int l = max(c * LCH, 1); l <= l_end; l++) {
      const double* row = S.tab[c & 1] + (l - c * LCH) * LROW;
      if (l >= 2) {
        const double w = row[57] * S.cl[c & 1][4][l - c * LCH] * row[56];
        s1 += w * c11;
        s2 += w * c1m1;
      }
      if (l < P.lmax) {
        lens_step(row, F_11, mu, c11, p11);
        lens_step(row, F_1M1, mu, c1m1, p1m1);
      }
    }
    __syncwarp();
  }
  if (blockIdx.x * 32 + threadIdx.x < P.num_mu) {
    P.cgl[i] = s1 / (4. * CLPP_PI);
    P.cgl2[i] = s2 / (4. * CLPP_PI);
  }
}

// lensed (minus unlensed) correlation functions ksi, ksiX, ksi+, ksi- at one angle per thread (:628-738)
__global__ void __launch_bounds__(32) lens_ksi_kernel(LensParams P) {
  __shared__ __align__(16) LensStage S;
  const int lmax = P.lmax;
  const bool live = (int)(blockIdx.x * 32 + threadIdx.x) < P.n_int;
  const int i = min((int)(blockIdx.x * 32 + threadIdx.x), P.n_int - 1);
  const double mu = P.mu[i];
  const double sigma2 = P.cgl[P.num_mu - 1] - P.cgl[i];
  const double cgl2 = P.cgl2[i];
  const double op = 1. + mu, om = 1. - mu;

  double cur[NFAM], prev[NFAM];
#pragma unroll
  for (int f = 0; f < NFAM; f++) { cur[f] = 0.; prev[f] = 0.; }
  // first multipoles: closed forms of d^l_mn at l = max(|m|,|n|), times sqrt((2l+1)/2); l = 0 and 1 here
  prev[F_00] = 1. / sqrt(2.);
  cur[F_00] = mu * sqrt(3. / 2.);
  cur[F_11] = op / 2. * sqrt(3. / 2.);
  cur[F_1M1] = om / 2. * sqrt(3. / 2.);

  double ksi = 0., ksiX = 0., ksip = 0., ksim = 0.;
  int ig = 0;
  int lg = P.l_size > 0 ? P.lgrid[0] : -1;
  const int nch = lmax / LCH + 1;
  lens_prefetch(S, 0, P, 0);
  for (int c = 0; c < nch; c++) {
    if (c + 1 < nch) lens_prefetch(S, (c + 1) & 1, P, (c + 1) * LCH);
    else __pipeline_commit();
    __pipeline_wait_prior(1);
    __syncwarp();
    const int l_end = min((c + 1) * LCH - 1, lmax);
#pragma unroll 1
    for (int l = max(c * LCH, 1); l <= l_end; l++) {
      const double* row = S.tab[c & 1] + (l - c * LCH) * LROW;
      if (l == 2) {
        cur[F_2M2] = om * om / 4. * sqrt(5. / 2.);
        cur[F_22] = op * op / 4. * sqrt(5. / 2.);
        cur[F_20] = sqrt(15.) / 4. * (1. - mu * mu);
      } else if (l == 3) {
        cur[F_31] = sqrt(105. / 2.) * op * op * om / 8.;
        cur[F_3M1] = sqrt(105. / 2.) * op * om * om / 8.;
        cur[F_3M3] = sqrt(7. / 2.) * om * om * om / 8.;
      } else if (l == 4) {
        cur[F_40] = sqrt(315.) * op * op * om * om / 16.;
        cur[F_4M2] = sqrt(126.) * op * om * om * om / 16.;
        cur[F_4M4] = sqrt(9. / 2.) * om * om * om * om / 16.;
      }
      if (l >= 2) {
        const double nl = row[56];
        const double d00 = cur[F_00] * nl, d11 = cur[F_11] * nl, d1m1 = cur[F_1M1] * nl, d2m2 = cur[F_2M2] * nl;
        const double d22 = cur[F_22] * nl, d20 = cur[F_20] * nl, d31 = cur[F_31] * nl, d3m1 = cur[F_3M1] * nl;
        const double d3m3 = cur[F_3M3] * nl, d40 = cur[F_40] * nl, d4m2 = cur[F_4M2] * nl, d4m4 = cur[F_4M4] * nl;
        if (l == lg) {
          if (live) {
            const size_t st = (size_t)P.l_size * P.n_int;
            const size_t o = (size_t)ig * P.n_int + i;
            P.dgrid[o] = d00; P.dgrid[st + o] = d20; P.dgrid[2 * st + o] = d22; P.dgrid[3 * st + o] = d2m2;
          }
          ig++;
          lg = ig < P.l_size ? P.lgrid[ig] : -1;
        }
        const double cl_tt = S.cl[c & 1][0][l - c * LCH], cl_te = S.cl[c & 1][1][l - c * LCH];
        const double cl_ee = S.cl[c & 1][2][l - c * LCH], cl_bb = S.cl[c & 1][3][l - c * LCH];
        const double2 lfa = *reinterpret_cast<const double2*>(row + 48), lfb = *reinterpret_cast<const double2*>(row + 50);
        const double2 lfc = *reinterpret_cast<const double2*>(row + 52), lfd = *reinterpret_cast<const double2*>(row + 54);
        const double fac = lfa.x, fac1 = lfa.y;
        const double X_000 = exp(-fac * sigma2);
        const double X_p000 = -fac * X_000;
        const double X_220 = lfb.y * X_000;
        {
          double lens = X_000 * X_000 * d00 + X_p000 * X_p000 * d1m1 * cgl2 * lfb.x +
                        (X_p000 * X_p000 * d00 + X_220 * X_220 * d2m2) * cgl2 * cgl2;
          if (P.subtract_unlensed) lens -= d00;
          ksi += fac1 * cl_tt * lens;
        }
        if (P.has_te | P.has_pol) {
          const double X_022 = X_000 * (1. + sigma2 * (1. + 0.5 * sigma2));
          const double X_p022 = -(fac - 1.) * X_022;
          const double X_242 = lfc.x * X_000;
          double X_121 = 0., X_132 = 0.;
          if (P.has_pol) {
            X_121 = lfc.y * X_000 * (1. + 2. / 3. * sigma2);
            X_132 = lfd.x * X_000 * (1. + 5. / 3. * sigma2);
          }
          if (P.has_te) {
            double lens = X_022 * X_000 * d20 + cgl2 * X_p000 * lfd.y * (X_121 * d11 + X_132 * d3m1) +
                          0.5 * cgl2 * cgl2 * ((2. * X_p022 * X_p000 + X_220 * X_220) * d20 + X_220 * X_242 * d4m2);
            if (P.subtract_unlensed) lens -= d20;
            ksiX += fac1 * cl_te * lens;
          }
          if (P.has_pol) {
            double lensp = X_022 * X_022 * d22 + 2. * cgl2 * X_132 * X_121 * d31 +
                           cgl2 * cgl2 * (X_p022 * X_p022 * d22 + X_242 * X_220 * d40);
            double lensm = X_022 * X_022 * d2m2 + cgl2 * (X_121 * X_121 * d1m1 + X_132 * X_132 * d3m3) +
                           0.5 * cgl2 * cgl2 * (2. * X_p022 * X_p022 * d2m2 + X_220 * X_220 * d00 + X_242 * X_242 * d4m4);
            if (P.subtract_unlensed) { lensp -= d22; lensm -= d2m2; }
            ksip += fac1 * (cl_ee + cl_bb) * lensp;
            ksim += fac1 * (cl_ee - cl_bb) * lensm;
          }
        }
      }
      if (l < lmax) {
#pragma unroll
        for (int f = 0; f < NFAM; f++) lens_step(row, f, mu, cur[f], prev[f]);
      }
    }
    __syncwarp();
  }
  if (live) {
    P.ksi[i] = ksi;
    P.ksi[P.n_int + i] = ksiX;
    P.ksi[2 * (size_t)P.n_int + i] = ksip;
    P.ksi[3 * (size_t)P.n_int + i] = ksim;
  }
}

// back-transform at the output multipoles (:1152-1230): one CTA per multipole, reduction over the angles
__global__ void lens_cl_kernel(LensParams P) {
  const int il = blockIdx.x;
  const size_t st = (size_t)P.l_size * P.n_int;
  const double* d = P.dgrid + (size_t)il * P.n_int;
  double s[4] = {0., 0., 0., 0.};
  for (int i = threadIdx.x; i < P.n_int; i += blockDim.x) {
    const double w = P.w8[i];
    s[0] += P.ksi[i] * d[i] * w;
    s[1] += P.ksi[P.n_int + i] * d[st + i] * w;
    s[2] += P.ksi[2 * (size_t)P.n_int + i] * d[2 * st + i] * w;
    s[3] += P.ksi[3 * (size_t)P.n_int + i] * d[3 * st + i] * w;
  }
  __shared__ double red[4][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
    if (lane == 0) red[j][wid] = s[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    double t[4] = {0., 0., 0., 0.};
    for (int j = 0; j < 4; j++)
      for (int w = 0; w < nw; w++) t[j] += red[j][w];
    double* o = P.out + (size_t)il * 4;
    o[0] = t[0] * 2. * CLPP_PI;
    o[1] = t[1] * 2. * CLPP_PI;
    o[2] = (t[2] + t[3]) * CLPP_PI;
    o[3] = (t[2] - t[3]) * CLPP_PI;
  }
}

}  // namespace

int clpp_dev_lensing(clpp_ctx* c, const clpp_lensing_desc* ld, clpp_lensing_info* info, double* l_out, double* cl_lens_out,
                     char* err) {
  clpp_ctx::Dev* d = c->dev;
  CLPP_CHECK(c->has_cl, err, "no C_l table: run clpp_spectra_compute first");
  const HostTable& t = c->clt;
  const clpp_spectra_info& S = c->sinfo;
  CLPP_CHECK(S.index_ct_tt >= 0 && S.index_ct_pp >= 0, err,
             "lensed C_l need the temperature and lensing-potential spectra (output must include tCl and lCl)");
  const int lt_size = S.ct_size;
  const int lmax = (int)t.x[t.n_lines - 1];  // l_unlensed_max_ = l_max_tot_ (:992)
  const int l_lensed_max = lmax - ld->delta_l_max;
  CLPP_CHECK(l_lensed_max >= 2, err, "delta_l_max=%d leaves no lensed multipole below l_max=%d", ld->delta_l_max, lmax);
  // output multipoles: the spectra grid up to l_lensed_max, plus the following points for the interpolation (:996-1004)
  int il = 0;
  while (il < t.n_lines && t.x[il] <= l_lensed_max) il++;
  if (il < t.n_lines) il++;
  const int l_size = il + 1 <= t.n_lines ? il + 1 : t.n_lines;
  const bool has_te = S.index_ct_te >= 0, has_pol = S.index_ct_ee >= 0 || S.index_ct_bb >= 0;

  int num_mu;
  if (ld->accurate_lensing) {
    num_mu = lmax + ld->num_mu_minus_lmax;
    num_mu += num_mu % 2;
  } else {
    num_mu = (lmax * 2) / 16;
  }
  CLPP_CHECK(num_mu >= 2, err, "l_max=%d too small for the lensing quadrature", lmax);
  const int n_int = num_mu - 1;

  // host staging: [mu (num_mu) | w8 (n_int) | cl (5 x (lmax+1))], one upload
  std::vector<double>& h = c->lens_stage;
  const size_t n_stage = (size_t)num_mu + n_int + 5 * (size_t)(lmax + 1);
  h.assign(n_stage, 0.);
  double *mu = h.data(), *w8 = mu + num_mu, *cl5 = w8 + n_int;
  mu[num_mu - 1] = 1.;
  if (ld->accurate_lensing) {
    if (c->gl_n != n_int || c->gl_tol != ld->tol_gauss_legendre) {
      c->gl_nodes.resize(2 * (size_t)n_int);
      if (clpp_gauss_legendre(c->gl_nodes.data(), c->gl_nodes.data() + n_int, n_int, ld->tol_gauss_legendre, err)) return CLPP_FAILURE;
      c->gl_n = n_int;
      c->gl_tol = ld->tol_gauss_legendre;
    }
    for (int i = 0; i < n_int; i++) { mu[i] = c->gl_nodes[i]; w8[i] = c->gl_nodes[n_int + i]; }
  } else {
    const double delta_theta = CLPP_PI / 16. / (double)(num_mu - 1);
    for (int i = 0; i < n_int; i++) {
      const double theta = (i + 1) * delta_theta;
      mu[i] = cos(theta);
      w8[i] = sin(theta) * delta_theta;
    }
  }
  std::vector<double> row(lt_size);
  for (int l = 2; l <= lmax; l++) {
    if (clpp_host_cl_at_l(c, (double)l, row.data(), err)) return CLPP_FAILURE;
    cl5[l] = row[S.index_ct_tt];
    if (has_te) cl5[(size_t)(lmax + 1) + l] = row[S.index_ct_te];
    if (S.index_ct_ee >= 0) cl5[2 * (size_t)(lmax + 1) + l] = row[S.index_ct_ee];
    if (S.index_ct_bb >= 0) cl5[3 * (size_t)(lmax + 1) + l] = row[S.index_ct_bb];
    cl5[4 * (size_t)(lmax + 1) + l] = row[S.index_ct_pp];
  }
  std::vector<int> lgrid(l_size);
  for (int i = 0; i < l_size; i++) lgrid[i] = (int)t.x[i];

  if (clpp_dev_reserve(d, &d->lens_stage, n_stage, err)) return CLPP_FAILURE;
  if (clpp_dev_reserve(d, &d->lens_lgrid, (size_t)l_size, err)) return CLPP_FAILURE;
  if (clpp_dev_reserve(d, &d->lens_work, (size_t)2 * num_mu + 4 * (size_t)n_int + 4 * (size_t)l_size * n_int + 4 * (size_t)l_size, err))
    return CLPP_FAILURE;
  cudaStream_t s = d->stream;
  if (d->lens_lmax != lmax) {
    if (clpp_dev_reserve(d, &d->lens_coef, (size_t)LROW * (lmax + 1), err)) return CLPP_FAILURE;
    lens_table_kernel<<<(lmax + 128) / 128, 128, 0, s>>>(lmax, d->lens_coef);
    c->launches++;
    d->lens_lmax = lmax;
  }
  CLPP_CUDA(cudaMemcpyAsync(d->lens_stage, h.data(), n_stage * sizeof(double), cudaMemcpyHostToDevice, s), err);
  CLPP_CUDA(cudaMemcpyAsync(d->lens_lgrid, lgrid.data(), l_size * sizeof(int), cudaMemcpyHostToDevice, s), err);

  LensParams P;
  P.lmax = lmax; P.num_mu = num_mu; P.n_int = n_int; P.l_size = l_size;
  P.has_te = has_te; P.has_pol = has_pol; P.subtract_unlensed = ld->accurate_lensing ? 0 : 1;
  P.mu = d->lens_stage; P.w8 = P.mu + num_mu; P.cl = P.w8 + n_int;
  P.tab = d->lens_coef;
  P.lgrid = d->lens_lgrid;
  P.cgl = d->lens_work; P.cgl2 = P.cgl + num_mu; P.ksi = P.cgl2 + num_mu; P.dgrid = P.ksi + 4 * (size_t)n_int;
  P.out = P.dgrid + 4 * (size_t)l_size * n_int;
  cudaEventRecord(d->ev[0], s);
  lens_cgl_kernel<<<(num_mu + 31) / 32, 32, 0, s>>>(P);
  lens_ksi_kernel<<<(n_int + 31) / 32, 32, 0, s>>>(P);
  lens_cl_kernel<<<l_size, 128, 0, s>>>(P);
  cudaEventRecord(d->ev[1], s);
  c->launches += 3;
  CLPP_CUDA(cudaGetLastError(), err);
  std::vector<double> integ((size_t)l_size * 4);
  CLPP_CUDA(cudaMemcpyAsync(integ.data(), P.out, integ.size() * sizeof(double), cudaMemcpyDeviceToHost, s), err);
  CLPP_CUDA(cudaStreamSynchronize(s), err);
  { float ms = 0; cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]); d->t_lensing_ms = ms; }

  // table of lensed spectra on the output grid: unlensed values for the types lensing leaves alone (:1030-1034),
  // integrals (+ unlensed C_l in fast mode, :1166-1260) for tt, te, ee, bb; then the spline along l (:783-791)
  HostTable& L = c->cl_lens;
  L.n_lines = l_size;
  L.n_cols = lt_size;
  L.x.assign(t.x.begin(), t.x.begin() + l_size);
  L.y.resize((size_t)l_size * lt_size);
  L.ddy.resize(L.y.size());
  for (int i = 0; i < l_size; i++) {
    double* o = L.y.data() + (size_t)i * lt_size;
    if (clpp_host_cl_at_l(c, L.x[i], o, err)) return CLPP_FAILURE;
    const int l = lgrid[i];
    const double add = ld->accurate_lensing ? 0. : 1.;
    o[S.index_ct_tt] = integ[(size_t)i * 4 + 0] + add * cl5[l];
    if (has_te) o[S.index_ct_te] = integ[(size_t)i * 4 + 1] + add * cl5[(size_t)(lmax + 1) + l];
    if (has_pol) {
      if (S.index_ct_ee >= 0) o[S.index_ct_ee] = integ[(size_t)i * 4 + 2] + add * cl5[2 * (size_t)(lmax + 1) + l];
      if (S.index_ct_bb >= 0) o[S.index_ct_bb] = integ[(size_t)i * 4 + 3] + add * cl5[3 * (size_t)(lmax + 1) + l];
    }
  }
  clpp_spline_table_lines(L.x.data(), L.n_lines, L.y.data(), L.n_cols, L.ddy.data());
  c->l_lensed_max = l_lensed_max;
  c->has_cl_lens = true;

  if (info) {
    info->lt_size = lt_size; info->l_size = l_size; info->l_unlensed_max = lmax; info->l_lensed_max = l_lensed_max;
    info->index_lt_tt = S.index_ct_tt; info->index_lt_ee = S.index_ct_ee; info->index_lt_te = S.index_ct_te;
    info->index_lt_bb = S.index_ct_bb; info->index_lt_pp = S.index_ct_pp; info->index_lt_tp = S.index_ct_tp;
    info->index_lt_ep = S.index_ct_ep;
  }
  if (l_out) for (int i = 0; i < l_size; i++) l_out[i] = L.x[i];
  if (cl_lens_out) for (size_t i = 0; i < L.y.size(); i++) cl_lens_out[i] = L.y[i];
  return CLPP_SUCCESS;
}

// LensingModule::lensing_cl_at_l (lensing_module.cpp:111-140)
int clpp_host_lensing_cl_at_l(const clpp_ctx* c, int l, double* cl_lensed, char* err) {
  CLPP_CHECK(c && c->has_cl_lens, err, "no lensed C_l table: run clpp_lensing_compute first");
  CLPP_CHECK(l <= c->l_lensed_max, err,
             "you asked for lensed Cls at l=%d, they were computed only up to l=%d, you should increase l_max_scalars or "
             "decrease the precision parameter delta_l_max", l, c->l_lensed_max);
  int last = 0;
  if (clpp_interp_spline(c->cl_lens, (double)l, &last, cl_lensed, c->cl_lens.n_cols, err)) return CLPP_FAILURE;
  if (l > c->cl_l_max)
    for (int i = 0; i < c->cl_lens.n_cols; i++) cl_lensed[i] = 0.;
  return CLPP_SUCCESS;
}
