// Stage 2: line-of-sight transfer integrals on the device.
//
//   Delta_l^X(q) = int dtau S^X(k(q),tau) R_l^X(k (tau0 - tau))
//
// Reference path: TransferModule::transfer_init (source/transfer_module.cpp:127-345) ->
//   transfer_perturbation_copy_sources_and_nl_corrections (:542-601)
//   transfer_perturbation_source_spline -> array_spline_table_columns2 (tools/arrays.c:967-1092)
//   hyperspherical_HIS_create (tools/hyperspherical.c:11-246)  [flat case: j_l(x), j_l'(x)]
//   transfer_compute_for_each_q (:1488-1715): interpolate_sources (:1767), transfer_sources (:1845),
//   transfer_can_be_neglected (:3187), transfer_late_source_can_be_neglected (:3229),
//   transfer_use_limber (:2661) / transfer_limber (:2912) / transfer_integrate (:2750) with
//   transfer_radial_function (:3274) over Hermite-4 interpolation (tools/hermite4_interpolation_csource.h).
//
// B200 mapping:
//   k_spline_kernel     one thread per (type, tau): tridiagonal sweep along k; the device layout
//                       [tp][k][tau] makes every step a fully coalesced FP64 row access.
//   bessel_table_kernel one thread per x node: backward (CF1 start) or forward recurrence in l,
//                       writes only the l of the list; 9.2k independent chains of <= 3001 steps.
//   los_kernel          one CTA per q: the sources of this q are interpolated in k once and staged
//                       in shared memory; each warp then owns (type, l) cells, lanes stride over tau,
//                       evaluate the cubic Hermite Bessel interpolant from the L2-resident table and
//                       reduce with warp shuffles.  FP64 vector pipe bound (~40 flop / point).
#include <cmath>

#include "device.h"

// =============================================================================================
// (a) cubic spline of the sources along k
// =============================================================================================
__global__ void k_spline_kernel(int ntp, int nk, int nt, const double* __restrict__ kk, const double* __restrict__ S,
                                double* __restrict__ dd, double* __restrict__ u) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ntp * nt) return;
  const int tp = idx / nt, it = idx % nt;
  const size_t base = (size_t)tp * nk * nt + it;
#define AT(a, ik) a[base + (size_t)(ik) * nt]
  const double x0 = kk[0], x1 = kk[1], x2 = kk[2];
  {
    const double slope0 = ((x2 - x0) * (x2 - x0) * (AT(S, 1) - AT(S, 0)) - (x1 - x0) * (x1 - x0) * (AT(S, 2) - AT(S, 0))) /
                          ((x2 - x0) * (x1 - x0) * (x2 - x1));
    AT(dd, 0) = -0.5;
    AT(u, 0) = (3. / (x1 - x0)) * ((AT(S, 1) - AT(S, 0)) / (x1 - x0) - slope0);
  }
  double d_prev = -0.5, u_prev = AT(u, 0);
  double y_m = AT(S, 0), y_c = AT(S, 1);
  for (int i = 1; i < nk - 1; i++) {
    const double xm = kk[i - 1], xc = kk[i], xp = kk[i + 1];
    const double y_p = AT(S, i + 1);
    const double sig = (xc - xm) / (xp - xm);
    const double p = sig * d_prev + 2.0;
    const double d_i = (sig - 1.0) / p;
    double t = (y_p - y_c) / (xp - xc) - (y_c - y_m) / (xc - xm);
    const double u_i = (6.0 * t / (xp - xm) - sig * u_prev) / p;
    AT(dd, i) = d_i;
    AT(u, i) = u_i;
    d_prev = d_i;
    u_prev = u_i;
    y_m = y_c;
    y_c = y_p;
  }
  {
    const int n = nk;
    const double xa = kk[n - 3], xb = kk[n - 2], xc = kk[n - 1];
    const double slopeN = ((xa - xc) * (xa - xc) * (AT(S, n - 2) - AT(S, n - 1)) -
                           (xb - xc) * (xb - xc) * (AT(S, n - 3) - AT(S, n - 1))) /
                          ((xa - xc) * (xb - xc) * (xa - xb));
    const double qn = 0.5;
    const double un = (3. / (xc - xb)) * (slopeN - (AT(S, n - 1) - AT(S, n - 2)) / (xc - xb));
    AT(dd, n - 1) = (un - qn * u_prev) / (qn * d_prev + 1.0);
  }
  double d_next = AT(dd, nk - 1);
  for (int i = nk - 2; i >= 0; i--) {
    const double v = AT(dd, i) * d_next + AT(u, i);
    AT(dd, i) = v;
    d_next = v;
  }
#undef AT
}

// phi+psi source times the nonlinear density correction (transfer_module.cpp:559-590)
__global__ void nl_correction_kernel(size_t n, double* __restrict__ S, const double* __restrict__ corr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) S[i] *= corr[i];
}

// =============================================================================================
// (b) flat Bessel table  Phi_l(x) = j_l(x), Phi_l'(x)  on a uniform x grid
// =============================================================================================
// Continued fraction CF1 for Phi'_l/Phi_l (modified Lentz), K=0, beta=1 (hyperspherical.c:677-716)
__device__ static bool bessel_cf1(int l, double cotK, double* CF, int* isign) {
  const double tiny = 1e-100, reltol = 2.2204460492503131e-16;
  double bj = l * cotK, fj = bj, Cj = bj, Dj = 0.0;
  *isign = 1;
  for (int j = 1; j <= 1000000; j++) {
    const double aj = -1.0;
    bj = (2 * (l + j) + 1) * cotK;
    Dj = bj + aj * Dj;
    if (Dj == 0.0) Dj = tiny;
    Cj = bj + aj / Cj;
    if (Cj == 0.0) Cj = tiny;
    Dj = 1.0 / Dj;
    const double Delj = Cj * Dj;
    fj = fj * Delj;
    if (Dj < 0) *isign *= -1;
    if (fabs(Delj - 1.0) < reltol) {
      *CF = fj;
      return true;
    }
  }
  return false;
}

// One thread per x node. lrec = l_max+1 is the top of the recurrence (the derivative of the last
// multipole needs Phi_{l_max+1}); only the l of llist are written.
__global__ void bessel_table_kernel(int nx, double xmin, double dx, int xfwdidx, int lrec, int nl,
                                    const int* __restrict__ llist, double* __restrict__ xout,
                                    double* __restrict__ phi, double* __restrict__ dphi, int* __restrict__ scale_count) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nx) return;
  const double x = xmin + j * dx;
  xout[j] = x;
  const double cotK = 1.0 / x;
  if (j < xfwdidx) {
    // ---- backward recurrence from l = lrec, started from CF1, rescaled by 1e-200 on overflow
    double phipr1 = 0.;
    int isign = 1;
    bessel_cf1(lrec, cotK, &phipr1, &isign);
    const double phi1 = isign;
    phipr1 *= phi1;
    double ph = phi1;                                    // Phi_l (unnormalised)
    double ph_plus_times_sqrtK = lrec * cotK * phi1 - phipr1;  // Phi_{l+1} (sqrtK = 1 in flat space)
    int nscale = 0;
    int li = nl - 1;  // next list entry to be stored (descending)
    while (li >= 0 && llist[li] > lrec) li--;
    if (li >= 0 && llist[li] == lrec) {  // cannot happen for lrec = l_max+1, kept for generality
      phi[(size_t)li * nx + j] = ph;
      dphi[(size_t)li * nx + j] = lrec * cotK * ph - ph_plus_times_sqrtK;
      scale_count[(size_t)li * nx + j] = nscale;
      li--;
    }
    const int l_align = lrec - lrec % 8;
    int l = lrec;
    auto step = [&]() {
      const double ph_minus = ((2 * l + 1) * cotK * ph - ph_plus_times_sqrtK);
      ph_plus_times_sqrtK = ph;
      ph = ph_minus;
      l--;  // ph is now Phi_l for the decremented l, ph_plus_times_sqrtK = Phi_{l+1}
      if (li >= 0 && llist[li] == l) {
        phi[(size_t)li * nx + j] = ph;
        dphi[(size_t)li * nx + j] = l * cotK * ph - ph_plus_times_sqrtK;
        scale_count[(size_t)li * nx + j] = nscale;
        li--;
      }
    };
    while (l > l_align) step();
    for (int l_ini = l_align; l_ini > 0; l_ini -= 8) {
      for (int s = 0; s < 8; s++) step();
      if (fabs(ph) > 1e200) {
        ph *= 1e-200;
        ph_plus_times_sqrtK *= 1e-200;
        nscale++;
      }
    }
    // normalise with the analytic Phi_0 = sin(x)/x
    const double phi0 = sin(x) / x;
    const double scaling = phi0 / ph;
    for (int i = 0; i < nl; i++) {
      if (llist[i] > lrec) continue;
      const size_t o = (size_t)i * nx + j;
      double a = phi[o], b = dphi[o];
      for (int r = scale_count[o]; r < nscale && (a != 0. || b != 0.); r++) {
        a *= 1e-200;
        b *= 1e-200;
      }
      phi[o] = a * scaling;
      dphi[o] = b * scaling;
    }
  } else {
    // ---- forward recurrence (stable for x > sqrt(l(l+1)))
    double p0 = sin(x) / x;
    double p1 = p0 * (cotK - 1.0 / tan(x));
    int li = 0;
    // l = 0 and 1 are never in the list (l >= 2)
    double pm = p0, pc = p1;  // Phi_{l-1}, Phi_l with l = 1
    for (int l = 2; l <= lrec; l++) {
      const double pn = (2 * l - 1) * cotK * pc - pm;  // Phi_l
      // when pc is Phi_{l-1} and it is a list multipole, its derivative needs pn = Phi_l
      if (li < nl && llist[li] == l - 1) {
        phi[(size_t)li * nx + j] = pc;
        dphi[(size_t)li * nx + j] = (l - 1) * cotK * pc - pn;
        li++;
      }
      pm = pc;
      pc = pn;
    }
  }
}

// =============================================================================================
// (c) line-of-sight integrals
// =============================================================================================
#define LOS_THREADS 256
#define LOS_MAX_TT 5

enum { RADIAL_T0 = 0, RADIAL_T1 = 1, RADIAL_T2 = 2, RADIAL_E = 3 };

struct LosParams {
  int nk, nt, nq, nl, ntt, nx;
  int q_begin, q_end;
  double tau0, tau_rec, tau0_minus_tau_cut, ra_rec;  // ra_rec = (tau0 - tau_rec) * angular_rescaling
  double xmin, dx;
  double neglect_delta_k[LOS_MAX_TT];
  int tp_of_tt[LOS_MAX_TT], radial[LOS_MAX_TT], is_lcmb[LOS_MAX_TT], late_neglect_ok[LOS_MAX_TT];
  int l_size_tt[LOS_MAX_TT];
  double l_late_threshold;  // transfer_neglect_late_source * angular_rescaling
  double l_switch_limber;
  double lcmb_rescale, lcmb_tilt, lcmb_pivot;
  int index_tau_min_lcmb;  // first tau index > tau_rec
};

// last i in [0, start] with pred(i), -1 if none; warp-cooperative ballot scan from the top
template <typename Pred>
__device__ __forceinline__ int warp_find_last(int start, int lane, Pred pred) {
  for (int top = start; top >= 0; top -= 32) {
    const int i = top - lane;
    const bool ok = (i >= 0) && pred(i);
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (m) return top - (__ffs(m) - 1);
  }
  return -1;
}

__global__ void __launch_bounds__(LOS_THREADS)
los_kernel(LosParams P, const double* __restrict__ kgrid, const double* __restrict__ tau, const double* __restrict__ q,
           const double* __restrict__ kq, const int* __restrict__ llist, const double* __restrict__ chi_at_phimin,
           const double* __restrict__ S, const double* __restrict__ Sdd, const double* __restrict__ bphi,
           const double* __restrict__ bdphi, double* __restrict__ transfer, unsigned long long* __restrict__ counters) {
  extern __shared__ double sm[];
  const int nt = P.nt;
  double* tmt = sm;               // tau0 - tau                       [nt]
  double* wtr = tmt + nt;         // trapezoid weights, full range    [nt]
  double* wtr_l = wtr + nt;       // trapezoid weights, lcmb range    [nt]
  double* src = wtr_l + nt;       // transfer sources per tt          [ntt][nt]
  const int iq = P.q_begin + blockIdx.x;
  if (iq >= P.q_end) return;
  const double k = kq[iq], qv = q[iq];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = LOS_THREADS / 32;

  // ---- bracket k(q) in the perturbation k grid (transfer_interpolate_sources, :1792-1799)
  __shared__ int s_ik;
  if (tid == 0) {
    int lo = 0, hi = P.nk - 1;  // largest ik with kgrid[ik] < k, clipped to nk-2 (same result as the linear walk)
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (kgrid[mid] < k) lo = mid; else hi = mid;
    }
    s_ik = lo;
  }
  __syncthreads();
  const int ik = s_ik;
  const double h = kgrid[ik + 1] - kgrid[ik];
  const double b = (k - kgrid[ik]) / h, a = 1. - b;
  const double ca = (a * a * a - a), cb = (b * b * b - b), h26 = h * h / 6.0;

  const int tmin = P.index_tau_min_lcmb;
  const int nt_l = nt - tmin;
  for (int i = tid; i < nt; i += LOS_THREADS) {
    tmt[i] = P.tau0 - tau[i];
  }
  __syncthreads();
  for (int i = tid; i < nt; i += LOS_THREADS) {
    // array_trapezoidal_mweights (tools/arrays.c:2856-2876) on tau0_minus_tau
    double w;
    if (i == 0) w = 0.5 * (tmt[0] - tmt[1]);
    else if (i == nt - 1) w = 0.5 * (tmt[nt - 2] - tmt[nt - 1]);
    else w = 0.5 * (tmt[i - 1] - tmt[i + 1]);
    wtr[i] = w;
    if (i >= tmin) {
      const int j = i - tmin;
      double wl;
      if (nt_l == 1) wl = 1.0;
      else if (j == 0) wl = 0.5 * (tmt[i] - tmt[i + 1]);
      else if (j == nt_l - 1) wl = 0.5 * (tmt[i - 1] - tmt[i]);
      else wl = 0.5 * (tmt[i - 1] - tmt[i + 1]);
      wtr_l[j] = wl;
    }
  }
  for (int tt = 0; tt < P.ntt; tt++) {
    const size_t row = ((size_t)P.tp_of_tt[tt] * P.nk + ik) * nt;
    double* dst = src + (size_t)tt * nt;
    const double lfac = P.is_lcmb[tt] ? P.lcmb_rescale * pow(k / P.lcmb_pivot, P.lcmb_tilt) : 1.0;
    for (int i = tid; i < nt; i += LOS_THREADS) {
      const double v = a * S[row + i] + b * S[row + nt + i] + (ca * Sdd[row + i] + cb * Sdd[row + nt + i]) * h26;
      if (!P.is_lcmb[tt]) {
        dst[i] = v;
      } else if (i >= tmin) {
        // lensing potential: W = (tau_rec - tau)/(tau0 - tau)/(tau0 - tau_rec), zero at tau0 (:1910-1972)
        double resc = 0.;
        if (i != nt - 1) resc = (P.tau_rec - tau[i]) / (P.tau0 - tau[i]) / (P.tau0 - P.tau_rec);
        dst[i - tmin] = v * resc * lfac;
      }
    }
  }
  __syncthreads();

  unsigned long long my_int = 0, my_pts = 0;
  const int ncell = P.ntt * P.nl;
  for (int cell = warp; cell < ncell; cell += nwarp) {
    const int tt = cell / P.nl, il = cell % P.nl;
    const double l = (double)llist[il];
    double* out = transfer + ((size_t)tt * P.nl + il) * P.nq + iq;
    double result = 0.;
    bool done = false;
    if (il >= P.l_size_tt[tt]) done = true;
    // transfer_can_be_neglected: l < (k - delta_k) * ra_rec  (uses q, :1582)
    if (!done && !P.is_lcmb[tt] && (l < (qv - P.neglect_delta_k[tt]) * P.ra_rec)) done = true;
    if (done) {
      if (lane == 0) *out = 0.;
      continue;
    }
    const bool lcmb = P.is_lcmb[tt];
    const double* s = src + (size_t)tt * nt;
    const double* t0 = lcmb ? tmt + tmin : tmt;
    const double* w = lcmb ? wtr_l : wtr;
    const int n = lcmb ? nt_l : nt;

    if (lcmb && l > P.l_switch_limber) {
      // ---- Limber approximation (transfer_limber :2912-3052, SCALAR_TEMPERATURE_0 branch)
      if (lane == 0) {
        const double tl = (l + 0.5) / qv;
        if ((tl > t0[0]) || (tl < t0[n - 1])) {
          result = 0.;
        } else {
          int it = 1;  // first index >= 1 with t0[it] <= tl, capped at n-2 (t0 is decreasing)
          {
            int lo = 1, hi = n - 2;
            if (!(t0[1] > tl)) it = 1;
            else if (t0[hi] > tl) it = hi;
            else {
              while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (t0[mid] > tl) lo = mid; else hi = mid;
              }
              it = hi;
            }
          }
          const double x1 = t0[it - 1], x2 = t0[it], x3 = t0[it + 1];
          const double y1 = s[it - 1] * x1, y2 = s[it] * x2;
          const double y3 = (it < n - 2) ? s[it + 1] * x3 : s[it] * x2;
          // array_interpolate_parabola (tools/arrays.c:2581-2620)
          const double bb = ((y1 - y2) * (x3 - x2) * (x3 + x2) - (y3 - y2) * (x1 - x2) * (x1 + x2)) / (x1 - x2) / (x3 - x2) / (x3 - x1);
          const double aa = (y1 - y2 - bb * (x1 - x2)) / (x1 - x2) / (x1 + x2);
          const double cc = y2 - bb * x2 - aa * x2 * x2;
          const double Sv = aa * tl * tl + bb * tl + cc;
          const double IPhiFlat = sqrt(CLPP_PI / (2. * l)) * (1. - 0.25 / l + 1. / 32. / (l * l));
          result = IPhiFlat * Sv / (l + 0.5);
        }
        *out = result;
      }
      continue;
    }

    // ---- full integral (transfer_integrate :2750-2892)
    const double tmin_bessel = chi_at_phimin[il] / k;
    if (tmin_bessel >= t0[0]) {
      if (lane == 0) *out = 0.;
      continue;
    }
    int imax = warp_find_last(n - 1, lane, [&](int i) { return !(t0[i] < tmin_bessel); });
    const int imax_bessel = imax;
    imax = warp_find_last(imax, lane, [&](int i) { return s[i] != 0.; });
    if (imax < 0) {
      if (lane == 0) *out = 0.;
      continue;
    }
    if (P.late_neglect_ok[tt] && l > P.l_late_threshold) {
      imax = warp_find_last(imax, lane, [&](int i) { return !(t0[i] < P.tau0_minus_tau_cut); });
      if (imax < 0) {
        if (lane == 0) *out = 0.;
        continue;
      }
    }
    const double* ph = bphi + (size_t)il * P.nx;
    const double* dph = bdphi + (size_t)il * P.nx;
    const int radial = P.radial[tt];
    const double lxlp1 = l * (l + 1.0);
    const double xmax = P.xmin + (P.nx - 1) * P.dx;
    const double efac = sqrt(3.0 / 8.0 * (l + 2.0) * (l + 1.0) * l * (l - 1.0));
    double acc = 0., r_last = 0.;
    for (int i = lane; i <= imax; i += 32) {
      const double x = k * t0[i];
      double R = 0.;
      if (!(x < P.xmin) && !(x > xmax)) {
        int ib = (int)((x - P.xmin) / P.dx) + 1;
        ib = max(1, min(P.nx - 1, ib));
        const double xl = P.xmin + (ib - 1) * P.dx, xr = P.xmin + ib * P.dx;
        const double ym = __ldg(ph + ib - 1), yp = __ldg(ph + ib), dym = __ldg(dph + ib - 1), dyp = __ldg(dph + ib);
        const double z = (x - xl) / P.dx, z2 = z * z, z3 = z2 * z;
        const double dxx = P.dx;
        if (radial == RADIAL_T0 || radial == RADIAL_E) {
          const double a0 = dym * dxx, a1 = -2 * dym * dxx - dyp * dxx - 3 * ym + 3 * yp, a2 = dym * dxx + dyp * dxx + 2 * ym - 2 * yp;
          const double Phi = ym + a0 * z + a1 * z2 + a2 * z3;
          if (radial == RADIAL_T0) R = Phi;
          else { const double csc = 1.0 / x; R = efac * csc * csc * Phi; }
        } else {
          // second (and third) derivatives at the nodes from the Bessel ODE
          const double cm = 1.0 / xl, cp = 1.0 / xr;
          const double d2ym = -2 * dym * cm + ym * (lxlp1 * cm * cm - 1.0);
          const double d2yp = -2 * dyp * cp + yp * (lxlp1 * cp * cp - 1.0);
          if (radial == RADIAL_T1) {
            const double b0 = d2ym * dxx, b1 = -2 * d2ym * dxx - d2yp * dxx - 3 * dym + 3 * dyp,
                         b2 = d2ym * dxx + d2yp * dxx + 2 * dym - 2 * dyp;
            R = dym + b0 * z + b1 * z2 + b2 * z3;
          } else {  // RADIAL_T2: (3 Phi'' + Phi)/2
            const double a0 = dym * dxx, a1 = -2 * dym * dxx - dyp * dxx - 3 * ym + 3 * yp, a2 = dym * dxx + dyp * dxx + 2 * ym - 2 * yp;
            const double Phi = ym + a0 * z + a1 * z2 + a2 * z3;
            const double d3ym = -2 * cm * d2ym - 2 * ym * lxlp1 * cm * cm * cm + dym * (-1.0 + (2 + lxlp1) * cm * cm);
            const double d3yp = -2 * cp * d2yp - 2 * yp * lxlp1 * cp * cp * cp + dyp * (-1.0 + (2 + lxlp1) * cp * cp);
            const double c0 = d3ym * dxx, c1 = -2 * d3ym * dxx - d3yp * dxx - 3 * d2ym + 3 * d2yp,
                         c2 = d3ym * dxx + d3yp * dxx + 2 * d2ym - 2 * d2yp;
            const double d2Phi = d2ym + c0 * z + c1 * z2 + c2 * z3;
            R = 0.5 * (3 * d2Phi + Phi);
          }
        }
      }
      acc += s[i] * R * w[i];
      if (i == imax) r_last = R;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc += __shfl_xor_sync(0xffffffffu, acc, o);
      r_last += __shfl_xor_sync(0xffffffffu, r_last, o);
    }
    if (lane == 0) {
      // Bessel truncation: replace the wrong last trapezoid triangle by the right one (:2883-2887)
      if ((imax != n - 1) && (imax == imax_bessel)) acc -= 0.5 * (t0[imax + 1] - tmin_bessel) * r_last * s[imax];
      *out = acc;
      my_int += 1;
      my_pts += (unsigned long long)(imax + 1);
    }
  }
  if (lane == 0 && counters) {
    atomicAdd(&counters[0], my_int);
    atomicAdd(&counters[1], my_pts);
  }
}

// =============================================================================================
// host driver
// =============================================================================================
#define dev_alloc(p, n, err) clpp_dev_reserve(d, p, n, err)

// work buffer of one stage call: grow-only per-context buffer, or (lean_scratch) a stream-ordered pool allocation that is
// returned to the pool when the stage has been enqueued -- the pool recycles it for the next context without any
// device-wide synchronisation (cudaMallocAsync / cudaFreeAsync)
struct ScratchList {
  cudaStream_t st;
  bool lean;
  std::vector<void**> owned;
  template <typename T>
  int get(clpp_ctx::Dev* d, T** p, size_t n, char* err) {
    if (!lean) return clpp_dev_reserve(d, p, n, err);
    if (*p && d->cap.count((void*)p)) { cudaFree(*p); d->cap.erase((void*)p); }  // left over from a non-lean call
    cudaError_t e = cudaMallocAsync((void**)p, (n > 0 ? n : 1) * sizeof(T), st);
    if (e != cudaSuccess) return clpp_fail(err, "cudaMallocAsync of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
    owned.push_back((void**)p);
    return CLPP_SUCCESS;
  }
  void release() {
    for (void** p : owned) { cudaFreeAsync(*p, st); *p = nullptr; }
    owned.clear();
  }
  ~ScratchList() { release(); }  // error paths
};

int clpp_dev_transfer_compute(clpp_ctx* c, const double* nl_corr_density, int q_begin, int q_end, char* err) {
  clpp_ctx::Dev* d = c->dev;
  const clpp_perturb_info& PI = c->pinfo;
  const clpp_transfer_info& TI = c->tinfo;
  const clpp_transfer_desc& td = c->td;
  const int nk = PI.k_size, nt = PI.tau_size, ntp = PI.tp_size;
  const size_t nsrc = (size_t)ntp * nk * nt;
  cudaStream_t st = d->stream;
  ScratchList scratch{st, c->lean_scratch, {}};
  if (c->lean_scratch) {
    static bool pool_ready = false;  // keep freed blocks cached in the pool (no trimming at synchronisation points)
    if (!pool_ready) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      pool_ready = true;
    }
  }
  CLPP_CHECK(TI.tt_size <= LOS_MAX_TT, err, "too many transfer types");
  CLPP_CHECK(nt >= 3 && nk >= 3, err, "source table too small");

  // grids
  if (dev_alloc(&d->q, TI.q_size, err) || dev_alloc(&d->kq, TI.q_size, err) || dev_alloc(&d->l, TI.l_size_max, err))
    return CLPP_FAILURE;
  CLPP_CUDA(cudaMemcpyAsync(d->q, c->q.data(), TI.q_size * sizeof(double), cudaMemcpyHostToDevice, st), err);
  CLPP_CUDA(cudaMemcpyAsync(d->kq, c->kq.data(), TI.q_size * sizeof(double), cudaMemcpyHostToDevice, st), err);
  CLPP_CUDA(cudaMemcpyAsync(d->l, c->l.data(), TI.l_size_max * sizeof(int), cudaMemcpyHostToDevice, st), err);
  if (!d->k || !d->tau) {
    if (dev_alloc(&d->k, nk, err) || dev_alloc(&d->tau, nt, err)) return CLPP_FAILURE;
  }
  CLPP_CUDA(cudaMemcpyAsync(d->k, c->k.data(), nk * sizeof(double), cudaMemcpyHostToDevice, st), err);
  CLPP_CUDA(cudaMemcpyAsync(d->tau, c->tau.data(), nt * sizeof(double), cudaMemcpyHostToDevice, st), err);

  // (a) sources seen by the transfer stage (+ nonlinear correction of phi+psi), spline along k
  if (scratch.get(d, &d->src_tr, nsrc, err) || scratch.get(d, &d->src_ddk, nsrc, err)) return CLPP_FAILURE;
  CLPP_CUDA(cudaMemcpyAsync(d->src_tr, d->sources, nsrc * sizeof(double), cudaMemcpyDeviceToDevice, st), err);
  std::vector<double> corr_t;
  if (!nl_corr_density && c->use_device_nl && c->nl_dev_valid && c->nl_dev_nk == nk && c->nl_dev_nt == nt && d->nl_corr2 &&
      PI.index_tp_phi_plus_psi >= 0) {
    // halofit ran on the device (clpp_nonlinear_halofit): total-matter correction, already in the [k][tau] layout
    const size_t n = (size_t)nk * nt;
    nl_correction_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, d->src_tr + (size_t)PI.index_tp_phi_plus_psi * n,
                                                                     d->nl_corr2);
    c->launches++;
  } else if (nl_corr_density && PI.index_tp_phi_plus_psi >= 0) {
    // reference layout [tau][k] -> device layout [k][tau]
    corr_t.resize((size_t)nk * nt);
    for (int it = 0; it < nt; it++)
      for (int ik = 0; ik < nk; ik++) corr_t[(size_t)ik * nt + it] = nl_corr_density[(size_t)it * nk + ik];
    if (dev_alloc(&d->nl_corr, corr_t.size(), err)) return CLPP_FAILURE;
    CLPP_CUDA(cudaMemcpyAsync(d->nl_corr, corr_t.data(), corr_t.size() * sizeof(double), cudaMemcpyHostToDevice, st), err);
    const size_t n = (size_t)nk * nt;
    nl_correction_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, d->src_tr + (size_t)PI.index_tp_phi_plus_psi * n,
                                                                     d->nl_corr);
    c->launches++;
  }
  {
    if (scratch.get(d, &d->spline_u, nsrc, err)) return CLPP_FAILURE;
    for (int i = 0; i < 6; i++)
      if (!d->ev2[i]) cudaEventCreate(&d->ev2[i]);
    const int n = ntp * nt;
    cudaEventRecord(d->ev2[0], st);
    k_spline_kernel<<<(n + 127) / 128, 128, 0, st>>>(ntp, nk, nt, d->k, d->src_tr, d->src_ddk, d->spline_u);
    cudaEventRecord(d->ev2[1], st);
    c->launches++;
    CLPP_CUDA(cudaGetLastError(), err);
  }

  // (b) flat Bessel table (hyperspherical_HIS_create with K=0, beta=1)
  const double tau0 = c->bg.conformal_age;
  const int lmax = c->l[TI.l_size_max - 1];
  {
    const double xmax = c->q[TI.q_size - 1] * tau0;
    const double xmin = td.hyper_x_min;
    const double lambda = 2 * CLPP_PI / 1.0;
    int nx = (int)((xmax - xmin) * td.hyper_sampling_flat / lambda);
    nx = std::max(nx, 2);
    const double dx = (xmax - xmin) / (nx - 1.0);
    const int l_rec_max = lmax;  // l_WKB = l_max+1: every l of the list is below it
    const double xfwd = sqrt(l_rec_max * (l_rec_max + 1.0)) / 1.0;
    const int xfwdidx = (int)((xfwd - xmin) / dx);
    d->bessel_nx = nx;
    d->bessel_dx = dx;
    d->bessel_xmin = xmin;
    c->tinfo.x_size = nx;
    const size_t nb = (size_t)TI.l_size_max * nx;
    if (scratch.get(d, &d->bessel_x, (size_t)nx, err) || scratch.get(d, &d->bessel_phi, nb, err) || scratch.get(d, &d->bessel_dphi, nb, err) ||
        dev_alloc(&d->chi_at_phimin, TI.l_size_max, err))
      return CLPP_FAILURE;
    if (scratch.get(d, &d->bessel_scale, nb, err)) return CLPP_FAILURE;
    int* scale_count = d->bessel_scale;
    cudaEventRecord(d->ev2[2], st);
    bessel_table_kernel<<<(nx + 63) / 64, 64, 0, st>>>(nx, xmin, dx, std::min(nx, xfwdidx), lmax + 1, TI.l_size_max, d->l,
                                                      d->bessel_x, d->bessel_phi, d->bessel_dphi, scale_count);
    cudaEventRecord(d->ev2[3], st);
    c->launches++;
    CLPP_CUDA(cudaGetLastError(), err);
    // chi_at_phimin (hyperspherical_get_xmin_from_approx, hyperspherical.c:1419-1457, K=0, nu=1)
    std::vector<double> chi(TI.l_size_max);
    for (int i = 0; i < TI.l_size_max; i++) {
      const double lph = c->l[i] + 0.5;
      const double lhs = 1.0 / lph * log(2 * td.hyper_phi_min_abs * lph);
      const double alpha = -2.0 * lhs / 5.0 * (1.0 + 2.0 * cosh(1.0 / 3.0 * acosh(1.0 + 375.0 / (16.0 * lhs * lhs))));
      chi[i] = lph / cosh(alpha) / 1.0;
    }
    c->chi_host = chi;  // must outlive the asynchronous copy
    CLPP_CUDA(cudaMemcpyAsync(d->chi_at_phimin, c->chi_host.data(), chi.size() * sizeof(double), cudaMemcpyHostToDevice, st), err);
  }

  // (c) line-of-sight integrals
  LosParams P;
  memset(&P, 0, sizeof(P));
  P.nk = nk; P.nt = nt; P.nq = TI.q_size; P.nl = TI.l_size; P.ntt = TI.tt_size; P.nx = d->bessel_nx;
  P.q_begin = q_begin; P.q_end = q_end;
  P.tau0 = tau0; P.tau_rec = c->th.tau_rec; P.tau0_minus_tau_cut = tau0 - c->th.tau_cut;
  P.ra_rec = (tau0 - c->th.tau_rec) * c->th.angular_rescaling;
  P.xmin = d->bessel_xmin; P.dx = d->bessel_dx;
  P.l_late_threshold = td.transfer_neglect_late_source * c->th.angular_rescaling;
  P.l_switch_limber = td.l_switch_limber;
  P.lcmb_rescale = td.lcmb_rescale; P.lcmb_tilt = td.lcmb_tilt; P.lcmb_pivot = td.lcmb_pivot;
  {
    int i = 0;
    while (c->tau[i] <= c->th.tau_rec) i++;
    P.index_tau_min_lcmb = i;
  }
  for (int tt = 0; tt < TI.tt_size; tt++) {
    P.l_size_tt[tt] = c->l_size_tt[tt];
    if (tt == TI.index_tt_t0) { P.tp_of_tt[tt] = PI.index_tp_t0; P.radial[tt] = RADIAL_T0; P.neglect_delta_k[tt] = td.transfer_neglect_delta_k_S_t0; }
    else if (tt == TI.index_tt_t1) { P.tp_of_tt[tt] = PI.index_tp_t1; P.radial[tt] = RADIAL_T1; P.neglect_delta_k[tt] = td.transfer_neglect_delta_k_S_t1; P.late_neglect_ok[tt] = 1; }
    else if (tt == TI.index_tt_t2) { P.tp_of_tt[tt] = PI.index_tp_t2; P.radial[tt] = RADIAL_T2; P.neglect_delta_k[tt] = td.transfer_neglect_delta_k_S_t2; P.late_neglect_ok[tt] = 1; }
    else if (tt == TI.index_tt_e) { P.tp_of_tt[tt] = PI.index_tp_p; P.radial[tt] = RADIAL_E; P.neglect_delta_k[tt] = td.transfer_neglect_delta_k_S_e; P.late_neglect_ok[tt] = 1; }
    else if (tt == TI.index_tt_lcmb) { P.tp_of_tt[tt] = PI.index_tp_phi_plus_psi; P.radial[tt] = RADIAL_T0; P.is_lcmb[tt] = 1; }
    CLPP_CHECK(P.tp_of_tt[tt] >= 0, err, "transfer type %d has no matching source type", tt);
  }
  const size_t ntr = (size_t)TI.tt_size * TI.l_size * TI.q_size;
  if (!d->transfer || d->transfer_count != ntr) {
    if (dev_alloc(&d->transfer, ntr, err)) return CLPP_FAILURE;
    d->transfer_count = ntr;
    CLPP_CUDA(cudaMemsetAsync(d->transfer, 0, ntr * sizeof(double), st), err);
  }
  if (!d->tr_counters) CLPP_CUDA(cudaMalloc((void**)&d->tr_counters, 2 * sizeof(unsigned long long)), err);
  CLPP_CUDA(cudaMemsetAsync(d->tr_counters, 0, 2 * sizeof(unsigned long long), st), err);
  const size_t smem = (size_t)(3 + TI.tt_size) * nt * sizeof(double);
  CLPP_CHECK(smem <= 200 * 1024, err, "tau_size=%d too large for the shared-memory staging of the LOS kernel", nt);
  { static const cudaError_t once = clpp_allow_max_dynamic_smem(los_kernel); CLPP_CUDA(once, err); }
  cudaEventRecord(d->ev[0], st);
  if (q_end > q_begin) {
    los_kernel<<<q_end - q_begin, LOS_THREADS, smem, st>>>(P, d->k, d->tau, d->q, d->kq, d->l, d->chi_at_phimin, d->src_tr,
                                                          d->src_ddk, d->bessel_phi, d->bessel_dphi, d->transfer,
                                                          d->tr_counters);
    c->launches++;
  }
  cudaEventRecord(d->ev[1], st);
  CLPP_CUDA(cudaGetLastError(), err);
  scratch.release();
  unsigned long long cnt[2];
  CLPP_CUDA(cudaMemcpyAsync(cnt, d->tr_counters, sizeof(cnt), cudaMemcpyDeviceToHost, st), err);
  CLPP_CUDA(cudaStreamSynchronize(st), err);
  {
    float ms = 0;
    cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]); d->t_los_ms = ms;
    cudaEventElapsedTime(&ms, d->ev2[0], d->ev2[1]); d->t_kspline_ms = ms;
    cudaEventElapsedTime(&ms, d->ev2[2], d->ev2[3]); d->t_bessel_ms = ms;
  }
  c->tinfo.n_integrals = (long)cnt[0];
  c->tinfo.n_points = (long)cnt[1];
  c->has_transfer = true;
  return CLPP_SUCCESS;
}

int clpp_dev_get_bessel(clpp_ctx* c, double* x, double* phi, double* dphi, double* chi, char* err) {
  clpp_ctx::Dev* d = c->dev;
  CLPP_CHECK(d->bessel_phi != nullptr, err, "the Bessel table of this context has been returned to the memory pool (lean_scratch)");
  const size_t nb = (size_t)c->tinfo.l_size_max * d->bessel_nx;
  if (x) CLPP_CUDA(cudaMemcpy(x, d->bessel_x, d->bessel_nx * sizeof(double), cudaMemcpyDeviceToHost), err);
  if (phi) CLPP_CUDA(cudaMemcpy(phi, d->bessel_phi, nb * sizeof(double), cudaMemcpyDeviceToHost), err);
  if (dphi) CLPP_CUDA(cudaMemcpy(dphi, d->bessel_dphi, nb * sizeof(double), cudaMemcpyDeviceToHost), err);
  if (chi) CLPP_CUDA(cudaMemcpy(chi, d->chi_at_phimin, c->tinfo.l_size_max * sizeof(double), cudaMemcpyDeviceToHost), err);
  return CLPP_SUCCESS;
}
