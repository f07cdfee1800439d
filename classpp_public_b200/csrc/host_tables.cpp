// Host-side table machinery of the path: cubic-spline second derivatives and the two
// interpolators the reference calls everywhere on the path.
//
// Everything here must be BIT-IDENTICAL to the reference, because the k / tau / q sampling
// grids are built from interpolated background/thermodynamics values and the grids are
// required to match bit for bit (SURVEY.md 8a', 8d).  Hence: same arithmetic expressions in
// the same association order as tools/arrays.c, compiled without FMA contraction.
//
//   clpp_spline_table_lines      <-> array_spline_table_lines (tools/arrays.c:514-660), _SPLINE_EST_DERIV_
//   clpp_interp_spline           <-> array_interpolate_spline (tools/arrays.c:1565-1634)
//   clpp_interp_spline_closeby   <-> array_interpolate_spline_growing_closeby (tools/arrays.c:2173-2232)
//   clpp_interp_linear           <-> array_interpolate_linear (tools/arrays.c:1690-1750)
//   clpp_background_at_tau       <-> BackgroundModule::background_at_tau (background_module.cpp:125-199)
//   clpp_thermodynamics_at_z     <-> ThermodynamicsModule::thermodynamics_at_z (thermodynamics_module.cpp:114-285)
#include <cmath>

#include "clpp_internal.h"

// Second derivatives of every column of a row-major table y[x_size][y_size] with the
// "estimated first derivative" end conditions: the end slopes come from the parabola through
// the first (last) three nodes, then the usual tridiagonal sweep (forward elimination storing
// the decomposition factor in ddy and the reduced right-hand side in u, then back-substitution).
void clpp_spline_table_lines(const double* x, int x_size, const double* y, int y_size, double* ddy) {
  const bool natural = (x_size == 2);
  std::vector<double> u((size_t)(x_size - 1) * y_size);
  auto Y = [&](int ix, int iy) { return y[(size_t)ix * y_size + iy]; };

  // first node
  for (int c = 0; c < y_size; c++) {
    if (natural) {
      ddy[c] = 0.0;
      u[c] = 0.0;
    } else {
      const double slope0 = ((x[2] - x[0]) * (x[2] - x[0]) * (Y(1, c) - Y(0, c)) -
                             (x[1] - x[0]) * (x[1] - x[0]) * (Y(2, c) - Y(0, c))) /
                            ((x[2] - x[0]) * (x[1] - x[0]) * (x[2] - x[1]));
      ddy[c] = -0.5;
      u[c] = (3. / (x[1] - x[0])) * ((Y(1, c) - Y(0, c)) / (x[1] - x[0]) - slope0);
    }
  }
  // interior nodes: forward elimination
  for (int i = 1; i < x_size - 1; i++) {
    const double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
    double* d_i = ddy + (size_t)i * y_size;
    const double* d_m = ddy + (size_t)(i - 1) * y_size;
    double* u_i = u.data() + (size_t)i * y_size;
    const double* u_m = u.data() + (size_t)(i - 1) * y_size;
    for (int c = 0; c < y_size; c++) {
      const double p = sig * d_m[c] + 2.0;
      d_i[c] = (sig - 1.0) / p;
      double t = (Y(i + 1, c) - Y(i, c)) / (x[i + 1] - x[i]) - (Y(i, c) - Y(i - 1, c)) / (x[i] - x[i - 1]);
      u_i[c] = (6.0 * t / (x[i + 1] - x[i - 1]) - sig * u_m[c]) / p;
    }
  }
  // last node
  const int n = x_size;
  for (int c = 0; c < y_size; c++) {
    double qn, un;
    if (natural) {
      qn = un = 0.0;
    } else {
      const double slopeN = ((x[n - 3] - x[n - 1]) * (x[n - 3] - x[n - 1]) * (Y(n - 2, c) - Y(n - 1, c)) -
                             (x[n - 2] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (Y(n - 3, c) - Y(n - 1, c))) /
                            ((x[n - 3] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (x[n - 3] - x[n - 2]));
      qn = 0.5;
      un = (3. / (x[n - 1] - x[n - 2])) * (slopeN - (Y(n - 1, c) - Y(n - 2, c)) / (x[n - 1] - x[n - 2]));
    }
    ddy[(size_t)(n - 1) * y_size + c] =
        (un - qn * u[(size_t)(n - 2) * y_size + c]) / (qn * ddy[(size_t)(n - 2) * y_size + c] + 1.0);
  }
  // back-substitution
  for (int i = n - 2; i >= 0; i--) {
    double* d_i = ddy + (size_t)i * y_size;
    const double* d_p = ddy + (size_t)(i + 1) * y_size;
    const double* u_i = u.data() + (size_t)i * y_size;
    for (int c = 0; c < y_size; c++) d_i[c] = d_i[c] * d_p[c] + u_i[c];
  }
}

static inline void cubic_eval(const HostTable& t, int inf, int sup, double x, double* result, int result_size) {
  const double h = t.x[sup] - t.x[inf];
  const double b = (x - t.x[inf]) / h;
  const double a = 1 - b;
  const double* y0 = &t.y[(size_t)inf * t.n_cols];
  const double* y1 = &t.y[(size_t)sup * t.n_cols];
  const double* d0 = &t.ddy[(size_t)inf * t.n_cols];
  const double* d1 = &t.ddy[(size_t)sup * t.n_cols];
  for (int i = 0; i < result_size; i++)
    result[i] = a * y0[i] + b * y1[i] + ((a * a * a - a) * d0[i] + (b * b * b - b) * d1[i]) * h * h / 6.;
}

static int bisect(const HostTable& t, double x, int* inf_out, int* sup_out, char* err) {
  int inf = 0, sup = t.n_lines - 1;
  if (t.x[inf] < t.x[sup]) {
    CLPP_CHECK(!(x < t.x[inf]), err, "x=%e < x_min=%e", x, t.x[inf]);
    CLPP_CHECK(!(x > t.x[sup]), err, "x=%e > x_max=%e", x, t.x[sup]);
    while (sup - inf > 1) {
      int mid = (int)(0.5 * (inf + sup));
      if (x < t.x[mid]) sup = mid; else inf = mid;
    }
  } else {
    CLPP_CHECK(!(x < t.x[sup]), err, "x=%e < x_min=%e", x, t.x[sup]);
    CLPP_CHECK(!(x > t.x[inf]), err, "x=%e > x_max=%e", x, t.x[inf]);
    while (sup - inf > 1) {
      int mid = (int)(0.5 * (inf + sup));
      if (x > t.x[mid]) sup = mid; else inf = mid;
    }
  }
  *inf_out = inf;
  *sup_out = sup;
  return CLPP_SUCCESS;
}

int clpp_interp_spline(const HostTable& t, double x, int* last_index, double* result, int result_size, char* err) {
  int inf, sup;
  if (bisect(t, x, &inf, &sup, err)) return CLPP_FAILURE;
  *last_index = inf;
  cubic_eval(t, inf, sup, x, result, result_size);
  return CLPP_SUCCESS;
}

int clpp_interp_linear(const HostTable& t, double x, int* last_index, double* result, int result_size, char* err) {
  int inf, sup;
  if (bisect(t, x, &inf, &sup, err)) return CLPP_FAILURE;
  *last_index = inf;
  const double h = t.x[sup] - t.x[inf];
  const double b = (x - t.x[inf]) / h;
  const double a = 1 - b;
  for (int i = 0; i < result_size; i++)
    result[i] = a * t.y[(size_t)inf * t.n_cols + i] + b * t.y[(size_t)sup * t.n_cols + i];
  return CLPP_SUCCESS;
}

int clpp_interp_spline_closeby(const HostTable& t, double x, int* last_index, double* result, int result_size,
                               char* err) {
  int inf = *last_index;
  CLPP_CHECK(inf >= 0 && inf <= t.n_lines - 1, err, "*lastindex=%d out of range [0:%d]", inf, t.n_lines - 1);
  while (x < t.x[inf]) {
    inf--;
    CLPP_CHECK(inf >= 0, err, "x=%e < x_min=%e", x, t.x[0]);
  }
  int sup = inf + 1;
  while (x > t.x[sup]) {
    sup++;
    CLPP_CHECK(sup <= t.n_lines - 1, err, "x=%e > x_max=%e", x, t.x[t.n_lines - 1]);
  }
  inf = sup - 1;
  *last_index = inf;
  cubic_eval(t, inf, sup, x, result, result_size);
  return CLPP_SUCCESS;
}

int clpp_background_at_tau(const clpp_ctx* c, double tau, int size, int mode, int* last_index, double* pvecback,
                           char* err) {
  const HostTable& t = c->bgt;
  CLPP_CHECK(!(tau < t.x[0]), err,
             "out of range: tau=%e < tau_min=%e, you should decrease the precision parameter a_ini_over_a_today_default",
             tau, t.x[0]);
  CLPP_CHECK(!(tau > t.x[t.n_lines - 1]), err, "out of range: tau=%e > tau_max=%e", tau, t.x[t.n_lines - 1]);
  if (mode == CLPP_INTER_NORMAL) return clpp_interp_spline(t, tau, last_index, pvecback, size, err);
  return clpp_interp_spline_closeby(t, tau, last_index, pvecback, size, err);
}

int clpp_thermodynamics_at_z(const clpp_ctx* c, double z, int mode, int* last_index, const double* pvecback,
                             double* pv, char* err) {
  const HostTable& t = c->tht;
  const clpp_thermo_desc& th = c->th;
  const clpp_background_desc& bg = c->bg;
  const int last = t.n_lines - 1;
  if (z >= t.x[last]) {
    // beyond the table: analytic early-time scalings (constant x_e, kappa' ~ (1+z)^2, T_b = T_cmb(1+z))
    const double* row = &t.y[(size_t)last * t.n_cols];
    const double x0 = row[th.index_th_xe];
    const double H = pvecback[bg.index_bg_H], Hp = pvecback[bg.index_bg_H_prime];
    pv[th.index_th_xe] = x0;
    pv[th.index_th_dkappa] = (1. + z) * (1. + z) * th.n_e * x0 * CLPP_sigma * CLPP_Mpc_over_m;
    pv[th.index_th_tau_d] = row[th.index_th_tau_d] * pow((1 + z) / (1. + t.x[last]), 2);
    if (th.compute_damping_scale) pv[th.index_th_r_d] = row[th.index_th_r_d] * pow((1 + z) / (1. + t.x[last]), -1.5);
    pv[th.index_th_ddkappa] = -H * 2. / (1. + z) * pv[th.index_th_dkappa];
    pv[th.index_th_dddkappa] = (H * H / (1. + z) - Hp) * 2. / (1. + z) * pv[th.index_th_dkappa];
    pv[th.index_th_exp_m_kappa] = 0.;
    pv[th.index_th_g] = 0.;
    pv[th.index_th_dg] = 0.;
    pv[th.index_th_ddg] = 0.;
    pv[th.index_th_Tb] = bg.T_cmb * (1. + z);
    pv[th.index_th_wb] = CLPP_k_B / (CLPP_c * CLPP_c * CLPP_m_H) *
                         (1. + (1. / CLPP_not4 - 1.) * th.YHe + x0 * (1. - th.YHe)) * bg.T_cmb * (1. + z);
    pv[th.index_th_cb2] = pv[th.index_th_wb] * 4. / 3.;
    if (th.compute_cb2_derivatives) {
      pv[th.index_th_dcb2] = -H * pvecback[bg.index_bg_a] * pv[th.index_th_cb2];
      pv[th.index_th_ddcb2] = -Hp * pvecback[bg.index_bg_a] * pv[th.index_th_cb2];
    }
    pv[th.index_th_rate] = pv[th.index_th_dkappa];
    return CLPP_SUCCESS;
  }
  if (((th.reio_parametrization == CLPP_REIO_HALF_TANH) && (z < 2 * th.z_reionization)) ||
      ((th.reio_parametrization == CLPP_REIO_INTER) && (z < 50.)))
    return clpp_interp_linear(t, z, last_index, pv, t.n_cols, err);
  if (mode == CLPP_INTER_NORMAL) return clpp_interp_spline(t, z, last_index, pv, t.n_cols, err);
  return clpp_interp_spline_closeby(t, z, last_index, pv, t.n_cols, err);
}


// Gauss-Legendre rule with n nodes on [-1,1] for the accurate lensing mode (lensing_module.cpp:237-248 calls
// quadrature_gauss_legendre, tools/quadrature.c:752-788): each root of P_n in the upper half is refined by Newton's method
// from the Chebyshev-like guess cos(pi (i - 1/4) / (n + 1/2)) until the update is below `tol`; the rule is symmetric.
// Nodes are returned in ascending order, weights w = 2 / ((1 - x^2) P_n'(x)^2).
// This is the textbook Newton iteration on the three-term recurrence (Numerical Recipes' `gauleg`, which the reference's
// routine also is): the same algorithm in the same evaluation order -- the nodes must be bit-identical to the reference's for
// the accurate lensing mode to reproduce its C_l -- so it is the standard routine restated, not an independent design.
int clpp_gauss_legendre(double* mu, double* w8, int n, double tol, char* err) {
  const int half = (n + 1) / 2;
  for (int i = 0; i < half; i++) {
    double x = cos(CLPP_PI * ((double)(i + 1) - 0.25) / ((double)n + 0.5));
    double dpn = 0.;
    for (int it = 0;; it++) {
      if (it == 10000) return clpp_fail(err, "maximum number of iteration reached: increase either _MAX_IT_ or tol\n");
      double pn = 1., pnm1 = 0.;  // P_j(x), P_{j-1}(x) by the three-term recurrence
      for (int j = 1; j <= n; j++) {
        const double pnm2 = pnm1;
        pnm1 = pn;
        pn = ((2.0 * j - 1.0) * x * pnm1 - (j - 1.0) * pnm2) / j;
      }
      dpn = n * (x * pn - pnm1) / (x * x - 1.0);
      const double x_old = x;
      x = x_old - pn / dpn;
      if (fabs(x - x_old) <= tol) break;
    }
    mu[i] = -x;
    mu[n - 1 - i] = x;
    w8[i] = w8[n - 1 - i] = 2.0 / ((1.0 - x * x) * dpn * dpn);
  }
  return CLPP_SUCCESS;
}
