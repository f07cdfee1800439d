// Internal state of libclpp.so (one clpp_ctx = one cosmology on one CUDA device).
#ifndef CLPP_INTERNAL_H
#define CLPP_INTERNAL_H

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "clpp.h"

// physical constants, same numerical values as the reference
// (include/common.h:66-126, source/thermodynamics.h:372-374)
#define CLPP_PI 3.1415926535897932384626433832795e0
#define CLPP_Mpc_over_m 3.085677581282e22
#define CLPP_c 2.99792458e8
#define CLPP_k_B 1.3806504e-23
#define CLPP_m_H 1.673575e-27
#define CLPP_not4 3.9715
#define CLPP_sigma 6.6524616e-29

// enums mirrored from source/perturbations.h:25-58
enum { CLPP_TCA_FIRST_ORDER_MB = 0, CLPP_TCA_FIRST_ORDER_CAMB, CLPP_TCA_FIRST_ORDER_CLASS,
       CLPP_TCA_SECOND_ORDER_CRS, CLPP_TCA_SECOND_ORDER_CLASS, CLPP_TCA_COMPROMISE_CLASS };
enum { CLPP_RSA_NULL = 0, CLPP_RSA_MD, CLPP_RSA_MD_WITH_REIO, CLPP_RSA_NONE };
enum { CLPP_UFA_MB = 0, CLPP_UFA_HU, CLPP_UFA_CLASS, CLPP_UFA_NONE };
enum { CLPP_NCDMFA_MB = 0, CLPP_NCDMFA_HU, CLPP_NCDMFA_CLASS, CLPP_NCDMFA_NONE };
// reionization_parametrization (source/thermodynamics.h:24-31)
enum { CLPP_REIO_NONE = 0, CLPP_REIO_CAMB, CLPP_REIO_BINS_TANH, CLPP_REIO_HALF_TANH, CLPP_REIO_MANY_TANH,
       CLPP_REIO_INTER };

struct clpp_error {};

inline int clpp_fail(char* err, const char* fmt, ...) {
  if (err) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, CLPP_ERRLEN, fmt, ap);
    va_end(ap);
  }
  return CLPP_FAILURE;
}

#define CLPP_CHECK(cond, err, ...) \
  do { if (!(cond)) return clpp_fail(err, __VA_ARGS__); } while (0)

#define CLPP_CUDA(call, err)                                                          \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess)                                                            \
      return clpp_fail(err, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
  } while (0)

// ---- host-side interpolation tables (tools/arrays.c restatement, see host_tables.cpp) ------
struct HostTable {
  int n_lines = 0, n_cols = 0;
  std::vector<double> x;    // [n_lines]
  std::vector<double> y;    // [n_lines*n_cols]
  std::vector<double> ddy;  // [n_lines*n_cols]
};

// Gauss-Legendre nodes (ascending) and weights on [-1,1] (quadrature_gauss_legendre, tools/quadrature.c:752-788)
int clpp_gauss_legendre(double* mu, double* w8, int n, double tol, char* err);
void clpp_spline_table_lines(const double* x, int x_size, const double* y, int y_size, double* ddy);
// array_interpolate_spline (bisection) / _growing_closeby (cursor)
int clpp_interp_spline(const HostTable& t, double x, int* last_index, double* result, int result_size, char* err);
int clpp_interp_spline_closeby(const HostTable& t, double x, int* last_index, double* result, int result_size,
                               char* err);
int clpp_interp_linear(const HostTable& t, double x, int* last_index, double* result, int result_size, char* err);

struct clpp_ctx {
  int device = -1;  // -1: host-only context (grids only)
  long launches = 0;
  void* stream = nullptr;  // cudaStream_t
  bool lean_scratch = false;  // stage-2 work buffers from the stream-ordered pool, returned after the stage
  int lane_path = -1;         // -1: by batch size; 0 / 1: force the warp-per-mode / thread-per-mode perturbation kernels

  // --- inputs
  bool has_bg = false, has_th = false;
  clpp_background_desc bg{};
  clpp_thermo_desc th{};
  HostTable bgt, tht;
  int N_ncdm = 0;
  std::vector<int> ncdm_q_size;
  std::vector<double> ncdm_q, ncdm_w, ncdm_dlnf0, ncdm_M, ncdm_factor;

  // --- stage 1
  bool has_pgrids = false, has_sources = false;
  clpp_perturb_desc pd{};
  clpp_perturb_info pinfo{};
  std::vector<double> k, tau;
  std::vector<clpp_kstat> kstat;

  // --- stage 2
  bool has_tgrids = false, has_transfer = false;
  bool nl_dev_valid = false;  // a device-resident halofit correction matches the current sources
  int nl_dev_nk = 0, nl_dev_nt = 0;  // grid the device-resident correction was built for
  bool use_device_nl = true;  // clpp_transfer_compute(nl_corr_density = NULL) applies it (option "use_device_nl")
  clpp_transfer_desc td{};
  clpp_transfer_info tinfo{};
  std::vector<int> l, l_size_tt;
  std::vector<double> q, kq, chi_host;

  // --- stage 3: C_l table on the l grid + its spline along l (spectra_cl_at_l)
  bool has_cl = false;
  HostTable clt;
  int cl_l_max = 0;  // l_max_ct: C_l are zero above it
  clpp_spectra_info sinfo{};

  // --- lensing: lensed C_l table on its l grid + spline along l (lensing_cl_at_l)
  bool has_cl_lens = false;
  HostTable cl_lens;
  int l_lensed_max = 0;
  std::vector<double> lens_stage, gl_nodes;  // host staging; cached Gauss-Legendre nodes | weights
  int gl_n = 0;
  double gl_tol = 0.;

  // --- device memory (managed in device.cu)
  struct Dev;
  Dev* dev = nullptr;
};

// host-side background_at_tau / thermodynamics_at_z
// (background_module.cpp:125-199, thermodynamics_module.cpp:114-285)
enum { CLPP_INTER_NORMAL = 0, CLPP_INTER_CLOSEBY = 1 };
int clpp_background_at_tau(const clpp_ctx* c, double tau, int size, int mode, int* last_index, double* pvecback,
                           char* err);
int clpp_thermodynamics_at_z(const clpp_ctx* c, double z, int mode, int* last_index, const double* pvecback,
                             double* pvecthermo, char* err);

// grids (grids.cpp)
int clpp_host_perturb_grids(clpp_ctx* c, char* err);
int clpp_host_transfer_grids(clpp_ctx* c, char* err);

// device side (implemented in the .cu files)
int clpp_dev_create(clpp_ctx* c, char* err);
void clpp_dev_destroy(clpp_ctx* c);
int clpp_dev_upload_tables(clpp_ctx* c, char* err);

#endif
