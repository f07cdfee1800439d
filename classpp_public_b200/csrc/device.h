// Device-resident state of a clpp_ctx (HBM layout, see DESIGN.md "Data layout in HBM").
#ifndef CLPP_DEVICE_H
#define CLPP_DEVICE_H

#include <cuda_runtime.h>

#include <map>

#include "clpp_internal.h"

#define CLPP_PT_MAX_CHUNKS 16

struct clpp_ctx::Dev {
  cudaStream_t stream = nullptr, stream2 = nullptr, stream_hi = nullptr;
  cudaStream_t chunk_stream[CLPP_PT_MAX_CHUNKS] = {};  // low-priority streams of the bulk chunks of a perturbation launch
  cudaEvent_t chunk_done[CLPP_PT_MAX_CHUNKS] = {};  // stream2: second group of the perturbation launch
  int sm_count = 0;
  // per-kernel device timings of the last stage calls (CUDA events on `stream`), in ms
  cudaEvent_t ev[2] = {nullptr, nullptr};
  double t_perturb_ms = 0., t_perturb_tail_ms = 0., t_kspline_ms = 0., t_bessel_ms = 0., t_los_ms = 0., t_spectra_ms = 0.;

  // upstream tables: row-major [n_lines][n_cols] + second derivatives, L2-resident (~6 MB)
  double *bg_tau = nullptr, *bg_y = nullptr, *bg_dd = nullptr;
  double *th_z = nullptr, *th_y = nullptr, *th_dd = nullptr;

  // stage 1
  double *k = nullptr, *tau = nullptr;
  double* sources = nullptr;  // [tp][k][tau], tau fastest (one k-mode writes contiguous rows)
  clpp_kstat* kstat = nullptr;
  int* k_order = nullptr;     // work queue: mode indices sorted by decreasing expected cost
  int* queue_head = nullptr;  // atomic cursor into k_order
  double* jac_scratch = nullptr;  // per-CTA global workspace: hub block of the Jacobian
  double* ncdm = nullptr;         // [3][nq_tot]: q, w, dlnf0/dlnq
  double* lane_scratch = nullptr; // per-thread slabs of the lane kernel (lane.cuh), global-memory fallback only
  unsigned char* ln_modes = nullptr;  // (cosmology, k) list of the lane kernel
  cudaStream_t lane_stream = nullptr;
  cudaEvent_t lane_done = nullptr, lane_go = nullptr;
  cudaStream_t tlane_stream = nullptr;  // lane tails (perturb_tail_lane_kernel) of a group that also has warp tails
  cudaEvent_t tlane_go = nullptr, tlane_done = nullptr;
  double* i2l1 = nullptr;         // 1/(2l+1)
  double* pt_tail = nullptr;      // hand-off records perturb_kernel -> perturb_tail_kernel
  size_t pt_tail_cap = 0;
  unsigned char *pt_cosmo = nullptr, *pt_modes = nullptr;  // batch descriptors of the last solve (PtCosmo[], int2[])
  size_t sources_count = 0;
  size_t k_cap = 0, tau_cap = 0, kstat_cap = 0, jac_cap = 0, ncdm_cap = 0, pt_cosmo_cap = 0, pt_modes_cap = 0;

  // stage 2
  double *q = nullptr, *kq = nullptr;
  int *l = nullptr;
  double *bessel_x = nullptr, *bessel_phi = nullptr, *bessel_dphi = nullptr, *chi_at_phimin = nullptr;
  int bessel_nx = 0;
  double bessel_dx = 0., bessel_xmin = 0.;
  double* src_tr = nullptr;   // sources seen by the transfer stage (after nl correction) [tp][k][tau]
  double* src_ddk = nullptr;  // d2S/dk2 for the cubic spline in k, same layout
  double* nl_corr = nullptr;
  double* nl_corr2 = nullptr;  // halofit on the device: R_NL [spec][k][tau] (spec 0 = total matter)
  int* hf_flags = nullptr;
  double t_halofit_ms = 0.;
  double* transfer = nullptr;  // [tt][l][q], q fastest (= reference layout)
  size_t transfer_count = 0;
  unsigned long long* tr_counters = nullptr;

  // stage-2 scratch kept across calls (no cudaMalloc/cudaFree, which synchronise the device, on the hot path)
  double* spline_u = nullptr;
  int* bessel_scale = nullptr;
  cudaEvent_t ev2[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::map<void*, size_t> cap;  // capacity (bytes) of the buffers managed by dev_reserve_any

  // stage 3
  double *pk = nullptr, *wq = nullptr, *cl = nullptr;

  // lensing
  double *lens_stage = nullptr, *lens_work = nullptr, *lens_coef = nullptr;
  int* lens_lgrid = nullptr;
  int lens_lmax = -1;  // l_max the recurrence-coefficient table was built for
  double t_lensing_ms = 0.;
};

// Dynamic shared memory above 48 KB needs a per-function opt-in that is PROCESS-GLOBAL state: raise it to the device limit
// (227 KB) once instead of setting the size of each call -- concurrent stage calls of different contexts (sweeps run one
// host thread per cosmology) would otherwise lower each other's limit between the attribute call and the launch.
template <typename K>
inline cudaError_t clpp_allow_max_dynamic_smem(K kernel) {
  int dev = 0, optin = 0;
  cudaFuncAttributes fa;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, kernel);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

// grow-only device buffer: reallocates only when the requested size exceeds the capacity.  Every allocation carries 25 %
// of headroom: in a sweep the contexts are reused for cosmology after cosmology whose grids differ by a few per cent
// (k_size, tau_size, bt_size, the Bessel x range), and a reallocation is a cudaFree -- which waits for the whole device,
// i.e. for the batched perturbation launch in flight -- so without headroom the host stages meant to run under that launch
// queue up behind it (profiles/r02_config5_lhs1024.json: launches of 9-13 s in a real sweep against 8 s in the bench, whose
// batch repeats).
template <typename T>
inline int clpp_dev_reserve(clpp_ctx::Dev* d, T** p, size_t n, char* err) {
  const size_t bytes = (n > 0 ? n : 1) * sizeof(T);
  auto it = d->cap.find((void*)p);
  if (*p && it != d->cap.end() && it->second >= bytes) return CLPP_SUCCESS;
  if (*p) { cudaFree(*p); *p = nullptr; }
  const size_t padded = (bytes + bytes / 4 + 255) & ~(size_t)255;
  cudaError_t e = cudaMalloc((void**)p, padded);
  if (e != cudaSuccess) return clpp_fail(err, "cudaMalloc of %zu bytes failed: %s", padded, cudaGetErrorString(e));
  d->cap[(void*)p] = padded;
  return CLPP_SUCCESS;
}

#endif
