// C-ABI glue of libclpp.so: context management, table upload, grid entry points, accessors.
// The numerics live in perturb.cu / transfer.cu / spectra.cu (device) and grids.cpp /
// host_tables.cpp (host, bit-exact grids).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>

#include "device.h"

// implemented in the stage files
int clpp_dev_perturb_solve(clpp_ctx* c, int k_begin, int k_end, char* err);
int clpp_dev_perturb_solve_batch(clpp_ctx** cs, int n_ctx, const int* k_begin, const int* k_end, const int* k_list,
                                 int n_list, char* err);
int clpp_dev_perturb_solve_list(clpp_ctx* c, const int* k_list, int n, char* err);
int clpp_dev_transfer_compute(clpp_ctx* c, const double* nl_corr_density, int q_begin, int q_end, char* err);
int clpp_dev_spectra(clpp_ctx* c, const double* primordial_pk, int q_begin, int q_end, clpp_spectra_info* info,
                     double* cl_out, char* err);
int clpp_host_cl_at_l(const clpp_ctx* c, double l, double* cl_tot, char* err);
int clpp_dev_halofit(clpp_ctx* c, const clpp_halofit_desc* hd, const double* primordial_pk, double* nl_corr_out,
                     int* index_tau_min_nl, char* err);
int clpp_dev_pk_linear(clpp_ctx* c, const double* primordial_pk, int index_tau, int cb, double* pk_out, char* err);
int clpp_dev_sources_at_tau(clpp_ctx* c, int index_tp, double tau, double* psource, char* err);
int clpp_dev_lensing(clpp_ctx* c, const clpp_lensing_desc* ld, clpp_lensing_info* info, double* l_out, double* cl_lens_out,
                     char* err);
int clpp_host_lensing_cl_at_l(const clpp_ctx* c, int l, double* cl_lensed, char* err);
int clpp_dev_get_bessel(clpp_ctx* c, double* x, double* phi, double* dphi, double* chi, char* err);

template <typename T>
static int upload(clpp_ctx::Dev* d, T** dptr, const T* src, size_t n, cudaStream_t s, char* err) {
  if (n == 0) return CLPP_SUCCESS;
  if (clpp_dev_reserve(d, dptr, n, err)) return CLPP_FAILURE;  // grow-only: no cudaFree/cudaMalloc when re-used
  if (src) CLPP_CUDA(cudaMemcpyAsync(*dptr, src, n * sizeof(T), cudaMemcpyHostToDevice, s), err);
  return CLPP_SUCCESS;
}

extern "C" {

const char* clpp_version(void) { return "clpp-b200 0.1 (sm_100a)"; }

int clpp_ctx_create(int device, clpp_ctx** out, char* err) {
  CLPP_CHECK(out != nullptr, err, "null ctx pointer");
  *out = nullptr;
  clpp_ctx* c = new clpp_ctx();
  c->device = device;
  if (device >= 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= device) {
      delete c;
      return clpp_fail(err, "no usable CUDA device %d (%s): the B200 path has no CPU fallback", device,
                       e == cudaSuccess ? "device index out of range" : cudaGetErrorString(e));
    }
    cudaSetDevice(device);
    c->dev = new clpp_ctx::Dev();
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->dev->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->dev->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c->dev;
      delete c;
      return clpp_fail(err, "cudaStreamCreate failed");
    }
    c->stream = c->dev->stream;
    cudaEventCreate(&c->dev->ev[0]);
    cudaEventCreate(&c->dev->ev[1]);
  }
  *out = c;
  return CLPP_SUCCESS;
}

void clpp_ctx_destroy(clpp_ctx* c) {
  if (!c) return;
  if (c->dev) {
    cudaSetDevice(c->device);
    clpp_ctx::Dev* d = c->dev;
    cudaStreamSynchronize(d->stream);
    void* ptrs[] = {d->bg_tau, d->bg_y, d->bg_dd, d->th_z, d->th_y, d->th_dd, d->k, d->tau, d->sources, d->kstat,
                    d->k_order, d->queue_head, d->jac_scratch, d->q, d->kq, d->l, d->bessel_x, d->bessel_phi,
                    d->bessel_dphi, d->chi_at_phimin, d->src_tr, d->src_ddk, d->nl_corr, d->transfer, d->tr_counters,
                    d->pk, d->wq, d->cl, d->ncdm, d->pt_cosmo, d->pt_modes, d->spline_u, d->bessel_scale, d->pt_tail, d->nl_corr2, d->hf_flags,
                    d->lens_stage, d->lens_work, d->lens_coef, d->lens_lgrid};
    for (void* p : ptrs)
      if (p) cudaFree(p);
    for (int i = 0; i < 6; i++)
      if (d->ev2[i]) cudaEventDestroy(d->ev2[i]);
    cudaEventDestroy(d->ev[0]);
    cudaEventDestroy(d->ev[1]);
    if (d->stream2) cudaStreamDestroy(d->stream2);
    if (d->stream_hi) cudaStreamDestroy(d->stream_hi);
    if (d->lane_stream) cudaStreamDestroy(d->lane_stream);
    if (d->lane_done) cudaEventDestroy(d->lane_done);
    if (d->lane_go) cudaEventDestroy(d->lane_go);
    if (d->tlane_stream) cudaStreamDestroy(d->tlane_stream);
    if (d->tlane_go) cudaEventDestroy(d->tlane_go);
    if (d->tlane_done) cudaEventDestroy(d->tlane_done);
    for (int i = 0; i < CLPP_PT_MAX_CHUNKS; i++) {
      if (d->chunk_stream[i]) cudaStreamDestroy(d->chunk_stream[i]);
      if (d->chunk_done[i]) cudaEventDestroy(d->chunk_done[i]);
    }
    cudaStreamDestroy(d->stream);
    delete d;
  }
  delete c;
}

long clpp_ctx_launch_count(const clpp_ctx* c) { return c ? c->launches : 0; }

int clpp_ctx_get_stream(clpp_ctx* c, void** stream) {
  if (!c || !c->dev || !stream) return CLPP_FAILURE;
  *stream = (void*)c->dev->stream;
  return CLPP_SUCCESS;
}

int clpp_ctx_get_kernel_ms(const clpp_ctx* c, double out[8]) {
  if (!c || !c->dev) return CLPP_FAILURE;
  out[0] = c->dev->t_perturb_ms; out[1] = c->dev->t_kspline_ms; out[2] = c->dev->t_bessel_ms;
  out[3] = c->dev->t_los_ms; out[4] = c->dev->t_spectra_ms; out[5] = c->dev->t_perturb_tail_ms; out[6] = c->dev->t_halofit_ms;
  out[7] = c->dev->t_lensing_ms;
  return CLPP_SUCCESS;
}

__global__ void dfma_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6.,
         a7 = a0 + 7.;
  const double b = 1.0000001, cc = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, cc); a1 = fma(a1, b, cc); a2 = fma(a2, b, cc); a3 = fma(a3, b, cc);
    a4 = fma(a4, b, cc); a5 = fma(a5, b, cc); a6 = fma(a6, b, cc); a7 = fma(a7, b, cc);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int clpp_ctx_set_option(clpp_ctx* ctx, const char* name, double value, char* err) {
  CLPP_CHECK(ctx && name, err, "null context / option name");
  const std::string n(name);
  if (n == "lean_scratch") ctx->lean_scratch = value != 0.;
  else if (n == "lane_path") ctx->lane_path = (int)value;
  else if (n == "use_device_nl") ctx->use_device_nl = value != 0.;
  else return clpp_fail(err, "unknown context option '%s'", name);
  return CLPP_SUCCESS;
}

int clpp_measure_fp64_peak(clpp_ctx* c, double* tflops, char* err) {
  CLPP_CHECK(c && c->dev && tflops, err, "no CUDA device");
  cudaSetDevice(c->device);
  clpp_ctx::Dev* d = c->dev;
  const int blocks = d->sm_count * 8, threads = 256, iters = 20000;
  double* out = nullptr;
  CLPP_CUDA(cudaMalloc((void**)&out, (size_t)blocks * threads * sizeof(double)), err);
  double best = 0.;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(d->ev[0], d->stream);
    dfma_peak_kernel<<<blocks, threads, 0, d->stream>>>(out, iters);
    cudaEventRecord(d->ev[1], d->stream);
    CLPP_CUDA(cudaStreamSynchronize(d->stream), err);
    float ms = 0;
    cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]);
    const double fl = 2.0 * 8.0 * iters * (double)blocks * threads;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaFree(out);
  *tflops = best;
  return CLPP_SUCCESS;
}

int clpp_set_background(clpp_ctx* c, const clpp_background_desc* desc, const double* tau_table,
                        const double* background_table, char* err) {
  CLPP_CHECK(c && desc && tau_table && background_table, err, "null argument");
  CLPP_CHECK(desc->bt_size >= 3 && desc->bg_size >= 1, err, "background table too small (%d x %d)", desc->bt_size,
             desc->bg_size);
  c->bg = *desc;
  HostTable& t = c->bgt;
  t.n_lines = desc->bt_size;
  t.n_cols = desc->bg_size;
  t.x.assign(tau_table, tau_table + t.n_lines);
  t.y.assign(background_table, background_table + (size_t)t.n_lines * t.n_cols);
  t.ddy.resize(t.y.size());
  clpp_spline_table_lines(t.x.data(), t.n_lines, t.y.data(), t.n_cols, t.ddy.data());
  c->has_bg = true;
  if (c->dev) {
    cudaSetDevice(c->device);
    clpp_ctx::Dev* d = c->dev;
    if (upload(d, &d->bg_tau, t.x.data(), t.x.size(), d->stream, err)) return CLPP_FAILURE;
    if (upload(d, &d->bg_y, t.y.data(), t.y.size(), d->stream, err)) return CLPP_FAILURE;
    if (upload(d, &d->bg_dd, t.ddy.data(), t.ddy.size(), d->stream, err)) return CLPP_FAILURE;
  }
  return CLPP_SUCCESS;
}

int clpp_set_thermo(clpp_ctx* c, const clpp_thermo_desc* desc, const double* z_table, const double* th_table,
                    char* err) {
  CLPP_CHECK(c && desc && z_table && th_table, err, "null argument");
  CLPP_CHECK(desc->tt_size >= 3 && desc->th_size >= 1, err, "thermodynamics table too small");
  c->th = *desc;
  HostTable& t = c->tht;
  t.n_lines = desc->tt_size;
  t.n_cols = desc->th_size;
  t.x.assign(z_table, z_table + t.n_lines);
  t.y.assign(th_table, th_table + (size_t)t.n_lines * t.n_cols);
  t.ddy.resize(t.y.size());
  clpp_spline_table_lines(t.x.data(), t.n_lines, t.y.data(), t.n_cols, t.ddy.data());
  c->has_th = true;
  if (c->dev) {
    cudaSetDevice(c->device);
    clpp_ctx::Dev* d = c->dev;
    if (upload(d, &d->th_z, t.x.data(), t.x.size(), d->stream, err)) return CLPP_FAILURE;
    if (upload(d, &d->th_y, t.y.data(), t.y.size(), d->stream, err)) return CLPP_FAILURE;
    if (upload(d, &d->th_dd, t.ddy.data(), t.ddy.size(), d->stream, err)) return CLPP_FAILURE;
  }
  return CLPP_SUCCESS;
}

int clpp_set_ncdm(clpp_ctx* c, int N_ncdm, const int* q_size, const double* q, const double* w,
                  const double* dlnf0_dlnq, const double* M, const double* factor, char* err) {
  CLPP_CHECK(c, err, "null ctx");
  CLPP_CHECK(N_ncdm >= 0 && N_ncdm <= CLPP_MAX_NCDM, err, "N_ncdm=%d out of range [0,%d]", N_ncdm, CLPP_MAX_NCDM);
  c->N_ncdm = N_ncdm;
  c->ncdm_q_size.assign(q_size, q_size + N_ncdm);
  size_t tot = 0;
  for (int n = 0; n < N_ncdm; n++) tot += q_size[n];
  c->ncdm_q.assign(q, q + tot);
  c->ncdm_w.assign(w, w + tot);
  c->ncdm_dlnf0.assign(dlnf0_dlnq, dlnf0_dlnq + tot);
  c->ncdm_M.assign(M, M + N_ncdm);
  c->ncdm_factor.assign(factor, factor + N_ncdm);
  return CLPP_SUCCESS;
}

// ---- stage 1 ---------------------------------------------------------------------------------
int clpp_perturb_grids(clpp_ctx* c, const clpp_perturb_desc* desc, clpp_perturb_info* info, char* err) {
  CLPP_CHECK(c && desc, err, "null argument");
  c->pd = *desc;
  c->has_sources = false;
  c->nl_dev_valid = false;  // a device-resident halofit table belongs to the previous sources
  if (clpp_host_perturb_grids(c, err)) return CLPP_FAILURE;
  if (info) *info = c->pinfo;
  return CLPP_SUCCESS;
}

int clpp_perturb_solve(clpp_ctx* c, int k_begin, int k_end, char* err) {
  CLPP_CHECK(c && c->has_pgrids, err, "clpp_perturb_grids must be called first");
  CLPP_CHECK(c->dev, err, "this context has no CUDA device: the B200 path has no CPU fallback");
  CLPP_CHECK(0 <= k_begin && k_begin <= k_end && k_end <= c->pinfo.k_size, err, "bad k range [%d,%d)", k_begin, k_end);
  cudaSetDevice(c->device);
  return clpp_dev_perturb_solve(c, k_begin, k_end, err);
}

int clpp_perturb_solve_batch(clpp_ctx** ctxs, int n_ctx, char* err) {
  CLPP_CHECK(ctxs && n_ctx >= 1, err, "empty batch");
  std::vector<int> kb(n_ctx, 0), ke(n_ctx, 0);
  for (int b = 0; b < n_ctx; b++) {
    clpp_ctx* c = ctxs[b];
    CLPP_CHECK(c && c->has_pgrids, err, "clpp_perturb_grids must be called first (cosmology %d of the batch)", b);
    CLPP_CHECK(c->dev, err, "this context has no CUDA device: the B200 path has no CPU fallback");
    ke[b] = c->pinfo.k_size;
  }
  cudaSetDevice(ctxs[0]->device);
  return clpp_dev_perturb_solve_batch(ctxs, n_ctx, kb.data(), ke.data(), nullptr, 0, err);
}

int clpp_perturb_solve_list(clpp_ctx* c, const int* k_indices, int n, char* err) {
  CLPP_CHECK(c && c->has_pgrids, err, "clpp_perturb_grids must be called first");
  CLPP_CHECK(c->dev, err, "this context has no CUDA device: the B200 path has no CPU fallback");
  CLPP_CHECK(k_indices != nullptr && n >= 0, err, "bad mode list");
  cudaSetDevice(c->device);
  return clpp_dev_perturb_solve_list(c, k_indices, n, err);
}

int clpp_perturb_get_k(const clpp_ctx* c, double* k) {
  if (!c || !c->has_pgrids) return CLPP_FAILURE;
  std::copy(c->k.begin(), c->k.end(), k);
  return CLPP_SUCCESS;
}
int clpp_perturb_get_tau(const clpp_ctx* c, double* tau) {
  if (!c || !c->has_pgrids) return CLPP_FAILURE;
  std::copy(c->tau.begin(), c->tau.end(), tau);
  return CLPP_SUCCESS;
}
int clpp_perturb_get_kstat(const clpp_ctx* c, clpp_kstat* out) {
  if (!c || c->kstat.empty()) return CLPP_FAILURE;
  std::copy(c->kstat.begin(), c->kstat.end(), out);
  return CLPP_SUCCESS;
}

int clpp_perturb_get_sources(clpp_ctx* c, double* out, char* err) {
  CLPP_CHECK(c && c->dev && c->has_sources, err, "no device-resident sources");
  cudaSetDevice(c->device);
  const int nk = c->pinfo.k_size, nt = c->pinfo.tau_size, ntp = c->pinfo.tp_size;
  std::vector<double> tmp((size_t)ntp * nk * nt);
  CLPP_CUDA(cudaMemcpyAsync(tmp.data(), c->dev->sources, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost,
                            c->dev->stream), err);
  CLPP_CUDA(cudaStreamSynchronize(c->dev->stream), err);
  // device [tp][k][tau] -> reference [tp][tau][k]
  for (int tp = 0; tp < ntp; tp++)
    for (int ik = 0; ik < nk; ik++) {
      const double* src = &tmp[((size_t)tp * nk + ik) * nt];
      double* dst = out + (size_t)tp * nt * nk + ik;
      for (int it = 0; it < nt; it++) dst[(size_t)it * nk] = src[it];
    }
  return CLPP_SUCCESS;
}

int clpp_perturb_set_sources(clpp_ctx* c, const clpp_perturb_info* info, const double* k, const double* tau,
                             const double* sources, char* err) {
  CLPP_CHECK(c && info && k && tau && sources, err, "null argument");
  CLPP_CHECK(c->dev, err, "this context has no CUDA device: the B200 path has no CPU fallback");
  cudaSetDevice(c->device);
  c->pinfo = *info;
  c->k.assign(k, k + info->k_size);
  c->tau.assign(tau, tau + info->tau_size);
  c->has_pgrids = true;
  c->nl_dev_valid = false;
  const int nk = info->k_size, nt = info->tau_size, ntp = info->tp_size;
  std::vector<double> tmp((size_t)ntp * nk * nt);
  for (int tp = 0; tp < ntp; tp++)
    for (int it = 0; it < nt; it++) {
      const double* src = sources + ((size_t)tp * nt + it) * nk;
      for (int ik = 0; ik < nk; ik++) tmp[((size_t)tp * nk + ik) * nt + it] = src[ik];
    }
  clpp_ctx::Dev* d = c->dev;
  if (upload(d, &d->k, c->k.data(), c->k.size(), d->stream, err)) return CLPP_FAILURE;
  if (upload(d, &d->tau, c->tau.data(), c->tau.size(), d->stream, err)) return CLPP_FAILURE;
  if (upload(d, &d->sources, tmp.data(), tmp.size(), d->stream, err)) return CLPP_FAILURE;
  d->sources_count = tmp.size();
  CLPP_CUDA(cudaStreamSynchronize(d->stream), err);
  c->has_sources = true;
  return CLPP_SUCCESS;
}

int clpp_perturb_device_sources(clpp_ctx* c, void** dptr, long* count, char* err) {
  CLPP_CHECK(c && c->dev && c->dev->sources, err, "no device-resident sources");
  *dptr = c->dev->sources;
  *count = (long)c->dev->sources_count;
  return CLPP_SUCCESS;
}

// ---- stage 2 ---------------------------------------------------------------------------------
int clpp_transfer_grids(clpp_ctx* c, const clpp_transfer_desc* desc, clpp_transfer_info* info, char* err) {
  CLPP_CHECK(c && desc, err, "null argument");
  c->td = *desc;
  c->has_transfer = false;
  if (clpp_host_transfer_grids(c, err)) return CLPP_FAILURE;
  if (info) *info = c->tinfo;
  return CLPP_SUCCESS;
}

int clpp_transfer_compute(clpp_ctx* c, const double* nl_corr_density, int q_begin, int q_end, char* err) {
  CLPP_CHECK(c && c->has_tgrids, err, "clpp_transfer_grids must be called first");
  CLPP_CHECK(c->dev, err, "this context has no CUDA device: the B200 path has no CPU fallback");
  CLPP_CHECK(c->has_sources, err, "no sources: run clpp_perturb_solve or clpp_perturb_set_sources first");
  CLPP_CHECK(0 <= q_begin && q_begin <= q_end && q_end <= c->tinfo.q_size, err, "bad q range [%d,%d)", q_begin, q_end);
  cudaSetDevice(c->device);
  return clpp_dev_transfer_compute(c, nl_corr_density, q_begin, q_end, err);
}

int clpp_transfer_get_l(const clpp_ctx* c, int* l, int* l_size_tt) {
  if (!c || !c->has_tgrids) return CLPP_FAILURE;
  if (l) std::copy(c->l.begin(), c->l.end(), l);
  if (l_size_tt) std::copy(c->l_size_tt.begin(), c->l_size_tt.end(), l_size_tt);
  return CLPP_SUCCESS;
}
int clpp_transfer_get_q(const clpp_ctx* c, double* q, double* k) {
  if (!c || !c->has_tgrids) return CLPP_FAILURE;
  if (q) std::copy(c->q.begin(), c->q.end(), q);
  if (k) std::copy(c->kq.begin(), c->kq.end(), k);
  return CLPP_SUCCESS;
}

int clpp_transfer_get_transfer(clpp_ctx* c, double* out, char* err) {
  CLPP_CHECK(c && c->dev && c->has_transfer, err, "no device-resident transfer functions");
  cudaSetDevice(c->device);
  CLPP_CUDA(cudaMemcpyAsync(out, c->dev->transfer, c->dev->transfer_count * sizeof(double), cudaMemcpyDeviceToHost,
                            c->dev->stream), err);
  CLPP_CUDA(cudaStreamSynchronize(c->dev->stream), err);
  return CLPP_SUCCESS;
}

int clpp_transfer_set_transfer(clpp_ctx* c, const double* transfer, char* err) {
  CLPP_CHECK(c && c->dev && c->has_tgrids && transfer, err, "transfer grids / device missing");
  cudaSetDevice(c->device);
  clpp_ctx::Dev* d = c->dev;
  const size_t n = (size_t)c->tinfo.tt_size * c->tinfo.l_size * c->tinfo.q_size;
  if (upload(d, &d->transfer, transfer, n, d->stream, err)) return CLPP_FAILURE;
  d->transfer_count = n;
  if (upload(d, &d->kq, c->kq.data(), c->kq.size(), d->stream, err)) return CLPP_FAILURE;
  CLPP_CUDA(cudaStreamSynchronize(d->stream), err);
  c->has_transfer = true;
  return CLPP_SUCCESS;
}

int clpp_transfer_device_transfer(clpp_ctx* c, void** dptr, long* count, char* err) {
  CLPP_CHECK(c && c->dev && c->dev->transfer, err, "no device-resident transfer functions");
  *dptr = c->dev->transfer;
  *count = (long)c->dev->transfer_count;
  return CLPP_SUCCESS;
}

int clpp_transfer_get_bessel(clpp_ctx* c, double* x, double* phi, double* dphi, double* chi, char* err) {
  CLPP_CHECK(c && c->dev && c->dev->bessel_phi, err, "no Bessel table on the device (run clpp_transfer_compute)");
  cudaSetDevice(c->device);
  return clpp_dev_get_bessel(c, x, phi, dphi, chi, err);
}

// ---- stage 3 ---------------------------------------------------------------------------------
int clpp_spectra_compute_range(clpp_ctx* c, const double* primordial_pk, int q_begin, int q_end,
                               clpp_spectra_info* info, double* cl_out, char* err) {
  CLPP_CHECK(c && primordial_pk && cl_out, err, "null argument");
  CLPP_CHECK(c->dev, err, "this context has no CUDA device: the B200 path has no CPU fallback");
  CLPP_CHECK(c->has_transfer, err, "no transfer functions: run clpp_transfer_compute first");
  CLPP_CHECK(0 <= q_begin && q_begin <= q_end && q_end <= c->tinfo.q_size, err, "bad q range [%d,%d)", q_begin, q_end);
  cudaSetDevice(c->device);
  return clpp_dev_spectra(c, primordial_pk, q_begin, q_end, info, cl_out, err);
}

int clpp_pk_linear(clpp_ctx* c, const double* primordial_pk, int index_tau, int cb, double* pk_out, char* err) {
  CLPP_CHECK(c && primordial_pk && pk_out, err, "null argument");
  CLPP_CHECK(c->dev, err, "this context has no CUDA device: the B200 path has no CPU fallback");
  CLPP_CHECK(c->has_sources, err, "no sources: run clpp_perturb_solve first");
  cudaSetDevice(c->device);
  return clpp_dev_pk_linear(c, primordial_pk, index_tau, cb, pk_out, err);
}

int clpp_perturb_sources_at_tau(clpp_ctx* c, int index_tp, double tau, double* psource, char* err) {
  CLPP_CHECK(c && psource, err, "null argument");
  CLPP_CHECK(c->dev, err, "this context has no CUDA device: the B200 path has no CPU fallback");
  CLPP_CHECK(c->has_sources, err, "no sources: run clpp_perturb_solve first");
  cudaSetDevice(c->device);
  return clpp_dev_sources_at_tau(c, index_tp, tau, psource, err);
}

int clpp_nonlinear_halofit(clpp_ctx* c, const clpp_halofit_desc* desc, const double* primordial_pk, double* nl_corr_out,
                           int* index_tau_min_nl, char* err) {
  CLPP_CHECK(c && desc && primordial_pk, err, "null argument");
  CLPP_CHECK(c->dev, err, "this context has no CUDA device: the B200 path has no CPU fallback");
  CLPP_CHECK(c->has_sources, err, "no sources: run clpp_perturb_solve first");
  cudaSetDevice(c->device);
  return clpp_dev_halofit(c, desc, primordial_pk, nl_corr_out, index_tau_min_nl, err);
}

int clpp_lensing_compute(clpp_ctx* c, const clpp_lensing_desc* desc, clpp_lensing_info* info, double* l_out,
                         double* cl_lens_out, char* err) {
  CLPP_CHECK(c && desc, err, "null argument");
  CLPP_CHECK(c->dev, err, "this context has no CUDA device: the B200 path has no CPU fallback");
  cudaSetDevice(c->device);
  return clpp_dev_lensing(c, desc, info, l_out, cl_lens_out, err);
}

int clpp_lensing_cl_at_l(const clpp_ctx* c, int l, double* cl_lensed, char* err) {
  CLPP_CHECK(c && cl_lensed, err, "null argument");
  return clpp_host_lensing_cl_at_l(c, l, cl_lensed, err);
}

int clpp_spectra_cl_at_l(const clpp_ctx* c, double l, double* cl_tot, char* err) {
  CLPP_CHECK(c && cl_tot, err, "null argument");
  return clpp_host_cl_at_l(c, l, cl_tot, err);
}

int clpp_spectra_cl_output(const clpp_ctx* c, int lmax, double* out, char* err) {
  CLPP_CHECK(c && out, err, "null argument");
  CLPP_CHECK(c->has_cl, err, "Error: Cls have not been computed! lmax = %d", lmax);
  const int l_max_tot = (int)c->clt.x[c->clt.n_lines - 1];
  CLPP_CHECK(lmax >= 0 && lmax <= l_max_tot, err, "Error: lmax = %d is outside the allowed range [0, %d]", lmax, l_max_tot);
  const int ct = c->clt.n_cols;
  for (int i = 0; i < 2 * ct && i < (lmax + 1) * ct; i++) out[i] = 0.;
  for (int l = 2; l <= lmax; l++)
    if (clpp_host_cl_at_l(c, (double)l, out + (size_t)l * ct, err)) return CLPP_FAILURE;
  return CLPP_SUCCESS;
}

int clpp_spectra_compute(clpp_ctx* c, const double* primordial_pk, clpp_spectra_info* info, double* cl_out,
                         char* err) {
  CLPP_CHECK(c, err, "null ctx");
  return clpp_spectra_compute_range(c, primordial_pk, 0, c->tinfo.q_size, info, cl_out, err);
}

}  // extern "C"
