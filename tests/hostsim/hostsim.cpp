// TEST INFRASTRUCTURE (developer harness), not product code and never linked into libclpp.so.
//
// Executes the device "lane program" of classpp_public_b200/csrc/lane.cuh (one thread = one k mode) on the CPU, one mode
// after the other, so that its logic can be debugged and checked against the golden vectors in a container without a
// GPU.  The product path has no CPU fallback: libclpp.so only ever launches the CUDA kernels.
#define CLPP_HOST_SIM 1
#include <cstdlib>
#include <vector>

#include "clpp_internal.h"
#include "lane.cuh"

extern "C" int clpp_pt_fill_common(const clpp_ctx* c, PtParams* P, char* err);  // libclpp.so (perturb.cu)

extern "C" int hostsim_solve(void* ctx, const int* k_list, int n, double* sources, clpp_kstat* kstat, char* err) {
  clpp_ctx* c = (clpp_ctx*)ctx;
  PtParams P;
  if (clpp_pt_fill_common(c, &P, err)) return CLPP_FAILURE;
  std::vector<double> i2l1(P.n_i2l1);
  for (int l = 0; l < P.n_i2l1; l++) i2l1[l] = 1.0 / (2.0 * l + 1.0);
  P.i2l1 = i2l1.data();
  PtCosmo Q;
  memset(&Q, 0, sizeof(Q));
  Q.bg_tau = c->bgt.x.data(); Q.bg_y = c->bgt.y.data(); Q.bg_dd = c->bgt.ddy.data();
  Q.th_z = c->tht.x.data(); Q.th_y = c->tht.y.data(); Q.th_dd = c->tht.ddy.data();
  Q.ncdm_q = c->ncdm_q.data(); Q.ncdm_w = c->ncdm_w.data(); Q.ncdm_dlnf0 = c->ncdm_dlnf0.data();
  Q.k = c->k.data(); Q.tau = c->tau.data(); Q.sources = sources; Q.kstat = kstat;
  Q.bt_size = c->bg.bt_size; Q.tt_size = c->th.tt_size; Q.k_size = c->pinfo.k_size; Q.tau_size = c->pinfo.tau_size;
  Q.th_linear_below_z = -1.;
  if (c->th.reio_parametrization == CLPP_REIO_HALF_TANH) Q.th_linear_below_z = 2 * c->th.z_reionization;
  if (c->th.reio_parametrization == CLPP_REIO_INTER) Q.th_linear_below_z = 50.;
  Q.n_e = c->th.n_e; Q.YHe = c->th.YHe; Q.T_cmb = c->bg.T_cmb; Q.tau_free_streaming = c->th.tau_free_streaming;
  Q.a_today = c->bg.a_today;
  for (int s = 0; s < P.N_ncdm; s++) { Q.ncdm_M[s] = c->ncdm_M[s]; Q.ncdm_factor[s] = c->ncdm_factor[s]; }
  if (getenv("HOSTSIM_DENSE")) P.ln_structured = 0; else if (getenv("HOSTSIM_STRUCTURED")) P.ln_structured = 1;
  std::vector<double> mem((size_t)P.ln_words + 4096);
  if (getenv("HOSTSIM_VERBOSE"))
    fprintf(stderr, "[hostsim] neq_max %d nh_max %d words %d\n", P.neq_max, P.nh_max, P.ln_words);
  for (int i = 0; i < n; i++) {
    std::fill(mem.begin(), mem.end(), 0.);
    ln_mode(P, mem.data(), &Q, k_list[i]);
  }
  return CLPP_SUCCESS;
}
