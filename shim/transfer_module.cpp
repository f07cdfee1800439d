// Drop-in TransferModule: the reference's class (source/transfer_module.h:7-54, unchanged header) over the C ABI of
// libclpp.so.  Replaces source/transfer_module.cpp: l / q / k(q) grids on the host (bit-exact), k-spline of the sources,
// flat Bessel table and the line-of-sight integrals on the GPU, Delta_l(q) handed back in the reference layout
// transfer_[index_md][((index_ic*tt_size+index_tt)*l_size+index_l)*q_size+index_q].
#include <stdexcept>
#include <vector>

#include "background_module.h"
#include "nonlinear_module.h"
#include "perturbations_module.h"
#include "thermodynamics_module.h"
#include "transfer_module.h"

#include "clpp_shim.h"

TransferModule::TransferModule(InputModulePtr input_module, BackgroundModulePtr background_module,
                               ThermodynamicsModulePtr thermodynamics_module, PerturbationsModulePtr perturbations_module,
                               NonlinearModulePtr nonlinear_module)
    : BaseModule(std::move(input_module)),
      background_module_(std::move(background_module)),
      thermodynamics_module_(std::move(thermodynamics_module)),
      perturbations_module_(std::move(perturbations_module)),
      nonlinear_module_(std::move(nonlinear_module)) {
  if (transfer_init() != _SUCCESS_) throw std::runtime_error(error_message_);
}

TransferModule::~TransferModule() { transfer_free(); }

int TransferModule::transfer_init() {
  nz_size_ = 0;
  nz_evo_size_ = 0;
  has_cls_ = ppt->has_cls;
  md_size_ = 0;
  if (ppt->has_cls == _FALSE_) {
    if (ptr->transfer_verbose > 0) printf("No harmonic space transfer functions to compute. Transfer module skipped.\n");
    return _SUCCESS_;
  }
  if (ptr->transfer_verbose > 0) printf("Computing transfers (B200 path)\n");
  md_size_ = perturbations_module_->md_size_;
  clpp_ctx* ctx = clpp_shim::ctx_of(perturbations_module_.get());
  CLPP_SHIM_TEST(ctx == nullptr, error_message_, "the PerturbationsModule holds no device context");
  CLPP_SHIM_TEST(ppt->has_nc_density == _TRUE_ || ppt->has_nc_rsd == _TRUE_ || ppt->has_nc_lens == _TRUE_ || ppt->has_nc_gr == _TRUE_ ||
                     ppt->has_cl_lensing_potential == _TRUE_,
                 error_message_, "number count / galaxy lensing transfer types are not supported by the B200 path");

  clpp_transfer_desc d;
  memset(&d, 0, sizeof(d));
  d.has_cl_cmb_temperature = ppt->has_cl_cmb_temperature; d.has_cl_cmb_polarization = ppt->has_cl_cmb_polarization;
  d.has_cl_cmb_lensing_potential = ppt->has_cl_cmb_lensing_potential; d.l_scalar_max = ppt->l_scalar_max;
#define PR(x) d.x = ppr->x
  PR(l_logstep); PR(l_linstep); PR(hyper_x_min); PR(hyper_sampling_flat); PR(hyper_phi_min_abs);
  PR(q_linstep); PR(q_logstep_spline); PR(q_logstep_open);
  PR(transfer_neglect_delta_k_S_t0); PR(transfer_neglect_delta_k_S_t1); PR(transfer_neglect_delta_k_S_t2);
  PR(transfer_neglect_delta_k_S_e); PR(transfer_neglect_late_source); PR(l_switch_limber);
#undef PR
  d.lcmb_rescale = ptr->lcmb_rescale; d.lcmb_tilt = ptr->lcmb_tilt; d.lcmb_pivot = ptr->lcmb_pivot;
  clpp_transfer_info ti;
  CLPP_SHIM_CALL(clpp_transfer_grids(ctx, &d, &ti, clpp_err_), error_message_);

  // transfer types (transfer_indices_of_transfers :402-470, scalar branch) and grids
  index_tt_t0_ = ti.index_tt_t0; index_tt_t1_ = ti.index_tt_t1; index_tt_t2_ = ti.index_tt_t2; index_tt_e_ = ti.index_tt_e;
  index_tt_lcmb_ = ti.index_tt_lcmb;
  tt_size_ = (int*)malloc(sizeof(int));
  tt_size_[0] = ti.tt_size;
  l_size_max_ = ti.l_size_max;
  l_size_ = (int*)malloc(sizeof(int));
  l_size_[0] = ti.l_size;
  l_ = (int*)malloc(ti.l_size_max * sizeof(int));
  l_size_tt_ = (int**)malloc(sizeof(int*));
  l_size_tt_[0] = (int*)malloc(ti.tt_size * sizeof(int));
  clpp_transfer_get_l(ctx, l_, l_size_tt_[0]);
  q_size_ = ti.q_size;
  q_ = (double*)malloc(ti.q_size * sizeof(double));
  k_ = (double**)malloc(sizeof(double*));
  k_[0] = (double*)malloc(ti.q_size * sizeof(double));
  clpp_transfer_get_q(ctx, q_, k_[0]);
  index_q_flat_approximation_ = ti.q_size;  // flat space: no q uses the flat rescaling approximation

  // non-linear corrections of phi+psi (transfer_perturbation_copy_sources_and_nl_corrections :542-601)
  const double* nl_corr = nullptr;
  if (pnl->method != nl_none && ppt->has_cl_cmb_lensing_potential == _TRUE_) {
    CLPP_SHIM_TEST(nonlinear_module_->has_pk_m_ == _FALSE_, error_message_,
                   "non-linear corrections requested but the NonlinearModule holds no total matter spectrum");
    nl_corr = nonlinear_module_->nl_corr_density_[nonlinear_module_->index_pk_m_];
  }
  CLPP_SHIM_CALL(clpp_transfer_compute(ctx, nl_corr, 0, ti.q_size, clpp_err_), error_message_);
  transfer_ = (double**)malloc(sizeof(double*));
  transfer_[0] = (double*)malloc((size_t)ti.tt_size * ti.l_size * ti.q_size * sizeof(double));
  CLPP_SHIM_CALL(clpp_transfer_get_transfer(ctx, transfer_[0], clpp_err_), error_message_);
  return _SUCCESS_;
}

int TransferModule::transfer_free() {
  if (has_cls_ == _TRUE_ && md_size_ > 0) {
    free(l_size_tt_[0]); free(transfer_[0]); free(k_[0]);
    free(tt_size_); free(l_size_tt_); free(l_size_); free(l_); free(q_); free(k_); free(transfer_);
    md_size_ = 0;
  }
  return _SUCCESS_;
}
