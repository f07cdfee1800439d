// Drop-in SpectraModule: the reference's class (source/spectra_module.h:11-81, unchanged header) over the C ABI of
// libclpp.so.  Replaces source/spectra_module.cpp: the k-quadrature C_l = 4 pi int dk/k P(k) Delta^X Delta^Y runs on the
// GPU from the device-resident transfer functions; the table cl_[index_md][(index_l*ic_ic_size+0)*ct_size+index_ct] and
// the accessors (spectra_cl_at_l, cl_output*) keep the reference's layout and conventions.
#include <stdexcept>
#include <vector>

#include "exceptions.h"
#include "nonlinear_module.h"
#include "perturbations_module.h"
#include "primordial_module.h"
#include "spectra_module.h"
#include "transfer_module.h"

#include "clpp_shim.h"

SpectraModule::SpectraModule(InputModulePtr input_module, PerturbationsModulePtr perturbations_module,
                             PrimordialModulePtr primordial_module, NonlinearModulePtr nonlinear_module,
                             TransferModulePtr transfer_module)
    : BaseModule(std::move(input_module)),
      perturbations_module_(std::move(perturbations_module)),
      primordial_module_(std::move(primordial_module)),
      nonlinear_module_(std::move(nonlinear_module)),
      transfer_module_(std::move(transfer_module)) {
  if (spectra_init() != _SUCCESS_) throw std::runtime_error(error_message_);
}

SpectraModule::~SpectraModule() { spectra_free(); }

int SpectraModule::spectra_init() {
  md_size_ = 0;
  ct_size_ = 0;
  d_size_ = 0;
  has_tt_ = has_ee_ = has_te_ = has_bb_ = has_pp_ = has_tp_ = has_ep_ = _FALSE_;
  has_dd_ = has_td_ = has_pd_ = has_ll_ = has_tl_ = has_dl_ = _FALSE_;
  if (ppt->has_cls == _FALSE_) {
    if (psp->spectra_verbose > 0) printf("No spectra requested. Spectra module skipped.\n");
    return _SUCCESS_;
  }
  if (psp->spectra_verbose > 0) printf("Computing unlensed harmonic spectra (B200 path)\n");
  clpp_ctx* ctx = clpp_shim::ctx_of(perturbations_module_.get());
  CLPP_SHIM_TEST(ctx == nullptr, error_message_, "the PerturbationsModule holds no device context");

  // modes, pairs of initial conditions (spectra_indices :527-560): one mode, one adiabatic pair here
  md_size_ = perturbations_module_->md_size_;
  index_md_scalars_ = perturbations_module_->index_md_scalars_;
  const int md = index_md_scalars_;
  ic_size_ = (int*)malloc(sizeof(int) * md_size_);
  ic_ic_size_ = (int*)malloc(sizeof(int) * md_size_);
  is_non_zero_ = (short**)malloc(sizeof(short*) * md_size_);
  ic_size_[md] = primordial_module_->ic_size_[md];
  ic_ic_size_[md] = primordial_module_->ic_ic_size_[md];
  is_non_zero_[md] = (short*)malloc(sizeof(short) * ic_ic_size_[md]);
  for (int i = 0; i < ic_ic_size_[md]; i++) is_non_zero_[md][i] = primordial_module_->is_non_zero_[md][i];
  CLPP_SHIM_TEST(ic_size_[md] != 1, error_message_, "the B200 path computes adiabatic initial conditions only");

  // primordial spectrum on the k(q) grid (spectra_compute_cl :996)
  const int nq = transfer_module_->q_size_;
  std::vector<double> pk(nq);
  for (int iq = 0; iq < nq; iq++) {
    double v[4];
    class_call(primordial_module_->primordial_spectrum_at_k(md, linear, transfer_module_->k_[md][iq], v),
               primordial_module_->error_message_, error_message_);
    pk[iq] = v[0];
  }
  clpp_spectra_info si;
  const int nl = transfer_module_->l_size_[md];
  std::vector<double> table((size_t)nl * 16);
  CLPP_SHIM_CALL(clpp_spectra_compute(ctx, pk.data(), &si, table.data(), clpp_err_), error_message_);
  CLPP_SHIM_TEST(si.l_size != nl, error_message_, "multipole grids of the transfer and spectra stages differ (%d vs %d)", si.l_size, nl);

  // C_l types (spectra_indices :560-640) and their l_max (:712-795)
  ct_size_ = si.ct_size;
  has_tt_ = si.index_ct_tt >= 0; has_ee_ = si.index_ct_ee >= 0; has_te_ = si.index_ct_te >= 0; has_bb_ = si.index_ct_bb >= 0;
  has_pp_ = si.index_ct_pp >= 0; has_tp_ = si.index_ct_tp >= 0; has_ep_ = si.index_ct_ep >= 0;
  index_ct_tt_ = si.index_ct_tt; index_ct_ee_ = si.index_ct_ee; index_ct_te_ = si.index_ct_te; index_ct_bb_ = si.index_ct_bb;
  index_ct_pp_ = si.index_ct_pp; index_ct_tp_ = si.index_ct_tp; index_ct_ep_ = si.index_ct_ep;
  l_max_ = (int*)malloc(sizeof(int) * md_size_);
  l_max_ct_ = (int**)malloc(sizeof(int*) * md_size_);
  l_max_ct_[md] = (int*)calloc(ct_size_, sizeof(int));
  const int idx[6] = {index_ct_tt_, index_ct_ee_, index_ct_te_, index_ct_pp_, index_ct_tp_, index_ct_ep_};
  for (int i = 0; i < 6; i++)
    if (idx[i] >= 0) l_max_ct_[md][idx[i]] = ppt->l_scalar_max;  // BB of scalar modes stays 0
  l_max_[md] = 0;
  for (int ct = 0; ct < ct_size_; ct++) l_max_[md] = MAX(l_max_[md], l_max_ct_[md][ct]);
  l_max_tot_ = l_max_[md];

  // table and its spline along l (spectra_cls :804-939)
  l_size_ = (int*)malloc(sizeof(int) * md_size_);
  l_size_[md] = nl;
  l_size_max_ = nl;
  l_ = (double*)malloc(sizeof(double) * nl);
  for (int il = 0; il < nl; il++) l_[il] = transfer_module_->l_[il];
  cl_ = (double**)malloc(sizeof(double*) * md_size_);
  ddcl_ = (double**)malloc(sizeof(double*) * md_size_);
  cl_[md] = (double*)malloc(sizeof(double) * nl * ct_size_);
  ddcl_[md] = (double*)malloc(sizeof(double) * nl * ct_size_);
  memcpy(cl_[md], table.data(), sizeof(double) * nl * ct_size_);
  class_call(array_spline_table_lines(l_, nl, cl_[md], ic_ic_size_[md] * ct_size_, ddcl_[md], _SPLINE_EST_DERIV_, error_message_),
             error_message_, error_message_);
  return _SUCCESS_;
}

int SpectraModule::spectra_free() {
  if (ppt->has_cls == _FALSE_ || md_size_ == 0) return _SUCCESS_;
  const int md = index_md_scalars_;
  free(l_max_ct_[md]); free(cl_[md]); free(ddcl_[md]);
  free(l_); free(l_size_); free(l_max_ct_); free(l_max_); free(cl_); free(ddcl_);
  free(is_non_zero_[md]); free(is_non_zero_); free(ic_size_); free(ic_ic_size_);
  md_size_ = 0;
  return _SUCCESS_;
}

// C_l at any multipole (spectra_module.cpp:220-427): spline in l through the table, zero above the l_max of each type.
// One mode and one initial condition on this path, so only cl_tot is filled (as the reference does in that case).
int SpectraModule::spectra_cl_at_l(double l, double* cl_tot, double** cl_md, double** cl_md_ic) const {
  (void)cl_md; (void)cl_md_ic;
  const int md = index_md_scalars_;
  if ((int)l <= l_[l_size_[md] - 1]) {
    int last_index;
    class_call(array_interpolate_spline(l_, l_size_[md], cl_[md], ddcl_[md], ct_size_, l, &last_index, cl_tot, ct_size_, error_message_),
               error_message_, error_message_);
    for (int ct = 0; ct < ct_size_; ct++)
      if ((int)l > l_max_ct_[md][ct]) cl_tot[ct] = 0.;
  } else {
    for (int ct = 0; ct < ct_size_; ct++) cl_tot[ct] = 0.;
  }
  return _SUCCESS_;
}

std::map<std::string, int> SpectraModule::cl_output_index_map() const {
  std::map<std::string, int> m;
  if (has_tt_) m["tt"] = index_ct_tt_;
  if (has_ee_) m["ee"] = index_ct_ee_;
  if (has_te_) m["te"] = index_ct_te_;
  if (has_bb_) m["bb"] = index_ct_bb_;
  if (has_pp_) m["pp"] = index_ct_pp_;
  if (has_tp_) m["tp"] = index_ct_tp_;
  if (has_ep_) m["ep"] = index_ct_ep_;
  return m;  // density / lensing-potential number-count types do not exist on this path
}

void SpectraModule::cl_output_no_copy(int lmax, std::vector<double*>& output_pointers) const {
  ThrowRuntimeErrorIf((lmax > l_max_tot_) || (lmax < 0), "Error: lmax = %d is outside the allowed range [0, %d]\n", lmax, l_max_tot_);
  ThrowRuntimeErrorIf((int)output_pointers.size() != ct_size_, "Error: Size of input vector (%d) does not match ct_size = %d\n",
                      (int)output_pointers.size(), ct_size_);
  std::vector<double> row(ct_size_);
  for (int l = 0; l <= lmax; l++) {
    if (l >= 2) {
      const int status = spectra_cl_at_l(l, row.data(), nullptr, nullptr);
      ThrowRuntimeErrorIf(status != _SUCCESS_, "Error in SpectraModule::cl_output: %s", error_message_);
    }
    for (int ct = 0; ct < ct_size_; ct++) output_pointers[ct][l] = (l < 2) ? 0.0 : row[ct];
  }
}

std::map<std::string, std::vector<double>> SpectraModule::cl_output(int lmax) const {
  ThrowRuntimeErrorIf(ppt->has_cls == _FALSE_, "Error: Cls have not been computed! lmax = %d\n", lmax);
  ThrowRuntimeErrorIf((lmax > l_max_tot_) || (lmax < 0), "Error: lmax = %d is outside the allowed range [0, %d]\n", lmax, l_max_tot_);
  std::vector<std::vector<double>> cols(ct_size_, std::vector<double>(lmax + 1, 0.0));
  std::vector<double*> ptrs(ct_size_);
  for (int ct = 0; ct < ct_size_; ct++) ptrs[ct] = cols[ct].data();
  cl_output_no_copy(lmax, ptrs);
  std::map<std::string, std::vector<double>> out;
  for (const auto& kv : cl_output_index_map()) out[kv.first] = std::move(cols[kv.second]);
  return out;
}
