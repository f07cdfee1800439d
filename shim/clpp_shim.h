// Shared helpers of the drop-in constructors (shim/*.cpp): the three hot-path modules of CLASS++ re-implemented as thin
// C++ classes over the C ABI of libclpp.so (include/clpp.h).  They keep the reference's OWN headers
// (source/perturbations_module.h, transfer_module.h, spectra_module.h, included from the reference tree where it lies),
// so every other module of the reference, Cosmology (source/cosmology.cpp:30-79) and classy (classy.pyx) compile and link
// against them unchanged.  One clpp_ctx (device-resident state of one cosmology) is created by PerturbationsModule and
// shared with the TransferModule / SpectraModule built on top of it.
#ifndef CLPP_SHIM_H
#define CLPP_SHIM_H

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "clpp.h"

class PerturbationsModule;

namespace clpp_shim {
// context registry keyed by the PerturbationsModule that owns it
void register_ctx(const PerturbationsModule* p, clpp_ctx* c);
clpp_ctx* ctx_of(const PerturbationsModule* p);  // nullptr when absent
void release_ctx(const PerturbationsModule* p);
// CUDA device of this process: CLPP_DEVICE, else LOCAL_RANK (torchrun: one process per GPU), else 0
inline int device() {
  const char* e = getenv("CLPP_DEVICE");
  if (!e) e = getenv("LOCAL_RANK");
  return e ? atoi(e) : 0;
}
}  // namespace clpp_shim

// run a clpp call; on failure copy its message into the module's ErrorMsg and return _FAILURE_
#define CLPP_SHIM_CALL(call, errbuf)                                     \
  do {                                                                   \
    char clpp_err_[CLPP_ERRLEN];                                         \
    clpp_err_[0] = 0;                                                    \
    if ((call) != CLPP_SUCCESS) {                                        \
      snprintf(errbuf, sizeof(ErrorMsg), "%s(L:%d) : %s", __func__, __LINE__, clpp_err_); \
      return _FAILURE_;                                                  \
    }                                                                    \
  } while (0)
#define CLPP_SHIM_TEST(cond, errbuf, ...)                                \
  do {                                                                   \
    if (cond) {                                                          \
      char msg_[1024];                                                   \
      snprintf(msg_, sizeof(msg_), __VA_ARGS__);                         \
      snprintf(errbuf, sizeof(ErrorMsg), "%s(L:%d) : %s", __func__, __LINE__, msg_); \
      return _FAILURE_;                                                  \
    }                                                                    \
  } while (0)

#endif
