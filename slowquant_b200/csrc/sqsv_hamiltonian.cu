// sigma-build and reduced density matrices (placeholder; filled in below)
#include "sqsv_internal.h"

extern "C" int sq_sigma(sq_space* sp, double e_core, const double* h_act_host, const double* g_act_host,
                        const double* in_dev, double* out_dev, void* stream) {
  sq_set_error("sq_sigma: not built yet");
  return SQ_ERR_UNSUPPORTED;
}
extern "C" int sq_rdm12(sq_space* sp, const double* bra_dev, const double* ket_dev, double* rdm1_host,
                        double* rdm2_host, void* stream) {
  sq_set_error("sq_rdm12: not built yet");
  return SQ_ERR_UNSUPPORTED;
}
