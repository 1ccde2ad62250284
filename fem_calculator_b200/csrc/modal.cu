// Modal solver placeholder — replaced by the LOBPCG / shift-invert implementation.
#include "common.cuh"
namespace femb {
int run_modal(femb_handle* h, const femb_eig_opts&, double*, double*, int32_t*, femb_stats*) {
  return fail(h, FEMB_ERR_ARG, "modal solver not built yet");
}
}  // namespace femb
