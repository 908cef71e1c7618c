// Run-time compiled models: user vector fields given as CUDA C++ source.
#pragma once
#include <string>

#include "model_ops.cuh"

namespace pnde {

// f_body / jac_body: statement lists assigning du[i] (from u[], p[]; generic in the scalar type T, may use
// + - * / exp log sin cos sqrt) and J[i][j] (doubles).  jac_body may be null for EK0.  Returns nullptr and
// fills err on failure.  ieks: the filter kernels carry the IEKS linearisation policy (ieks_kernel.cuh).
// adaptive: 0 / 1 compile only the fixed-step / adaptive filter kernel, -1 both.  quirk_check: -DPNDE_QUIRK_CHECK
// (PNDE_FLAG_REFERENCE_QUIRKS (b), filter_kernel.cuh).  lane_groups: dense EK1 with d (q+1) >= 10 and even d gets the
// lane-group filter and smoother (wide_filter.cuh, wide_smoother.cuh) like the catalogue models.  The returned object is owned by the caller (rtc_destroy).
const ModelOps* rtc_build(int alg, int q, bool mvdyn, int d, int np, const char* f_body, const char* jac_body,
                          std::string& err, bool ieks = false, int adaptive = -1, bool quirk_check = false,
                          bool lane_groups = true);
void rtc_destroy(const ModelOps* ops);
bool rtc_rolled(int alg, int d, int q);  // would this user model be built with rolled loops (general-(d, q) fallback)?
// compile-only validation of the source (needs libnvrtc, not a GPU)
bool rtc_check(int alg, int q, bool mvdyn, int d, int np, const char* f_body, const char* jac_body, std::string& err,
               bool ieks = false);

}  // namespace pnde
