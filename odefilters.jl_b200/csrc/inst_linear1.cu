// Template instantiations for the linear1 vector field (one translation unit per field so that the
// build parallelises).
#include "inst_common.cuh"
namespace pnde {
PNDE_DEFINE_OPS(ops_linear1, VfLinear1)
}  // namespace pnde
