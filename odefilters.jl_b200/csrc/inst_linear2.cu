// Template instantiations for the linear2 vector field (one translation unit per field so that the
// build parallelises).
#include "inst_common.cuh"
namespace pnde {
PNDE_DEFINE_OPS(ops_linear2, VfLinear2)
}  // namespace pnde
