// Template instantiations for the logistic vector field (one translation unit per field so that the
// build parallelises).
#include "inst_common.cuh"
namespace pnde {
PNDE_DEFINE_OPS(ops_logistic, VfLogistic)
}  // namespace pnde
