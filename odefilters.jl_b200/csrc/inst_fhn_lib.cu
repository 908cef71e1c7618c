// Template instantiations for the fhn_lib vector field (one translation unit per field so that the
// build parallelises).
#include "inst_common.cuh"
namespace pnde {
PNDE_DEFINE_OPS(ops_fhn_lib, VfFhnLib)
}  // namespace pnde
