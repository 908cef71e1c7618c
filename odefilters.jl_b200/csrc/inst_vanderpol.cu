// Template instantiations for the vanderpol vector field (one translation unit per field so that the
// build parallelises).
#include "inst_common.cuh"
namespace pnde {
PNDE_DEFINE_OPS(ops_vanderpol, VfVanDerPol)
}  // namespace pnde
