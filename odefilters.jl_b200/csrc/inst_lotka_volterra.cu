// Template instantiations for the lotka_volterra vector field (one translation unit per field so that the
// build parallelises).
#include "inst_common.cuh"
namespace pnde {
PNDE_DEFINE_OPS(ops_lotka_volterra, VfLotkaVolterra)
}  // namespace pnde
