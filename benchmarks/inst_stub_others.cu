// Experiment builds only (make exp): every field except fhn_readme resolves to "no kernel".
#include "model_ops.cuh"
namespace pnde {
const ModelOps* ops_fhn_lib(int, int, bool) { return nullptr; }
const ModelOps* ops_lotka_volterra(int, int, bool) { return nullptr; }
const ModelOps* ops_vanderpol(int, int, bool) { return nullptr; }
const ModelOps* ops_linear2(int, int, bool) { return nullptr; }
const ModelOps* ops_logistic(int, int, bool) { return nullptr; }
const ModelOps* ops_linear1(int, int, bool) { return nullptr; }
}  // namespace pnde
