// placeholder until the tcgen05 kernel lands
#include "qce_common.cuh"
namespace qce {
bool tc_supported(const qce_model*, int) { return false; }
qce_status tc_pack_params(qce_model*, cudaStream_t) { return QCE_OK; }
void tc_free(qce_model*) {}
qce_status launch_dense_tc(const qce_model*, cudaStream_t, const double*, int64_t, int, int, double, double*, double*,
                           const void*, int, double*) {
    set_error("tensor-core kernel not built");
    return QCE_ERR_UNSUPPORTED;
}
}  // namespace qce
