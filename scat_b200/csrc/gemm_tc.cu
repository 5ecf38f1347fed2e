// placeholder until the tcgen05 kernel lands
#include "kernels.h"
namespace scat {
bool gemm_tc_supported(const GemmArgs&) { return false; }
int launch_gemm_tc(const GemmArgs&, int, cudaStream_t) {
    set_last_error("tcgen05 GEMM not built");
    return kErrUnsupported;
}
}  // namespace scat
