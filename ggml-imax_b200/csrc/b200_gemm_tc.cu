// placeholder until the tcgen05 GEMM lands
#include "b200_internal.cuh"
bool b200_gemm_available(void) { return false; }
int b200_launch_gemm(b200_ctx *ctx, const b200_gemm_params &) {
    b200_set_error(ctx, "tcgen05 GEMM not built");
    return B200_ERR_UNSUPPORTED;
}
