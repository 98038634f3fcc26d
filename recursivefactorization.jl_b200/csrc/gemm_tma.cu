// gemm_tma.cu -- placeholder until the TMA-fed DMMA kernel lands (see gemm.cu for the generic path).
#include "rfb_internal.h"

int rfb_launch_gemm_f64_tma(rfb_ctx *ctx, double *C, const double *A, const double *B, int64_t m, int64_t n,
                            int64_t k, int64_t lda, bool *handled) {
    (void)ctx; (void)C; (void)A; (void)B; (void)m; (void)n; (void)k; (void)lda;
    *handled = false;
    return RFB_OK;
}
