// panel_nopiv.cu -- K1': LU of a tall m x n panel (n <= 64) WITHOUT pivoting, one launch.
//
// Reference: `_generic_lufact!(A, Val(false), ipiv, info)` (src/lu.jl:290-338 with Pivot = false:
// kp = k, no interchange, reciprocal scaling :317-320, rank-1 update :330-334, a zero pivot is
// recorded as NEGATIVE info and the column is left unscaled :321-327) and the bottom levels of
// `reckernel!` (:189-263) below the leaf width.  The operation order per element is the unblocked
// loop's own (multiply by the correctly rounded reciprocal, then one FMA per earlier column in
// column order), so the result is bit-identical to that loop.
//
// B200 design: without a pivot search there is nothing to exchange between CTAs.  Every CTA
// factors the n x n diagonal block REDUNDANTLY (n threads, one row each, kept in registers as the
// same sliding window as K1) and publishes row k of U to shared memory at step k; the CTA's other
// 128 threads each own one row below the diagonal block and eliminate it against that row in the
// same step -- one __syncthreads per column, no grid-wide communication, no cooperative launch.
// The panel is read once and written once (2*s*m*n bytes); the column of L produced at step k is
// stored straight to global memory, coalesced down the column.
#include "rfb_internal.h"

namespace {

constexpr int kRowThreads = 128;     // rows below the diagonal block per CTA

__device__ __forceinline__ double rcp_rn(double v) { return __drcp_rn(v); }
__device__ __forceinline__ float rcp_rn(float v) { return __frcp_rn(v); }

template <typename T, int NB>
struct NoPivShared {
    T uw[NB][NB];      // uw[k][j] = U[k][k + j] (window-relative), zero beyond the panel width
    T rinv[NB];        // 1 / U[k][k], or 1 when the pivot is exactly zero (column stays unscaled)
    int first_zero;    // 1-based column of the first exactly-zero pivot, 0 = none
};

template <typename T, int NB>
__global__ void __launch_bounds__(NB + kRowThreads)
panel_nopiv_kernel(T *__restrict__ A, int m, int n, long long lda, long long *__restrict__ info,
                   long long col_offset) {
    __shared__ NoPivShared<T, NB> sh;
    const int tid = threadIdx.x;
    const bool diag = tid < NB;                                   // owns row `tid` of the diagonal block
    const long long row = diag ? tid : (long long)n + (long long)blockIdx.x * kRowThreads + (tid - NB);
    const bool have = diag ? (tid < n) : (row < m);
    const bool writer = !diag || blockIdx.x == 0;                 // the diagonal block is written once

    T reg[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) reg[j] = (have && j < n) ? A[row + (long long)j * lda] : T(0);
    if (tid == 0) sh.first_zero = 0;
    bool alive = have;

#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        if (diag && tid == k) {                                   // row k of U is final: publish it
            const T pv = reg[0];
#pragma unroll
            for (int j = 0; j < NB; ++j) sh.uw[k][j] = reg[j];    // reg[j] == 0 beyond column n
            sh.rinv[k] = (pv != T(0)) ? rcp_rn(pv) : T(1);        // :316-320 / :321-327
            if (pv == T(0) && sh.first_zero == 0) sh.first_zero = k + 1;
            if (writer) {
                const int rem = n - k;
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if (j < rem) A[k + (long long)(k + j) * lda] = reg[j];
            }
            alive = false;
        }
        __syncthreads();
        if (alive) {
            const T l = reg[0] * sh.rinv[k];
            if (writer) A[row + (long long)k * lda] = l;
            const T nl = -l;
#pragma unroll
            for (int j = 1; j < NB; ++j) reg[j - 1] = fma(nl, sh.uw[k][j], reg[j]);   // slide the window
            reg[NB - 1] = T(0);
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0 && sh.first_zero != 0 && *info == 0)
        *info = -(col_offset + sh.first_zero);                    // Julia >= 1.11: negative for NoPivot
}

template <typename T, int NB>
int launch_inst(rfb_ctx *ctx, T *A, int m, int n, int64_t lda, int64_t *info, int64_t col_offset) {
    const int rows_below = m - n;
    const int G = rows_below > 0 ? (rows_below + kRowThreads - 1) / kRowThreads : 1;
    RfbLaunchScope scope(ctx, RFB_KC_PANEL, (double)m * n * n - (double)n * n * n / 3.0);
    panel_nopiv_kernel<T, NB><<<G, NB + kRowThreads, 0, ctx->stream>>>(A, m, n, (long long)lda, (long long *)info,
                                                                     (long long)col_offset);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

// identity pivots for a user-supplied ipiv (src/lu.jl:107-113)
__global__ void iota_kernel(long long *p, long long n, long long first) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = first + i;
}

}  // namespace

template <typename T>
int rfb_launch_panel_nopiv(rfb_ctx *ctx, T *A, int64_t m, int64_t n, int64_t lda, int64_t *info_dev, int64_t col_offset) {
    if (n <= 0 || m <= 0) return RFB_OK;
    if (n > RFB_MAX_NB) return ctx->fail(RFB_ERR_UNSUPPORTED, "panel width %lld > %d", (long long)n, RFB_MAX_NB);
    if (m < n) return ctx->fail(RFB_ERR_ARG, "panel needs m >= n (got %lld x %lld)", (long long)m, (long long)n);
    if (n <= 16) return launch_inst<T, 16>(ctx, A, (int)m, (int)n, lda, info_dev, col_offset);
    if (n <= 32) return launch_inst<T, 32>(ctx, A, (int)m, (int)n, lda, info_dev, col_offset);
    return launch_inst<T, 64>(ctx, A, (int)m, (int)n, lda, info_dev, col_offset);
}
template int rfb_launch_panel_nopiv<double>(rfb_ctx *, double *, int64_t, int64_t, int64_t, int64_t *, int64_t);
template int rfb_launch_panel_nopiv<float>(rfb_ctx *, float *, int64_t, int64_t, int64_t, int64_t *, int64_t);

int rfb_launch_iota(rfb_ctx *ctx, int64_t *p_dev, int64_t n, int64_t first) {
    if (n <= 0) return RFB_OK;
    RfbLaunchScope scope(ctx, RFB_KC_OTHER);
    iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((long long *)p_dev, n, first);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}
