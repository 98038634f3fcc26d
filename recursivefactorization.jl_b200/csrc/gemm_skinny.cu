// gemm_skinny.cu -- K4 for a handful of right-hand-side columns: C (m x n) <- C - A (m x k) B (k x n), n <= 8.
//
// This is the shape of every off-diagonal update when `ldiv!(F, b)` is given a vector or a few
// columns (src/lu.jl:60-64; runtests.jl:21-28, :82, :122-127 all solve with vectors / 3 columns).
// It is a GEMV: 2 flops per 8-byte element of A, HBM-bound (intensity n/4 flop/B), so tensor tiles
// are the wrong tool -- a 128-column DMMA tile would compute 1/128 useful work.  B200 design:
//   * a CTA owns 32 rows (lane = row) so every A access is a full-width coalesced 256-byte row of a
//     column; its 8 warps split k, eight independent column loads in flight per warp;
//   * B values are warp-uniform broadcast loads (L1/L2 resident: k * n * 8 bytes);
//   * the 8 partial sums are combined through shared memory in a FIXED order (deterministic, no
//     atomics) and subtracted from C once -- the accumulate-from-zero-then-add form of
//     `schur_complement!` (src/lu.jl:269-273);
//   * algorithmic bytes: s * (m k + k n + 2 m n), i.e. A is read exactly once.
#include "rfb_internal.h"

namespace {

constexpr int kSkWarps = 8;
constexpr int kSkRows = 32;

template <typename T, int N>
__global__ void __launch_bounds__(kSkWarps * 32)
gemm_skinny_kernel(T *__restrict__ C, const T *__restrict__ A, const T *__restrict__ B, int m, int n, int k, long long lda) {
    __shared__ T part[kSkWarps][N][kSkRows];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * kSkRows + lane;
    const bool live = row < m;
    const T *a = A + (live ? row : 0);
    // warp w takes the k-range [k0, k1): contiguous slices, multiples of 8 columns
    const int per = (((k + kSkWarps - 1) / kSkWarps) + 7) & ~7;
    const int k0 = w * per, k1 = min(k, k0 + per);
    T acc[N];
#pragma unroll
    for (int c = 0; c < N; ++c) acc[c] = T(0);
    int kk = k0;
    for (; kk + 8 <= k1; kk += 8) {
        T av[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) av[u] = live ? a[(long long)(kk + u) * lda] : T(0);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < N; ++c)
                if (c < n) acc[c] = fma(av[u], __ldg(B + (kk + u) + (long long)c * lda), acc[c]);
    }
    for (; kk < k1; ++kk) {
        const T av = live ? a[(long long)kk * lda] : T(0);
#pragma unroll
        for (int c = 0; c < N; ++c)
            if (c < n) acc[c] = fma(av, __ldg(B + kk + (long long)c * lda), acc[c]);
    }
#pragma unroll
    for (int c = 0; c < N; ++c) part[w][c][lane] = acc[c];
    __syncthreads();
    // fixed-order combine: thread (c, lane) with c = warp index
    for (int c = w; c < n; c += kSkWarps) {
        T s = part[0][c][lane];
#pragma unroll
        for (int q = 1; q < kSkWarps; ++q) s += part[q][c][lane];
        if (live) C[row + (long long)c * lda] = C[row + (long long)c * lda] - s;
    }
}

template <typename T, int N>
int launch(rfb_ctx *ctx, T *C, const T *A, const T *B, int64_t m, int64_t n, int64_t k, int64_t lda) {
    RfbLaunchScope scope(ctx, RFB_KC_GEMM, 2.0 * (double)m * (double)n * (double)k);
    gemm_skinny_kernel<T, N><<<(unsigned)((m + kSkRows - 1) / kSkRows), kSkWarps * 32, 0, ctx->stream>>>(
        C, A, B, (int)m, (int)n, (int)k, (long long)lda);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

}  // namespace

template <typename T>
int rfb_launch_gemm_skinny(rfb_ctx *ctx, T *C, const T *A, const T *B, int64_t m, int64_t n, int64_t k, int64_t lda) {
    if (n == 1) return launch<T, 1>(ctx, C, A, B, m, n, k, lda);
    if (n == 2) return launch<T, 2>(ctx, C, A, B, m, n, k, lda);
    if (n <= 4) return launch<T, 4>(ctx, C, A, B, m, n, k, lda);
    return launch<T, 8>(ctx, C, A, B, m, n, k, lda);
}
template int rfb_launch_gemm_skinny<double>(rfb_ctx *, double *, const double *, const double *, int64_t, int64_t, int64_t, int64_t);
template int rfb_launch_gemm_skinny<float>(rfb_ctx *, float *, const float *, const float *, int64_t, int64_t, int64_t, int64_t);
