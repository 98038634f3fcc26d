// trsm.cu -- K3: B <- unitlower(L)^-1 B, left side, in place.
//
// Replaces TriangularSolve.ldiv!(UnitLowerTriangular(A11), A12, thread) at src/lu.jl:235 and :153
// (TriangularSolve.jl is an un-vendored dependency of the reference; its contract is forward
// substitution with an implied unit diagonal that reads only the strict lower triangle).
//
// Blocked recursively on the host: diagonal TB x TB blocks are solved by the kernel below, the
// off-diagonal work (all but a TB/k fraction of the flops) is the K4 GEMM.
#include "rfb_internal.h"

namespace {

// Diagonal block solve.  One thread owns one right-hand-side column and keeps its TB values in
// registers for the whole solve; L is staged once in shared memory and read as warp-wide
// broadcasts.  Memory is touched in exactly two round trips: all TB loads of the L tile are in
// flight together, then all TB loads of the thread's own column (each thread reads and writes whole
// 32-byte sectors of its column, and the 128-byte lines are shared through L1, so going straight
// from global memory to registers costs no extra DRAM traffic).
template <typename T, int TB, int COLS>
__global__ void __launch_bounds__(COLS)
trsm_diag_kernel(const T *__restrict__ L, int kb, T *__restrict__ B, long long nrhs, long long lda) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sL = reinterpret_cast<T *>(smem_raw);          // [TB cols][TB rows], zero outside strict lower
    static_assert(COLS == TB, "the L tile load assumes one thread per tile row");
    const int tid = threadIdx.x;
    const long long col = (long long)blockIdx.x * COLS + tid;
    const bool active = col < nrhs;
    {
        T tl[TB];
#pragma unroll
        for (int c = 0; c < TB; ++c) tl[c] = (tid < kb && c < kb && tid > c) ? L[tid + (long long)c * lda] : T(0);
#pragma unroll
        for (int c = 0; c < TB; ++c) sL[c * TB + tid] = tl[c];
    }
    T *xcol = B + (active ? col : 0) * lda;
    T x[TB];
#pragma unroll
    for (int r = 0; r < TB; ++r) x[r] = (active && r < kb) ? xcol[r] : T(0);
    __syncthreads();
#pragma unroll
    for (int c = 0; c < TB - 1; ++c) {
        const T nxc = -x[c];
#pragma unroll
        for (int r = c + 1; r < TB; ++r) x[r] = fma(sL[c * TB + r], nxc, x[r]);
    }
    if (active) {
#pragma unroll
        for (int r = 0; r < TB; ++r)
            if (r < kb) xcol[r] = x[r];
    }
}

template <typename T, int TB>
int launch_diag(rfb_ctx *ctx, const T *L, int kb, T *B, int64_t nrhs, int64_t lda) {
    constexpr int COLS = TB;
    constexpr size_t smem = sizeof(T) * (TB * TB);
    auto kern = trsm_diag_kernel<T, TB, COLS>;
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)kern, smem));
    RfbLaunchScope scope(ctx, RFB_KC_TRSM, (double)kb * (double)kb * (double)nrhs);
    kern<<<(unsigned int)((nrhs + COLS - 1) / COLS), COLS, smem, ctx->stream>>>(L, kb, B, nrhs, lda);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

template <typename T>
int trsm_rec(rfb_ctx *ctx, const T *L, int64_t k, T *B, int64_t nrhs, int64_t lda, int tb,
             const rfb_opts *opts) {
    if (k <= tb) {
        if (tb == 32) return launch_diag<T, 32>(ctx, L, (int)k, B, nrhs, lda);
        return launch_diag<T, 64>(ctx, L, (int)k, B, nrhs, lda);
    }
    // split at a multiple of the diagonal block nearest to k/2
    int64_t k1 = ((k / 2 + tb - 1) / tb) * tb;
    if (k1 >= k) k1 = ((k - 1) / tb) * tb;
    RFB_TRY(trsm_rec<T>(ctx, L, k1, B, nrhs, lda, tb, opts));
    RFB_TRY(rfb_launch_gemm<T>(ctx, B + k1, L + k1, B, k - k1, nrhs, k1, lda, opts));
    return trsm_rec<T>(ctx, L + k1 + k1 * lda, k - k1, B + k1, nrhs, lda, tb, opts);
}

}  // namespace

template <typename T>
int rfb_launch_trsm(rfb_ctx *ctx, const T *L, int64_t k, T *B, int64_t nrhs, int64_t lda,
                    const rfb_opts *opts) {
    if (k <= 0 || nrhs <= 0) return RFB_OK;
    int tb = (opts && opts->trsm_block == 32) ? 32 : 64;
    return trsm_rec<T>(ctx, L, k, B, nrhs, lda, tb, opts);
}

template int rfb_launch_trsm<double>(rfb_ctx *, const double *, int64_t, double *, int64_t, int64_t, const rfb_opts *);
template int rfb_launch_trsm<float>(rfb_ctx *, const float *, int64_t, float *, int64_t, int64_t, const rfb_opts *);
