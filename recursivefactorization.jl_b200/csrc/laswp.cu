// laswp.cu -- K2: row interchanges (src/lu.jl:164-188, apply_permutation!).
//
// Sequential-swap semantics are kept exactly: for i = 0..npiv-1 swap rows i and ipiv[i]-1-sub of
// the column block; a later swap sees the effect of the earlier ones and ipiv values may repeat.
#include "rfb_internal.h"

namespace {

constexpr int kLaswpThreads = 128;
constexpr int kPivChunk = 1024;

// One thread per column, pivots staged through shared memory in chunks.  Swaps of one column are
// independent of every other column (the reference's threaded form :164-175 makes the same cut).
template <typename T>
__global__ void __launch_bounds__(kLaswpThreads)
laswp_ipiv_kernel(T *__restrict__ A, long long ncols, long long lda, const long long *__restrict__ ipiv,
                  int npiv, long long sub) {
    __shared__ int s_piv[kPivChunk];
    const long long col = (long long)blockIdx.x * kLaswpThreads + threadIdx.x;
    T *c = A + (col < ncols ? col : 0) * lda;
    for (int c0 = 0; c0 < npiv; c0 += kPivChunk) {
        const int cn = npiv - c0 < kPivChunk ? npiv - c0 : kPivChunk;
        __syncthreads();
        for (int i = threadIdx.x; i < cn; i += kLaswpThreads) s_piv[i] = (int)(ipiv[c0 + i] - 1 - sub);
        __syncthreads();
        if (col < ncols) {
            for (int i = 0; i < cn; ++i) {
                const int r = s_piv[i];
                const int p = c0 + i;
                if (r != p) {                       // serial form skips i' == i (:180)
                    const T t = c[p];
                    c[p] = c[r];
                    c[r] = t;
                }
            }
        }
    }
}

// List-driven form used by the whole-path driver.  K1 already composed the interchanges of each
// panel into a permutation ("row r ends up at row d", at most 2*NB rows per panel), so within one
// panel all reads can be issued before all writes: one warp owns one column, its 32 lanes gather
// up to 128 values in parallel, then scatter them.  Panels are applied in order, which keeps the
// sequential-swap semantics across panels; the pivot rows of a panel are contiguous, so half of
// every gather/scatter is coalesced down the column.  Replaces 2*npiv dependent memory round
// trips per thread of the ipiv-driven kernel by one round trip per panel.
constexpr int kListWarps = 8;
template <typename T>
__global__ void __launch_bounds__(kListWarps * 32)
laswp_list_kernel(T *__restrict__ A, long long ncols, long long lda, const int *__restrict__ perm_dst,
                  const int *__restrict__ perm_src, const int *__restrict__ perm_width, int k0, int k1) {
    const long long colidx = (long long)blockIdx.x * kListWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (colidx >= ncols) return;
    T *col = A + colidx * lda - k0;                 // col[r] addresses absolute row r
    int c = k0;
    int w = perm_width[c];
    while (c < k1 && w > 0) {
        const int nslots = 2 * w;
        int d[4], s[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = lane + 32 * e;
            d[e] = idx < nslots ? perm_dst[2 * c + idx] : -1;
            s[e] = idx < nslots ? perm_src[2 * c + idx] : -1;
        }
        const int cn = c + w;
        const int wn = cn < k1 ? perm_width[cn] : 0;            // prefetch the next panel's width
        T v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (d[e] >= 0 && d[e] != s[e]) ? col[s[e]] : T(0);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (d[e] >= 0 && d[e] != s[e]) col[d[e]] = v[e];
        __syncwarp();
        c = cn;
        w = wn;
    }
}

__global__ void ipiv_shift_kernel(long long *ipiv, long long n, long long shift) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ipiv[i] += shift;
}

}  // namespace

template <typename T>
int rfb_launch_laswp(rfb_ctx *ctx, T *A, int64_t ncols, int64_t lda, const int64_t *ipiv_dev,
                     int64_t npiv, int64_t ipiv_sub) {
    if (ncols <= 0 || npiv <= 0) return RFB_OK;
    if (ctx->dry_run) { ctx->rec(RFB_T_LASWP, A, nullptr, nullptr, ncols, ipiv_sub, ipiv_sub + npiv); return RFB_OK; }
    const unsigned int blocks = (unsigned int)((ncols + kLaswpThreads - 1) / kLaswpThreads);
    RfbLaunchScope scope(ctx, RFB_KC_LASWP, 4.0 * sizeof(T) * (double)npiv * (double)ncols);
    laswp_ipiv_kernel<T><<<blocks, kLaswpThreads, 0, ctx->stream>>>(A, ncols, lda, (const long long *)ipiv_dev,
                                                                     (int)npiv, ipiv_sub);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

template <typename T>
int rfb_launch_laswp_lists(rfb_ctx *ctx, T *A, int64_t ncols, int64_t lda, int64_t k0, int64_t k1) {
    if (ncols <= 0 || k1 <= k0) return RFB_OK;
    if (ctx->dry_run) { ctx->rec(RFB_T_LASWP, A, nullptr, nullptr, ncols, k0, k1); return RFB_OK; }
    const unsigned int blocks = (unsigned int)((ncols + kListWarps - 1) / kListWarps);
    RfbLaunchScope scope(ctx, RFB_KC_LASWP, 4.0 * sizeof(T) * (double)(k1 - k0) * (double)ncols);
    laswp_list_kernel<T><<<blocks, kListWarps * 32, 0, ctx->stream>>>(A, ncols, lda, ctx->perm_dst, ctx->perm_src,
                                                                      ctx->perm_width, (int)k0, (int)k1);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

int rfb_launch_ipiv_shift(rfb_ctx *ctx, int64_t *ipiv_dev, int64_t n, int64_t shift) {
    if (n <= 0) return RFB_OK;
    RfbLaunchScope scope(ctx, RFB_KC_OTHER);
    ipiv_shift_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, ctx->stream>>>((long long *)ipiv_dev, n, shift);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

template int rfb_launch_laswp<double>(rfb_ctx *, double *, int64_t, int64_t, const int64_t *, int64_t, int64_t);
template int rfb_launch_laswp<float>(rfb_ctx *, float *, int64_t, int64_t, const int64_t *, int64_t, int64_t);
template int rfb_launch_laswp_lists<double>(rfb_ctx *, double *, int64_t, int64_t, int64_t, int64_t);
template int rfb_launch_laswp_lists<float>(rfb_ctx *, float *, int64_t, int64_t, int64_t, int64_t);
