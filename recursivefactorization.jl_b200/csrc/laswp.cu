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

__global__ void ipiv_shift_kernel(long long *ipiv, long long n, long long shift) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ipiv[i] += shift;
}

}  // namespace

template <typename T>
int rfb_launch_laswp(rfb_ctx *ctx, T *A, int64_t ncols, int64_t lda, const int64_t *ipiv_dev,
                     int64_t npiv, int64_t ipiv_sub) {
    if (ncols <= 0 || npiv <= 0) return RFB_OK;
    const unsigned int blocks = (unsigned int)((ncols + kLaswpThreads - 1) / kLaswpThreads);
    RfbLaunchScope scope(ctx, RFB_KC_LASWP, 4.0 * sizeof(T) * (double)npiv * (double)ncols);
    laswp_ipiv_kernel<T><<<blocks, kLaswpThreads, 0, ctx->stream>>>(A, ncols, lda, (const long long *)ipiv_dev,
                                                                     (int)npiv, ipiv_sub);
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
