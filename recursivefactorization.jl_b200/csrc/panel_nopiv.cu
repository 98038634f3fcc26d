// panel_nopiv.cu -- K1': LU of a tall m x n panel (n <= 64) WITHOUT pivoting, one launch.
//
// Reference: `_generic_lufact!(A, Val(false), ipiv, info)` (src/lu.jl:290-338 with Pivot = false:
// kp = k, no interchange, reciprocal scaling :317-320, rank-1 update :330-334, a zero pivot is
// recorded as NEGATIVE info and the column is left unscaled :321-327) and the bottom levels of
// `reckernel!` (:189-263) below the leaf width.  The operation order per element is the unblocked
// loop's own (multiply by the correctly rounded reciprocal, then one FMA per earlier column in
// column order), so the result is bit-identical to that loop.
//
// B200 design: without a pivot search there is nothing to exchange between CTAs.
//   phase A -- every CTA factors the n x n diagonal block REDUNDANTLY in shared memory (256 threads,
//              one __syncthreads per column; 32 KB read from L2), so no CTA waits for another one;
//   phase B -- each of the CTA's 128 row threads owns one row below the diagonal block, loaded into
//              registers BEFORE phase A so the HBM latency hides behind it, and eliminates it against
//              the finished U with NO barrier at all: fully unrolled triangular loop, U values are
//              shared-memory broadcasts, every update a register FMA.
// The panel is read once and written once (2*s*m*n bytes).  (A first version published row k from the
// registers of one thread at every step: 64 dependent STS per column on the critical path, 74 us per
// 16384 x 64 panel -- profiles/r01_panel_nopiv_v1.txt.)
#include "rfb_internal.h"

namespace {

constexpr int kNpThreads = 256;      // threads per CTA (phase A uses all of them)
constexpr int kRowThreads = 128;     // rows below the diagonal block per CTA (phase B)

__device__ __forceinline__ double rcp_rn(double v) { return __drcp_rn(v); }
__device__ __forceinline__ float rcp_rn(float v) { return __frcp_rn(v); }

template <typename T, int NB>
struct NoPivShared {
    T d[NB][NB];       // diagonal block, row-major: d[i][j]; on exit of phase A row k holds U[k][k..] (j >= k)
    T rinv[NB];        // 1 / U[k][k], or 1 when the pivot is exactly zero (column stays unscaled, :321-327)
    int first_zero;    // 1-based column of the first exactly-zero pivot, 0 = none
    unsigned int ticket;
};

template <typename T, int NB>
__global__ void __launch_bounds__(kNpThreads)
panel_nopiv_kernel(T *__restrict__ A, int m, int n, long long lda, long long *__restrict__ info,
                   long long col_offset, unsigned int *__restrict__ loaded_counter) {
    __shared__ __align__(16) NoPivShared<T, NB> sh;
    const int tid = threadIdx.x;

    // phase B operands first: the loads are in flight while phase A runs
    const long long row = (long long)n + (long long)blockIdx.x * kRowThreads + tid;
    const bool have = tid < kRowThreads && row < m;
    T reg[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) reg[j] = (have && j < n) ? A[row + (long long)j * lda] : T(0);

    // ---- phase A: unblocked LU of the diagonal block in shared memory ---------------------------
    for (int e = tid; e < NB * NB; e += kNpThreads) {
        const int i = e / NB, j = e % NB;
        sh.d[i][j] = (i < n && j < n) ? A[i + (long long)j * lda] : T(0);
    }
    if (tid < NB) sh.rinv[tid] = T(1);
    if (tid == 0) sh.first_zero = 0;
    // The factored diagonal block goes back IN PLACE, but CTAs of a later wave may not have read the
    // original yet: the CTA that is LAST to finish loading it (ticket == gridDim.x - 1) is the one
    // that writes it, and it resets the counter for the next launch.  Nobody ever waits.
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        sh.ticket = atomicAdd(loaded_counter, 1u);
        const T p0 = sh.d[0][0];
        sh.rinv[0] = (p0 != T(0)) ? rcp_rn(p0) : T(1);           // :316-320
        if (p0 == T(0)) sh.first_zero = 1;
    }
    __syncthreads();
    const bool writer = sh.ticket == gridDim.x - 1;
    const int tx = tid % NB, ty = tid / NB;                      // column / row phase of this thread
    constexpr int TY = kNpThreads / NB;
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        __syncthreads();
        const T r = sh.rinv[k];
        // Branch-free trailing update of columns tx > k.  Column k itself is never rewritten in place: after
        // the loop d[i][k] (i > k) still holds the value the unblocked loop would scale at step k, so L is
        // formed once at the end as d[i][k] * rinv[k] -- the same product, bit for bit.  (Storing L from
        // inside this loop made it divergent and 6x slower: profiles/r01_panel_nopiv_v2.txt.)
        if (tx > k && tx < n) {
            const T ukj = sh.d[k][tx];
            int i = k + 1 + ty;
            if (tx == k + 1 && ty == 0) {
                // this thread produces the NEXT pivot first and computes its correctly rounded reciprocal at
                // once, so that dependent chain overlaps the other threads' updates of this step
                const T nd = fma(-(sh.d[k + 1][k] * r), ukj, sh.d[k + 1][tx]);
                sh.d[k + 1][tx] = nd;
                sh.rinv[k + 1] = (nd != T(0)) ? rcp_rn(nd) : T(1);   // :316-320
                if (nd == T(0) && sh.first_zero == 0) sh.first_zero = k + 2;
                i += TY;
            }
#pragma unroll 4
            for (; i < n; i += TY) sh.d[i][tx] = fma(-(sh.d[i][k] * r), ukj, sh.d[i][tx]);   // :330-334
        }
    }
    __syncthreads();
    if (writer) {                                                 // L\U of the diagonal block, written once
        for (int e = tid; e < n * n; e += kNpThreads) {
            const int i = e % n, j = e / n;
            A[i + (long long)j * lda] = (j >= i) ? sh.d[i][j] : sh.d[i][j] * sh.rinv[j];
        }
        if (tid == 0) {
            *loaded_counter = 0u;
            if (sh.first_zero != 0 && *info == 0) *info = -(col_offset + sh.first_zero);   // Julia >= 1.11: negative
        }
    }

    // ---- phase B: one row per thread against the finished U, no barriers ------------------------
    if (!have) return;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        const T l = reg[k] * sh.rinv[k];                          // columns >= n: 0 * 1
        reg[k] = l;
        const T nl = -l;
#pragma unroll
        for (int j = k + 1; j < NB; ++j) reg[j] = fma(nl, sh.d[k][j], reg[j]);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j)
        if (j < n) A[row + (long long)j * lda] = reg[j];
}

template <typename T, int NB>
int launch_inst(rfb_ctx *ctx, T *A, int m, int n, int64_t lda, int64_t *info, int64_t col_offset) {
    const int rows_below = m - n;
    const int G = rows_below > 0 ? (rows_below + kRowThreads - 1) / kRowThreads : 1;
    RfbLaunchScope scope(ctx, RFB_KC_PANEL, (double)m * n * n - (double)n * n * n / 3.0);
    panel_nopiv_kernel<T, NB><<<G, kNpThreads, 0, ctx->stream>>>(A, m, n, (long long)lda, (long long *)info,
                                                              (long long)col_offset, &ctx->xchg->pad[0]);
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
    if (ctx->dry_run) { ctx->rec(RFB_T_PANEL_NOPIV, A, nullptr, nullptr, m, n, col_offset); return RFB_OK; }
    if (n <= 16) return launch_inst<T, 16>(ctx, A, (int)m, (int)n, lda, info_dev, col_offset);
    if (n <= 32) return launch_inst<T, 32>(ctx, A, (int)m, (int)n, lda, info_dev, col_offset);
    return launch_inst<T, 64>(ctx, A, (int)m, (int)n, lda, info_dev, col_offset);
}
template int rfb_launch_panel_nopiv<double>(rfb_ctx *, double *, int64_t, int64_t, int64_t, int64_t *, int64_t);
template int rfb_launch_panel_nopiv<float>(rfb_ctx *, float *, int64_t, int64_t, int64_t, int64_t *, int64_t);

int rfb_launch_iota(rfb_ctx *ctx, int64_t *p_dev, int64_t n, int64_t first) {
    if (n <= 0) return RFB_OK;
    if (ctx->dry_run) { ctx->rec(RFB_T_IOTA, nullptr, nullptr, nullptr, n, first, 0); return RFB_OK; }
    RfbLaunchScope scope(ctx, RFB_KC_OTHER);
    iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((long long *)p_dev, n, first);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}
