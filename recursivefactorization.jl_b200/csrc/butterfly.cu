// butterfly.cu -- the random butterfly transform of the 🦋 solver (src/butterflylu.jl).
//
//   rfb_launch_butterfly_mul : `🦋mul!(A, uv)` (:93-113) = A <- U' A V, two butterfly levels
//                              (`🦋mul_level!`, :59-91) FUSED into one pass over the matrix;
//   rfb_launch_butterfly_vec : the two matrix-vector products of `🦋solve!` (:50-52),
//                              tmp = U' b and x = V tmp, applied in their factored O(n) form
//                              instead of through the dense U, V that `materializeUV` (:149-178)
//                              builds on the CPU.
//
// B200 design: the reference makes five sweeps (four quadrant sweeps of level 1, one sweep of
// level 2), each reading and writing its block.  Both levels only couple the 16 entries
// {m, m+M/4, m+M/2, m+3M/4} x {n, n+M/4, n+M/2, n+3M/4}, so one thread loads those 16 values, applies
// level 1 (inside each quadrant) and level 2 (across quadrants) in registers and stores them: the
// matrix is read once and written once (2*s*M^2 bytes, HBM-bound), every access is coalesced down
// the columns.  The arithmetic is the reference's own sequence of additions and (u * C) * v
// products, so the result is bit-identical to the two-level CPU loop.
#include "rfb_internal.h"

namespace {

// Round-to-nearest intrinsics: never contracted into an FMA.  Level 2 adds values that level 1 just
// produced as products; a fused multiply-add there would differ from the reference's two sweeps
// (which round the products when they store them) in the last bit.
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }

// 🦋mul_level! on one 2 x 2 stencil (src/butterflylu.jl:64-88); a = {A11, A21, A12, A22}
template <typename T>
__device__ __forceinline__ void level_2x2(T &a11, T &a21, T &a12, T &a22, T u1, T u2, T v1, T v2) {
    const T t1 = add_rn(a11, a12), t2 = add_rn(a21, a22), t3 = sub_rn(a11, a12), t4 = sub_rn(a21, a22);
    const T c11 = add_rn(t1, t2), c21 = sub_rn(t1, t2), c12 = add_rn(t3, t4), c22 = sub_rn(t3, t4);
    a11 = mul_rn(mul_rn(u1, c11), v1);
    a21 = mul_rn(mul_rn(u2, c21), v1);
    a12 = mul_rn(mul_rn(u1, c12), v2);
    a22 = mul_rn(mul_rn(u2, c22), v2);
}

constexpr int kBfThreads = 128;
constexpr int kBfCols = 4;         // n0 values per thread (loop), amortises the u loads

template <typename T>
__global__ void __launch_bounds__(kBfThreads)
butterfly_mul_kernel(T *__restrict__ A, int M, long long lda, const T *__restrict__ uv) {
    const int q = M >> 2, h = M >> 1;
    const int m0 = blockIdx.x * kBfThreads + threadIdx.x;
    if (m0 >= q) return;
    // level-1 row scales: quadrant rows [0,h) use U1 = uv[0:h], rows [h,M) use U2 = uv[M:M+h]   (:98-101)
    const T u1a = uv[m0], u1b = uv[m0 + q], u2a = uv[M + m0], u2b = uv[M + m0 + q];
    // level-2 row scales U = uv[2M:3M]                                                           (:108)
    const T *U = uv + 2 * (long long)M, *V = uv + 3 * (long long)M;
    const T ua = U[m0], ub = U[m0 + q], uc = U[m0 + h], ud = U[m0 + h + q];
    for (int c = 0; c < kBfCols; ++c) {
        const int n0 = blockIdx.y * kBfCols + c;
        if (n0 >= q) return;
        // level-1 column scales V1 = uv[h:M], V2 = uv[M+h:2M]
        const T v1a = uv[h + n0], v1b = uv[h + n0 + q], v2a = uv[M + h + n0], v2b = uv[M + h + n0 + q];
        const T va = V[n0], vb = V[n0 + q], vc = V[n0 + h], vd = V[n0 + h + q];
        T x[4][4];   // x[row group][col group], groups at offsets 0, q, h, h+q
        const long long ro[4] = {m0, m0 + q, m0 + h, m0 + h + q};
        const long long co[4] = {n0, n0 + q, n0 + h, n0 + h + q};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i][j] = A[ro[i] + co[j] * lda];
        // level 1: the four quadrants (:103-106)
        level_2x2(x[0][0], x[1][0], x[0][1], x[1][1], u1a, u1b, v1a, v1b);   // top-left     (U1, V1)
        level_2x2(x[2][0], x[3][0], x[2][1], x[3][1], u2a, u2b, v1a, v1b);   // bottom-left  (U2, V1)
        level_2x2(x[0][2], x[1][2], x[0][3], x[1][3], u1a, u1b, v2a, v2b);   // top-right    (U1, V2)
        level_2x2(x[2][2], x[3][2], x[2][3], x[3][3], u2a, u2b, v2a, v2b);   // bottom-right (U2, V2)
        // level 2: whole matrix, rows (i, i + h), columns (j, j + h) (:111)
        level_2x2(x[0][0], x[2][0], x[0][2], x[2][2], ua, uc, va, vc);
        level_2x2(x[1][0], x[3][0], x[1][2], x[3][2], ub, ud, va, vc);
        level_2x2(x[0][1], x[2][1], x[0][3], x[2][3], ua, uc, vb, vd);
        level_2x2(x[1][1], x[3][1], x[1][3], x[3][3], ub, ud, vb, vd);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) A[ro[i] + co[j] * lda] = x[i][j];
    }
}

// One butterfly block B = [D(y) D(z); D(y) -D(z)] (src/butterflylu.jl:134-147) applied to the pair
// (p, r) = (x[i], x[i + half]):   B' x -> (y (p + r), z (p - r));   B x -> (y p + z r, y p - z r).
template <typename T>
__device__ __forceinline__ void bt(T &p, T &r, T y, T z) { const T s = p + r, d = p - r; p = y * s; r = z * d; }
template <typename T>
__device__ __forceinline__ void bn(T &p, T &r, T y, T z) { const T a = y * p, b = z * r; p = a + b; r = a - b; }

// which == 0:  b <- U' b = Bu1' (Bu2' b)      (U = Bu2 Bu1, :176)
// which == 1:  b <- V  b = Bv2 (Bv1 b)        (V = Bv2 Bv1, :177)
template <typename T>
__global__ void __launch_bounds__(kBfThreads)
butterfly_vec_kernel(T *__restrict__ B, int M, int nrhs, long long ldb, const T *__restrict__ uv, int which) {
    const int q = M >> 2, h = M >> 1;
    const int i = blockIdx.x * kBfThreads + threadIdx.x;
    if (i >= q) return;
    T *b = B + (long long)blockIdx.y * ldb;
    T x0 = b[i], x1 = b[i + q], x2 = b[i + h], x3 = b[i + h + q];
    if (which == 0) {
        const T *u1 = uv, *u2 = uv + M, *u = uv + 2 * (long long)M;
        bt(x0, x1, u1[i], u1[i + q]);            // Bu2' : top half with U1, bottom half with U2
        bt(x2, x3, u2[i], u2[i + q]);
        bt(x0, x2, u[i], u[i + h]);              // Bu1'
        bt(x1, x3, u[i + q], u[i + q + h]);
    } else {
        const T *v1 = uv + h, *v2 = uv + M + h, *v = uv + 3 * (long long)M;
        bn(x0, x2, v[i], v[i + h]);              // Bv1
        bn(x1, x3, v[i + q], v[i + q + h]);
        bn(x0, x1, v1[i], v1[i + q]);            // Bv2
        bn(x2, x3, v2[i], v2[i + q]);
    }
    b[i] = x0; b[i + q] = x1; b[i + h] = x2; b[i + h + q] = x3;
}

template <typename T>
__global__ void set_diag_kernel(T *A, long long lda, int i0, int i1, T value) {
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < i1) A[i + i * lda] = value;
}

}  // namespace

// identity corner of `pad!` (src/butterflylu.jl:193-195): A[i, i] = value for i0 <= i < i1
template <typename T>
int rfb_launch_set_diag(rfb_ctx *ctx, T *A, int64_t lda, int64_t i0, int64_t i1, T value) {
    if (i1 <= i0) return RFB_OK;
    RfbLaunchScope scope(ctx, RFB_KC_OTHER);
    set_diag_kernel<T><<<(unsigned)((i1 - i0 + 63) / 64), 64, 0, ctx->stream>>>(A, (long long)lda, (int)i0, (int)i1, value);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}
template int rfb_launch_set_diag<double>(rfb_ctx *, double *, int64_t, int64_t, int64_t, double);
template int rfb_launch_set_diag<float>(rfb_ctx *, float *, int64_t, int64_t, int64_t, float);

template <typename T>
int rfb_launch_butterfly_mul(rfb_ctx *ctx, T *A, int64_t M, int64_t lda, const T *uv) {
    if (M <= 0) return RFB_OK;
    if (M % 4 != 0) return ctx->fail(RFB_ERR_ARG, "butterfly transform needs a size divisible by 4 (got %lld); pad first", (long long)M);
    if (M > 0x7fffffffLL) return ctx->fail(RFB_ERR_UNSUPPORTED, "dimension exceeds int32");
    const int q = (int)(M / 4);
    dim3 grid((q + kBfThreads - 1) / kBfThreads, (q + kBfCols - 1) / kBfCols);
    if (grid.y > 65535) return ctx->fail(RFB_ERR_UNSUPPORTED, "matrix too large for the butterfly grid");
    RfbLaunchScope scope(ctx, RFB_KC_OTHER, 2.0 * sizeof(T) * (double)M * (double)M);
    butterfly_mul_kernel<T><<<grid, kBfThreads, 0, ctx->stream>>>(A, (int)M, (long long)lda, uv);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}
template int rfb_launch_butterfly_mul<double>(rfb_ctx *, double *, int64_t, int64_t, const double *);
template int rfb_launch_butterfly_mul<float>(rfb_ctx *, float *, int64_t, int64_t, const float *);

template <typename T>
int rfb_launch_butterfly_vec(rfb_ctx *ctx, T *B, int64_t M, int64_t nrhs, int64_t ldb, const T *uv, int which) {
    if (M <= 0 || nrhs <= 0) return RFB_OK;
    if (M % 4 != 0) return ctx->fail(RFB_ERR_ARG, "butterfly transform needs a size divisible by 4 (got %lld); pad first", (long long)M);
    if (nrhs > 65535) return ctx->fail(RFB_ERR_UNSUPPORTED, "more than 65535 right-hand sides");
    const int q = (int)(M / 4);
    dim3 grid((q + kBfThreads - 1) / kBfThreads, (unsigned)nrhs);
    RfbLaunchScope scope(ctx, RFB_KC_OTHER, 2.0 * sizeof(T) * (double)M * (double)nrhs);
    butterfly_vec_kernel<T><<<grid, kBfThreads, 0, ctx->stream>>>(B, (int)M, (int)nrhs, (long long)ldb, uv, which);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}
template int rfb_launch_butterfly_vec<double>(rfb_ctx *, double *, int64_t, int64_t, int64_t, const double *, int);
template int rfb_launch_butterfly_vec<float>(rfb_ctx *, float *, int64_t, int64_t, int64_t, const float *, int);
