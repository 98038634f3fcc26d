// gemm.cu -- K4: trailing-submatrix (Schur complement) update  C <- C - A * B.
//
// Replaces schur_complement! (src/lu.jl:265-284): the product is accumulated from zero in its own
// accumulator and added to C exactly once (:269-273).  All three operands are column-major views
// of one allocation and share `lda` (A = L21 is m x k, B = U12 is k x n, C = A22 is m x n).
//
// Float64: sm_100a has no tcgen05 FP64 kind (ptxas: "Unknown modifier .kind::f64"); FP64 tensor
// math is the warp-level DMMA (mma.sync.m8n8k4.f64).  This file holds the *generic* tile kernel:
// any m, n, k, any alignment, operands staged with predicated 8-byte cp.async into padded shared
// tiles.  gemm_tma.cu holds the TMA-fed variant used when the views meet TMA's alignment rules.
// Float32: exact-FP32 FFMA register tiles (RFB_F32_FP32).
#include "rfb_internal.h"

namespace {

// ----------------------------------------------------------------------------------------------
// helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gmem_src, bool valid) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;   // src-size 0 => the 8 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ----------------------------------------------------------------------------------------------
// FP64 generic DMMA kernel: CTA tile 128 x 128 x 16, 8 warps (2 x 4), warp tile 64 x 32,
// 3-stage cp.async pipeline.  Shared tiles are padded so that the DMMA fragment loads are
// conflict-free: A tile [k][m] with row pitch 132 (== 4 mod 16), B tile [n][k] with pitch 20.
// ----------------------------------------------------------------------------------------------
constexpr int GBM = 128, GBN = 128, GBK = 16, GSTAGES = 3;
constexpr int GLDA = GBM + 4, GLDB = GBK + 4;
constexpr int GTHREADS = 256;
constexpr size_t kGemmF64Smem = sizeof(double) * GSTAGES * (GBK * GLDA + GBN * GLDB);

__global__ void __launch_bounds__(GTHREADS, 1)
gemm_f64_generic_kernel(double *__restrict__ C, const double *__restrict__ A, const double *__restrict__ B,
                        int M, int N, int K, long long lda) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);       // [stage][k][GLDA]
    double *sB = sA + GSTAGES * GBK * GLDA;                  // [stage][n][GLDB]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
    const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
    const int KT = (K + GBK - 1) / GBK;

    auto load_stage = [&](int s, int kt) {
        const int k0 = kt * GBK;
        double *dA = sA + s * GBK * GLDA;
        double *dB = sB + s * GBN * GLDB;
        {   // A: threads run down the column (coalesced)
            const int mm = tid & 127, kb = tid >> 7;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int kk = kb + 2 * i;
                const bool ok = (m0 + mm < M) && (k0 + kk < K);
                const double *src = ok ? A + (m0 + mm) + (long long)(k0 + kk) * lda : A;
                cp_async_8(dA + kk * GLDA + mm, src, ok);
            }
        }
        {   // B: k is the contiguous direction
            const int kk = tid & 15, nb = tid >> 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int nn = nb + 16 * i;
                const bool ok = (n0 + nn < N) && (k0 + kk < K);
                const double *src = ok ? B + (k0 + kk) + (long long)(n0 + nn) * lda : B;
                cp_async_8(dB + nn * GLDB + kk, src, ok);
            }
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + GSTAGES - 1;
            if (nk < KT) load_stage(nk % GSTAGES, nk);
            cp_async_commit();
        }
        const double *tA = sA + (kt % GSTAGES) * GBK * GLDA;
        const double *tB = sB + (kt % GSTAGES) * GBN * GLDB;
#pragma unroll
        for (int ks = 0; ks < GBK / 4; ++ks) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = tA[(ks * 4 + q) * GLDA + wm + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = tB[(wn + j * 8 + g) * GLDB + ks * 4 + q];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: C = C - acc  (the "+ (0 - sum)" of src/lu.jl:269-273); the 8 loads of a fragment row
    // are issued before the first store so that they overlap
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + wm + i * 8 + g;
        if (r >= M) continue;
        double cv[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = n0 + wn + j * 8 + 2 * q + e;
                cv[j][e] = c < N ? C[r + (long long)c * lda] : 0.0;
            }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = n0 + wn + j * 8 + 2 * q + e;
                if (c < N) C[r + (long long)c * lda] = cv[j][e] - acc[i][j][e];
            }
    }
}

// ----------------------------------------------------------------------------------------------
// FP32 exact FFMA kernel: CTA tile 128 x 128 x 16, 256 threads, 8 x 8 outputs per thread,
// register-staged double buffering.
// ----------------------------------------------------------------------------------------------
constexpr int SBM = 128, SBN = 128, SBK = 16, STHREADS = 256;

__global__ void __launch_bounds__(STHREADS, 2)
gemm_f32_simt_kernel(float *__restrict__ C, const float *__restrict__ A, const float *__restrict__ B,
                     int M, int N, int K, long long lda) {
    __shared__ __align__(16) float sA[2][SBK][SBM];
    __shared__ __align__(16) float sB[2][SBK][SBN + 4];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;          // thread owns rows tx*4..+3 (+64), cols ty*4..+3 (+64)
    const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;
    const int KT = (K + SBK - 1) / SBK;

    // global -> register staging maps
    const int a_m = tid & 127, a_k = tid >> 7;       // A: 8 loads, k = a_k + 2 i
    const int b_k = tid & 15, b_n = tid >> 4;        // B: 8 loads, n = b_n + 16 i
    float ra[8], rb[8];

    auto gload = [&](int kt) {
        const int k0 = kt * SBK;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kk = a_k + 2 * i;
            ra[i] = (m0 + a_m < M && k0 + kk < K) ? A[(m0 + a_m) + (long long)(k0 + kk) * lda] : 0.f;
            const int nn = b_n + 16 * i;
            rb[i] = (n0 + nn < N && k0 + b_k < K) ? B[(k0 + b_k) + (long long)(n0 + nn) * lda] : 0.f;
        }
    };
    auto sstore = [&](int s) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            sA[s][a_k + 2 * i][a_m] = ra[i];
            sB[s][b_k][b_n + 16 * i] = rb[i];
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    if (KT > 0) { gload(0); sstore(0); }
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int s = kt & 1;
        if (kt + 1 < KT) gload(kt + 1);
#pragma unroll
        for (int kk = 0; kk < SBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&sA[s][kk][tx * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&sA[s][kk][64 + tx * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&sB[s][kk][ty * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&sB[s][kk][64 + ty * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < KT) sstore(s ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = n0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
        if (c >= N) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
            if (r < M) {
                float *p = C + r + (long long)c * lda;
                *p = *p - acc[i][j];
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// DMMA peak: register-only mma.sync.m8n8k4.f64 chains (8 independent accumulator pairs per warp).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dmma_peak_kernel(double *out, int iters, double a0, double b0) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma_884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;   // keep the chain alive
}

__global__ void copy_kernel(double2 *__restrict__ dst, const double2 *__restrict__ src, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i];
}

}  // namespace

// defined in gemm_tc32.cu
int rfb_launch_gemm_f32_tc(rfb_ctx *ctx, float *C, const float *A, const float *B, int64_t m, int64_t n, int64_t k,
                           int64_t lda, bool *handled);
// defined in gemm_tma.cu
int rfb_launch_gemm_f64_tma(rfb_ctx *ctx, double *C, const double *A, const double *B, int64_t m, int64_t n,
                            int64_t k, int64_t lda, bool *handled);

template <>
int rfb_launch_gemm<double>(rfb_ctx *ctx, double *C, const double *A, const double *B, int64_t m, int64_t n,
                            int64_t k, int64_t lda, const rfb_opts *opts) {
    if (m <= 0 || n <= 0 || k <= 0) return RFB_OK;
    if (ctx->dry_run) { ctx->rec(RFB_T_GEMM, C, A, B, m, n, k); return RFB_OK; }
    const int path = opts ? opts->gemm_path : 0;
    if (n <= RFB_SKINNY_MAX_N && path == 0) return rfb_launch_gemm_skinny<double>(ctx, C, A, B, m, n, k, lda);   // GEMV-shaped: HBM-bound
    if (path != 1) {
        bool handled = false;
        RFB_TRY(rfb_launch_gemm_f64_tma(ctx, C, A, B, m, n, k, lda, &handled));
        if (handled) return RFB_OK;
        if (path == 2) return ctx->fail(RFB_ERR_UNSUPPORTED, "gemm_path=2 (TMA) but the views are not TMA-aligned");
    }
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)gemm_f64_generic_kernel, kGemmF64Smem));
    dim3 grid((unsigned int)((m + GBM - 1) / GBM), (unsigned int)((n + GBN - 1) / GBN));
    RfbLaunchScope scope(ctx, RFB_KC_GEMM, 2.0 * (double)m * (double)n * (double)k);
    gemm_f64_generic_kernel<<<grid, GTHREADS, kGemmF64Smem, ctx->stream>>>(C, A, B, (int)m, (int)n, (int)k, lda);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

template <>
int rfb_launch_gemm<float>(rfb_ctx *ctx, float *C, const float *A, const float *B, int64_t m, int64_t n,
                           int64_t k, int64_t lda, const rfb_opts *opts) {
    if (m <= 0 || n <= 0 || k <= 0) return RFB_OK;
    if (ctx->dry_run) { ctx->rec(RFB_T_GEMM, C, A, B, m, n, k); return RFB_OK; }
    if (n <= RFB_SKINNY_MAX_N && !(opts && opts->gemm_path != 0)) return rfb_launch_gemm_skinny<float>(ctx, C, A, B, m, n, k, lda);
    if (opts && opts->f32_mode == RFB_F32_TF32X3) {
        bool handled = false;
        RFB_TRY(rfb_launch_gemm_f32_tc(ctx, C, A, B, m, n, k, lda, &handled));
        if (handled) return RFB_OK;     // unaligned views fall through to the exact FFMA tiles
    }
    dim3 grid((unsigned int)((m + SBM - 1) / SBM), (unsigned int)((n + SBN - 1) / SBN));
    RfbLaunchScope scope(ctx, RFB_KC_GEMM, 2.0 * (double)m * (double)n * (double)k);
    gemm_f32_simt_kernel<<<grid, STHREADS, 0, ctx->stream>>>(C, A, B, (int)m, (int)n, (int)k, lda);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

int rfb_run_dmma_peak(rfb_ctx *ctx, int iters, double *tflops) {
    if (iters <= 0) iters = 20000;
    double *d = nullptr;
    RFB_CUDA(ctx, cudaMalloc(&d, 64));
    const int blocks = ctx->sm_count * 4, threads = 256;
    cudaEvent_t e0, e1;
    RFB_CUDA(ctx, cudaEventCreate(&e0));
    RFB_CUDA(ctx, cudaEventCreate(&e1));
    dmma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters / 10 + 1, 1.0, 1e-3);   // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        RFB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        dmma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters, 1.0, 1e-3);
        RFB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        RFB_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        RFB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double flops = (double)blocks * (threads / 32) * (double)iters * 16.0 * 512.0;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    ctx->launches += 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return RFB_OK;
}

int rfb_run_copy_bench(rfb_ctx *ctx, size_t bytes, int iters, double *gbs) {
    if (iters <= 0) iters = 5;
    bytes &= ~(size_t)15;
    double2 *a = nullptr, *b = nullptr;
    if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) {
        cudaFree(a);
        return ctx->fail(RFB_ERR_NOMEM, "copy bench: cannot allocate 2 x %zu bytes", bytes);
    }
    RFB_CUDA(ctx, cudaMemsetAsync(a, 1, bytes, ctx->stream));
    cudaEvent_t e0, e1;
    RFB_CUDA(ctx, cudaEventCreate(&e0));
    RFB_CUDA(ctx, cudaEventCreate(&e1));
    const int blocks = ctx->sm_count * 16;
    copy_kernel<<<blocks, 256, 0, ctx->stream>>>(b, a, bytes / 16);
    double best = 0;
    for (int rep = 0; rep < iters; ++rep) {
        RFB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        copy_kernel<<<blocks, 256, 0, ctx->stream>>>(b, a, bytes / 16);
        RFB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        RFB_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        RFB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double v = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
        if (v > best) best = v;
    }
    ctx->launches += iters + 1;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    *gbs = best;
    return RFB_OK;
}
