// umma_probe.cu -- standalone probe of tcgen05.mma kind::tf32 operand layouts (one CTA, no TMA).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
// mode 0: A K-major, B K-major      mode 1: A MN-major (as gemm_tc32.cu assumes), B K-major
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long make_desc(unsigned int saddr, unsigned int lbo, unsigned int sbo, unsigned int layout = 2) {
    unsigned long long d = (unsigned long long)((saddr & 0x3FFFFu) >> 4);
    d |= (unsigned long long)(lbo >> 4) << 16;
    d |= (unsigned long long)(sbo >> 4) << 32;
    d |= 1ull << 46;
    d |= (unsigned long long)layout << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1)
probe(const float *A, const float *B, float *D, int mode, unsigned int idesc, unsigned int *dbg) {
    // A: 128 x 32 (row i, col k) row-major input; B: 32 x 128 (k, j) row-major input; D: 128 x 128 row-major
    extern __shared__ unsigned char raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    float *sA = reinterpret_cast<float *>(base);              // 16 KB
    float *sB = reinterpret_cast<float *>(base + 16384);      // 16 KB
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(base + 32768);
    unsigned int *slot = reinterpret_cast<unsigned int *>(base + 32768 + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int idx = tid; idx < 128 * 32; idx += 128) {
        const int r = idx / 32, k = idx % 32;
        // B operand: N x K, K-major, 128B swizzle: row n at n*128, 16B chunk (k/4) ^ (n%8)
        sB[(r * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4) / 4] = B[k * 128 + r];
        if (mode == 0) {
            sA[(r * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4) / 4] = A[r * 32 + k];
        } else {
            // A operand: MN-major: 4 chunks of 32 rows (LBO = 4096 B); inside: k-row kk at kk*128 B, m%32 floats, swizzled by kk%8
            // SWIZZLE_128B_BASE32B: 32-byte chunk (ml/8) ^ (k%4), k-rows dense at 128 B
            const int chunk = r >> 5, ml = r & 31;
            sA[(chunk * 4096 + k * 128 + (((ml >> 3) ^ (k & 3)) << 5) + (ml & 7) * 4) / 4] = A[r * 32 + k];
        }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned int tmem = *slot;
    if (tid == 0) {
        dbg[0] = tmem;
        const unsigned int a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int ks = 0; ks < 4; ++ks) {
            unsigned long long da = mode == 0 ? make_desc(a0 + ks * 32, 16, 1024) : make_desc(a0 + ks * 1024, 4096, 512, 1);
            unsigned long long db = make_desc(b0 + ks * 32, 16, 1024);
            unsigned int acc = ks != 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    // everyone waits for the MMAs
    asm volatile(
        "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned int taddr = tmem + ((unsigned int)(warp * 32) << 16);
    for (int c0 = 0; c0 < 128; c0 += 16) {
        unsigned int v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr + (unsigned int)c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int u = 0; u < 16; ++u) D[(warp * 32 + lane) * 128 + c0 + u] = __uint_as_float(v[u]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

int main() {
    std::vector<float> A(128 * 32), B(32 * 128), D(128 * 128), W(128 * 128);
    srand(1);
    for (auto &x : A) x = (float)(rand() % 8);
    for (auto &x : B) x = (float)(rand() % 8);
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 128; ++j) {
            float s = 0;
            for (int k = 0; k < 32; ++k) s += A[i * 32 + k] * B[k * 128 + j];
            W[i * 128 + j] = s;
        }
    float *dA, *dB, *dD;
    unsigned int *dbg;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dbg, 64);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    for (int mode = 0; mode < 2; ++mode) {
        for (int variant = 0; variant < 2; ++variant) {
            // variant 0: idesc with M at bits [24,29); variant 1: M at bits [23,28) (older layout)
            unsigned int idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((mode == 1 ? 1u : 0u) << 15) | ((128u >> 3) << 17);
            idesc |= variant == 0 ? ((128u >> 4) << 24) : ((128u >> 4) << 23);
            cudaMemset(dD, 0xFF, D.size() * 4);
            probe<<<1, 128, 40000>>>(dA, dB, dD, mode, idesc, dbg);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            unsigned int t = 0;
            cudaMemcpy(&t, dbg, 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            int bad = 0;
            for (int i = 0; i < 128 * 128; ++i) {
                double d = fabs((double)D[i] - W[i]);
                if (!(d <= maxerr)) maxerr = d;
                bad += d > 1e-3;
            }
            printf("mode %d idesc-variant %d (0x%08x): %s tmem=0x%08x maxerr=%g bad=%d  D[0][0..3]=%g %g %g %g want %g %g %g %g\n", mode,
                   variant, idesc, cudaGetErrorString(e), t, maxerr, bad, D[0], D[1], D[2], D[3], W[0], W[1], W[2], W[3]);
        }
    }
    return 0;
}
