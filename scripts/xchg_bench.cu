// xchg_bench.cu -- micro-benchmark of inter-CTA exchange latency through L2 on B200.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o xchg_bench xchg_bench.cu
// Each CTA publishes an epoch-tagged 16-byte word per step and waits until it has seen the words of
// all G CTAs (the pattern of the panel kernel's pivot exchange).  Reports ns per step.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>

enum Mode { VOLATILE = 0, RELAXED_GPU = 1, ACQ_REL_GPU = 2, CG_PLAIN = 3, RELAXED_SYS = 4 };

template <int MODE>
__device__ __forceinline__ void st16(ulonglong2 *p, unsigned long long a, unsigned long long b) {
    if (MODE == VOLATILE) asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    if (MODE == RELAXED_GPU) asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    if (MODE == ACQ_REL_GPU) asm volatile("st.release.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    if (MODE == CG_PLAIN) asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    if (MODE == RELAXED_SYS) asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
template <int MODE>
__device__ __forceinline__ ulonglong2 ld16(const ulonglong2 *p) {
    ulonglong2 r;
    if (MODE == VOLATILE) asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    if (MODE == RELAXED_GPU) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    if (MODE == ACQ_REL_GPU) asm volatile("ld.acquire.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    if (MODE == CG_PLAIN) asm volatile("ld.global.cv.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    if (MODE == RELAXED_SYS) asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    return r;
}

// slots[parity][cta]; ROWW extra words per CTA published after the header by one thread and read
// by ROWW threads from the "winner" (cta = step % G) after the header gather (2nd round trip).
template <int MODE>
__global__ void __launch_bounds__(128, 1)
xchg_kernel(ulonglong2 *slots, ulonglong2 *rows, int steps, unsigned int epoch0, int roww, long long *cycles_out) {
    const int G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
    __shared__ unsigned long long s_acc;
    long long t0 = clock64();
    unsigned long long acc = 0;
    for (int k = 0; k < steps; ++k) {
        const int par = k & 1;
        const unsigned int epoch = epoch0 + k;
        if (tid == 0) {
            st16<MODE>(&slots[par * 512 + bid], acc + bid, ((unsigned long long)epoch << 32) | bid);
            for (int j = 0; j < roww; ++j) st16<MODE>(&rows[(par * 512 + bid) * 64 + j], acc + j, epoch);
        }
        for (int c = tid; c < G; c += blockDim.x) {
            ulonglong2 h;
            do { h = ld16<MODE>(&slots[par * 512 + c]); } while ((unsigned int)(h.y >> 32) != epoch);
            acc += h.x & 1;
        }
        __syncthreads();
        if (roww > 0) {
            const int w = k % G;
            if (tid < roww) {
                ulonglong2 d;
                do { d = ld16<MODE>(&rows[(par * 512 + w) * 64 + tid]); } while ((unsigned int)d.y != epoch);
                acc += d.x & 1;
            }
            __syncthreads();
        }
    }
    long long t1 = clock64();
    if (tid == 0) { s_acc = acc; if (bid == 0) cycles_out[0] = t1 - t0; }
    if (acc == 0xdeadbeefULL) cycles_out[1] = (long long)s_acc;
}

template <int MODE>
void run(const char *name, int G, int steps, int roww, ulonglong2 *slots, ulonglong2 *rows, long long *dcy, unsigned int &epoch) {
    void *args[] = {&slots, &rows, &steps, &epoch, &roww, &dcy};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchCooperativeKernel((const void *)xchg_kernel<MODE>, dim3(G), dim3(128), args, 0, 0);   // warm
    epoch += steps;
    cudaEventRecord(e0);
    cudaLaunchCooperativeKernel((const void *)xchg_kernel<MODE>, dim3(G), dim3(128), args, 0, 0);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    epoch += steps;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-12s G=%3d roww=%2d : %7.1f ns/step (%s)\n", name, G, roww, ms * 1e6 / steps, cudaGetErrorString(err));
}

int main() {
    ulonglong2 *slots, *rows;
    long long *dcy;
    cudaMalloc(&slots, 2 * 512 * sizeof(ulonglong2));
    cudaMalloc(&rows, 2 * 512 * 64 * sizeof(ulonglong2));
    cudaMalloc(&dcy, 64);
    cudaMemset(slots, 0, 2 * 512 * sizeof(ulonglong2));
    cudaMemset(rows, 0, 2 * 512 * 64 * sizeof(ulonglong2));
    unsigned int epoch = 1;
    const int steps = 2000;
    const int Gs[] = {1, 2, 4, 16, 32, 64, 128, 148};
    for (int roww : {0, 64}) {
        for (int G : Gs) {
            run<VOLATILE>("volatile", G, steps, roww, slots, rows, dcy, epoch);
            run<RELAXED_GPU>("relaxed.gpu", G, steps, roww, slots, rows, dcy, epoch);
            run<ACQ_REL_GPU>("acq_rel.gpu", G, steps, roww, slots, rows, dcy, epoch);
            run<CG_PLAIN>("cg/cv", G, steps, roww, slots, rows, dcy, epoch);
            run<RELAXED_SYS>("relaxed.sys", G, steps, roww, slots, rows, dcy, epoch);
        }
    }
    return 0;
}
