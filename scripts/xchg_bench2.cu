// xchg_bench2.cu -- closer mimic of the panel kernel's per-column exchange, to pick its layout.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o xchg_bench2 xchg_bench2.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>

template <int CV>
__device__ __forceinline__ void st16(ulonglong2 *p, unsigned long long a, unsigned long long b) {
    if (CV) asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    else asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
template <int CV>
__device__ __forceinline__ ulonglong2 ld16(const ulonglong2 *p) {
    ulonglong2 r;
    if (CV) asm volatile("ld.global.cv.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    else asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    return r;
}

// HS = header stride in 16-byte words (1 = packed, 8 = one 128-byte line per CTA)
// PUB = 0: one thread stores header + 64 row words; 1: row staged in smem, one warp stores it
// ALLROWS = 1: single round trip, every CTA reads every CTA's row (small G only)
template <int CV, int HS, int PUB, int ALLROWS>
__global__ void __launch_bounds__(128, 1)
xchg_kernel(ulonglong2 *hdr, ulonglong2 *rows, int steps, unsigned int epoch0, long long *sink) {
    const int G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double s_row[64];
    __shared__ double s_u[2][64];
    __shared__ unsigned int s_w[2];
    double regv = tid * 0.001;
    unsigned long long acc = 0;
    for (int k = 0; k < steps; ++k) {
        const int par = k & 1;
        const unsigned int epoch = epoch0 + k;
        const int pub = (k * 37 + bid * 5) & 127;            // publishing thread moves around
        // pretend local reduce
        __syncthreads();
        if (PUB == 0) {
            if (tid == pub) {
                st16<CV>(&hdr[(par * 512 + bid) * HS], __double_as_longlong(regv) + bid, ((unsigned long long)epoch << 32) | bid);
#pragma unroll
                for (int j = 0; j < 64; ++j) st16<CV>(&rows[(par * 512 + bid) * 64 + j], acc + j, epoch);
            }
        } else {
            if (tid == pub) {
                st16<CV>(&hdr[(par * 512 + bid) * HS], __double_as_longlong(regv) + bid, ((unsigned long long)epoch << 32) | bid);
#pragma unroll
                for (int j = 0; j < 64; ++j) s_row[j] = regv + j;
            }
            __syncwarp();
            if (warp == (pub >> 5)) {
                st16<CV>(&rows[(par * 512 + bid) * 64 + lane], __double_as_longlong(s_row[lane]), epoch);
                st16<CV>(&rows[(par * 512 + bid) * 64 + lane + 32], __double_as_longlong(s_row[lane + 32]), epoch);
            }
        }
        if (!ALLROWS) {
            unsigned int best = 0;
            for (int c = tid; c < G; c += blockDim.x) {
                ulonglong2 h;
                do { h = ld16<CV>(&hdr[(par * 512 + c) * HS]); } while ((unsigned int)(h.y >> 32) != epoch);
                best = (unsigned int)h.y;
            }
            if (tid == 0) s_w[par] = (k * 7) % G;            // pretend reduce result
            acc += best;
            __syncthreads();
            const int w = s_w[par];
            if (tid < 64) {
                ulonglong2 d;
                do { d = ld16<CV>(&rows[(par * 512 + w) * 64 + tid]); } while ((unsigned int)d.y != epoch);
                s_u[par][tid] = __longlong_as_double(d.x);
            }
            __syncthreads();
        } else {
            // every thread fetches a strided share of all G rows (+ headers)
            for (int idx = tid; idx < G * 64; idx += blockDim.x) {
                const int c = idx >> 6, j = idx & 63;
                ulonglong2 d;
                do { d = ld16<CV>(&rows[(par * 512 + c) * 64 + j]); } while ((unsigned int)d.y != epoch);
                if (c == (k * 7) % G) s_u[par][j] = __longlong_as_double(d.x);
            }
            for (int c = tid; c < G; c += blockDim.x) {
                ulonglong2 h;
                do { h = ld16<CV>(&hdr[(par * 512 + c) * HS]); } while ((unsigned int)(h.y >> 32) != epoch);
                acc += (unsigned int)h.y;
            }
            __syncthreads();
        }
        regv += s_u[par][tid & 63] * 1e-9;
    }
    if (acc == 0xdeadbeefULL || regv == 1.2345) sink[0] = (long long)acc;
}

template <int CV, int HS, int PUB, int ALLROWS>
void run(int G, int steps, ulonglong2 *hdr, ulonglong2 *rows, long long *sink, unsigned int &epoch) {
    void *args[] = {&hdr, &rows, &steps, &epoch, &sink};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto kern = xchg_kernel<CV, HS, PUB, ALLROWS>;
    cudaLaunchCooperativeKernel((const void *)kern, dim3(G), dim3(128), args, 0, 0);
    epoch += steps;
    cudaEventRecord(e0);
    cudaLaunchCooperativeKernel((const void *)kern, dim3(G), dim3(128), args, 0, 0);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    epoch += steps;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cv=%d hdr_stride=%3dB pub=%d allrows=%d G=%3d : %7.1f ns/step %s\n", CV, HS * 16, PUB, ALLROWS, G,
           ms * 1e6 / steps, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    ulonglong2 *hdr, *rows;
    long long *sink;
    cudaMalloc(&hdr, 2 * 512 * 16 * sizeof(ulonglong2));
    cudaMalloc(&rows, 2 * 512 * 64 * sizeof(ulonglong2));
    cudaMalloc(&sink, 64);
    cudaMemset(hdr, 0, 2 * 512 * 16 * sizeof(ulonglong2));
    cudaMemset(rows, 0, 2 * 512 * 64 * sizeof(ulonglong2));
    unsigned int epoch = 1;
    const int steps = 2000;
    for (int G : {1, 2, 8, 16, 32, 64, 128}) {
        run<0, 1, 0, 0>(G, steps, hdr, rows, sink, epoch);
        run<1, 1, 0, 0>(G, steps, hdr, rows, sink, epoch);
        run<0, 8, 0, 0>(G, steps, hdr, rows, sink, epoch);
        run<1, 8, 0, 0>(G, steps, hdr, rows, sink, epoch);
        run<1, 16, 0, 0>(G, steps, hdr, rows, sink, epoch);
        run<1, 8, 1, 0>(G, steps, hdr, rows, sink, epoch);
        if (G <= 32) run<1, 8, 1, 1>(G, steps, hdr, rows, sink, epoch);
        printf("\n");
    }
    return 0;
}
