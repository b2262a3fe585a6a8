// Throughput microbenchmarks for the instructions the sort network is made of (run on the B200 box).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
#define U 8
__device__ __forceinline__ uint32_t min2(uint32_t a, uint32_t b){ uint32_t r; asm volatile("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b){ uint32_t r; asm volatile("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t minu(uint32_t a, uint32_t b){ uint32_t r; asm volatile("min.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t maxu(uint32_t a, uint32_t b){ uint32_t r; asm volatile("max.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t madlo(uint32_t a, uint32_t b, uint32_t c){ uint32_t r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

template <int MODE> __global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, uint32_t one) {
    uint32_t a[U], b[U];
#pragma unroll
    for (int i = 0; i < U; ++i) { a[i] = seed * (threadIdx.x + 1) + i * 77u; b[i] = seed ^ (threadIdx.x * 31u + i); }
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            if (MODE == 0) { uint32_t lo = minu(a[i], b[i]), hi = maxu(a[i], b[i]); a[i] = lo; b[i] = hi; }            // 2 ALU
            if (MODE == 1) { uint32_t lo = min2(a[i], b[i]), hi = max2(a[i], b[i]); a[i] = lo; b[i] = hi; }            // 2 ALU (u16x2)
            if (MODE == 2) { a[i] = __shfl_xor_sync(0xFFFFFFFFu, a[i], 1); }                                           // 1 SHFL
            if (MODE == 3) { uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, a[i], 1); a[i] = minu(a[i], o); }              // SHFL + ALU
            if (MODE == 4) { uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, a[i], 1); uint32_t lo = minu(a[i], o), hi = maxu(a[i], o); a[i] = (threadIdx.x & 1) ? hi : lo; } // SHFL + 3 ALU (v3 pattern)
            if (MODE == 5) { uint32_t lo = minu(a[i], b[i]); uint32_t s = madlo(a[i], one, b[i]); b[i] = madlo(lo, 0u - one, s); a[i] = lo; } // 1 ALU + 2 FMA-pipe
            if (MODE == 6) { a[i] = madlo(a[i], one, b[i]); }                                                         // 1 IMAD
            if (MODE == 7) { uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, a[i], 1); bool p = (a[i] < o) != ((threadIdx.x & 1) != 0); a[i] = p ? a[i] : o; } // SHFL + ISETP + SEL
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < U; ++i) s += a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, double ops_per_it) {
    uint32_t *out; int n_sm; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = n_sm * 8;
    cudaMalloc(&out, blocks * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 12345u, 1u);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 12345u, 1u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inst = (double)blocks * 8 * N_IT * U * ops_per_it;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-44s %8.3f ms  %6.3f warp-inst/clk/SMSP (at %d MHz nominal)  [%g inst/it]\n", name, ms, warp_inst / cyc / (n_sm * 4), clk / 1000, ops_per_it);
    cudaFree(out);
}
int main() {
    run<0>("CE u32: min+max", 2);
    run<1>("CE u16x2: min2+max2", 2);
    run<2>("SHFL.BFLY", 1);
    run<3>("SHFL + min", 2);
    run<4>("SHFL + min + max + sel (v3 exchange)", 4);
    run<5>("CE as min + 2 IMAD", 3);
    run<6>("IMAD", 1);
    run<7>("SHFL + ISETP + SEL", 3);
    return 0;
}
