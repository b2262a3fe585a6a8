// How large can straight-line code be before instruction supply caps the issue rate? (run on the B200 box)
// A kernel whose loop body is NI independent-enough VIMNMX.U16x2 instructions over 32 registers (compare-exchange pairs in
// a fixed pattern, like a sorting network), executed by W warps per SM. Reports warp-instructions per clock per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu && ./icache
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int BASE> __device__ __forceinline__ void block64(uint32_t (&k)[32]) {
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        const int I = BASE + i, a = (I * 7) & 31, b = (a + 1 + ((I * 5) & 15)) & 31;
        const uint32_t lo = __vminu2(k[a], k[b]), hi = __vmaxu2(k[a], k[b]);
        k[a] = lo;
        k[b] = hi;
    }
}
#define B4(n) block64<(n)>(k); block64<(n) + 64>(k); block64<(n) + 128>(k); block64<(n) + 192>(k);

template <int NCE> __global__ void __launch_bounds__(32) body(uint32_t *out, uint32_t seed, int iters, int skew) {
    uint32_t k[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) k[i] = seed * (threadIdx.x + 1) + i * 2654435761u;
    if (skew) {  // take the warps of an SM out of step: a different delay per resident warp
        const unsigned slot = blockIdx.x / 148u;
        __nanosleep(slot * (unsigned)skew);
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (NCE >= 64) block64<0>(k);
        if (NCE >= 128) block64<64>(k);
        if (NCE >= 256) { block64<128>(k); block64<192>(k); }
        if (NCE >= 512) { B4(256) }
        if (NCE >= 768) { B4(512) }
        if (NCE >= 1024) { B4(768) }
        if (NCE >= 1536) { B4(1024) B4(1280) }
        if (NCE >= 2048) { B4(1536) B4(1792) }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s ^= k[i];
    out[blockIdx.x * 32 + threadIdx.x] = s;
}

template <int NCE> void run(int warps_per_sm, int skew) {
    int n_sm, clk;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    uint32_t *out;
    const int blocks = n_sm * warps_per_sm;
    cudaMalloc(&out, blocks * 32 * 4);
    const int iters = 4000000 / NCE;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    body<NCE><<<blocks, 32>>>(out, 12345u, iters, skew);
    cudaEventRecord(e0);
    body<NCE><<<blocks, 32>>>(out, 12345u, iters, skew);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)blocks * iters * NCE * 2.0, cyc = ms * 1e-3 * clk * 1e3;
    printf("body %5d instr (%3d KB)  %2d warps/SM  skew %4d ns  %.2f warp-inst/clk/SM  (ALU-bound limit 2.0)\n", NCE * 2, NCE * 2 * 16 / 1024,
           warps_per_sm, skew, inst / cyc / n_sm);
    cudaFree(out);
}

int main() {
    for (int skew : {0, 137, 1013}) {
        for (int w : {8, 20}) {
            run<256>(w, skew);
            run<512>(w, skew);
            run<768>(w, skew);
            run<1024>(w, skew);
            run<1536>(w, skew);
            run<2048>(w, skew);
        }
    }
    return 0;
}
