// Which pipe do the packed min/max flavours run on? (run on the B200 box)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu && ./pipes2
// If HMNMX2 (f16x2 / bf16x2 min/max) issues on another pipe than VIMNMX.U16x2, a mix of both runs faster than either.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
#define U 8
#define OP2(name, ptx) __device__ __forceinline__ uint32_t name(uint32_t a, uint32_t b){ uint32_t r; asm volatile(ptx " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OP2(minu2, "min.u16x2") OP2(maxu2, "max.u16x2")
OP2(minh2, "min.f16x2") OP2(maxh2, "max.f16x2")
OP2(minb2, "min.bf16x2") OP2(maxb2, "max.bf16x2")
OP2(mins2, "min.s16x2") OP2(maxs2, "max.s16x2")
OP2(minu, "min.u32") OP2(maxu, "max.u32")
__device__ __forceinline__ uint32_t min3(uint32_t a, uint32_t b, uint32_t c){ return min(min(a, b), c); }
__device__ __forceinline__ uint32_t push_le(uint32_t m, uint32_t ev, uint32_t q) {
    uint32_t r;
    asm volatile("{\n.reg .u32 t;\nsub.cc.u32 t, %1, %2;\naddc.u32 %0, %3, %3;\n}" : "=r"(r) : "r"(q), "r"(ev), "r"(m));
    return r;
}

template <int MODE> __global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, uint32_t one) {
    uint32_t a[U], b[U];
#pragma unroll
    for (int i = 0; i < U; ++i) { a[i] = (seed * (threadIdx.x + 1) + i * 77u) & 0x3BFF3BFFu; b[i] = (seed ^ (threadIdx.x * 31u + i)) & 0x3BFF3BFFu; }
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            if (MODE == 0) { uint32_t lo = minu2(a[i], b[i]), hi = maxu2(a[i], b[i]); a[i] = lo; b[i] = hi; }
            if (MODE == 1) { uint32_t lo = minh2(a[i], b[i]), hi = maxh2(a[i], b[i]); a[i] = lo; b[i] = hi; }
            if (MODE == 2) { uint32_t lo = minb2(a[i], b[i]), hi = maxb2(a[i], b[i]); a[i] = lo; b[i] = hi; }
            if (MODE == 3) {  // half the exchanges as u16x2, half as f16x2
                if (i & 1) { uint32_t lo = minu2(a[i], b[i]), hi = maxu2(a[i], b[i]); a[i] = lo; b[i] = hi; }
                else       { uint32_t lo = minh2(a[i], b[i]), hi = maxh2(a[i], b[i]); a[i] = lo; b[i] = hi; }
            }
            if (MODE == 4) {  // min as u16x2, max as f16x2
                uint32_t lo = minu2(a[i], b[i]), hi = maxh2(a[i], b[i]); a[i] = lo; b[i] = hi;
            }
            if (MODE == 5) { a[i] = min3(a[i], b[i], a[(i + 1) % U]); }
            if (MODE == 6) { a[i] = push_le(a[i], b[i], a[(i + 1) % U]); }
            if (MODE == 7) { uint32_t lo = mins2(a[i], b[i]), hi = maxs2(a[i], b[i]); a[i] = lo; b[i] = hi; }
            if (MODE == 8) { uint32_t lo = minu(a[i], b[i]), hi = maxu(a[i], b[i]); a[i] = lo; b[i] = hi; }
            if (MODE == 9) {  // min as u16x2, max as bf16x2
                uint32_t lo = minu2(a[i], b[i]), hi = maxb2(a[i], b[i]); a[i] = lo; b[i] = hi;
            }
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
    run<0>("CE u16x2: min+max", 2);
    run<1>("CE f16x2: min+max", 2);
    run<2>("CE bf16x2: min+max", 2);
    run<3>("CE alternating u16x2 / f16x2", 2);
    run<4>("CE min.u16x2 + max.f16x2", 2);
    run<9>("CE min.u16x2 + max.bf16x2", 2);
    run<5>("min3 u32", 1);
    run<6>("push_le (sub.cc + addc)", 2);
    run<7>("CE s16x2: min+max", 2);
    run<8>("CE u32: min+max", 2);
    return 0;
}
