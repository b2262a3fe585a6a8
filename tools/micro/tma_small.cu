// How fast can an SM pull many SMALL row slabs into shared memory? (decides per-row TMA vs cp.async staging)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_small tma_small.cu && ./tma_small
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// MODE 0: per batch, `rows` lanes each issue one TMA copy of `bytes` (rows*bytes = 4096) into the warp's slab; wait; touch.
// MODE 1: per batch, every lane issues 16 cp.async of 8 B (striped), wait_group, reads back its own elements.
// MODE 2: per batch, one TMA copy of 4096 contiguous bytes (what the tile design does).
template <int MODE> __global__ void __launch_bounds__(256) k(const uint2 *src, size_t n_iv, uint32_t *out, int iters, int rows, int warps_total) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + wid * (4096 + 64 * 16 + 16));
    uint2 *slab = reinterpret_cast<uint2 *>(bar + 2);
    if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + wid;
    const uint32_t per_row = 512 / rows;  // intervals per row
    uint32_t acc = 0, parity = 0;
    for (int it = 0; it < iters; ++it) {
        // batch b covers 512 intervals; rows are spread pseudo-randomly so that nothing is contiguous across rows
        const size_t b = ((size_t)it * warps_total + gw);
        if (MODE == 0) {
            if (lane == 0) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); mbar_expect_tx(bar, 4096); }
            __syncwarp();
            if (lane < (uint32_t)rows) {
                const size_t r = (b * rows + lane) * 2654435761ull % (n_iv / per_row);
                tma_load_1d(slab + lane * per_row, src + r * per_row, per_row * 8, bar);
            }
            mbar_wait(bar, parity); parity ^= 1;
#pragma unroll
            for (int t = 0; t < 16; ++t) { const uint2 v = slab[t * 32 + lane]; acc += v.x ^ v.y; }
        } else if (MODE == 1) {
            const uint32_t G = 32 / rows, j = lane / G, g = lane % G;
            const size_t r = (b * rows + j) * 2654435761ull % (n_iv / per_row);
            const uint2 *rp = src + r * per_row;
#pragma unroll
            for (int t = 0; t < 16; ++t)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(slab + t * 32 + lane)), "l"(rp + t * G + g) : "memory");
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int t = 0; t < 16; ++t) { const uint2 v = slab[t * 32 + lane]; acc += v.x ^ v.y; }
        } else {
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); mbar_expect_tx(bar, 4096);
                const size_t r = b * 2654435761ull % (n_iv / 512);
                tma_load_1d(slab, src + r * 512, 4096, bar);
            }
            mbar_wait(bar, parity); parity ^= 1;
#pragma unroll
            for (int t = 0; t < 16; ++t) { const uint2 v = slab[t * 32 + lane]; acc += v.x ^ v.y; }
        }
        __syncwarp();
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char *name, const uint2 *src, size_t n_iv, uint32_t *out, int rows) {
    int n_sm; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    const int warps_per_block = 8, blocks_per_sm = 4, blocks = n_sm * blocks_per_sm, iters = 200;
    const size_t smem = warps_per_block * (4096 + 64 * 16 + 16);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256, smem>>>(src, n_iv, out, iters, rows, blocks * warps_per_block);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256, smem>>>(src, n_iv, out, iters, rows, blocks * warps_per_block);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double batches = (double)blocks * warps_per_block * iters;
    printf("%-28s rows/batch %2d: %7.3f ms  %6.1f GB/s  %7.1f M rows/s  %6.1f ns/batch/SM  err=%s\n", name, rows, ms, batches * 4096 / ms / 1e6,
           batches * rows / ms / 1e3, ms * 1e6 / (batches / n_sm), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const size_t n_iv = 128u << 20;  // 1 GiB of intervals
    uint2 *src; uint32_t *out;
    cudaMalloc(&src, n_iv * 8); cudaMemset(src, 1, n_iv * 8); cudaMalloc(&out, 4 << 20);
    for (int rows : {1, 2, 4, 8, 16, 32}) run<0>("TMA per row", src, n_iv, out, rows);
    for (int rows : {2, 4, 8, 16, 32}) run<1>("cp.async 8B striped", src, n_iv, out, rows);
    run<2>("TMA 4 KB contiguous", src, n_iv, out, 1);
    return 0;
}
