// detect.cu — sm_100a kernels of the detect hot path (v2: one fused persistent kernel).
//
// Replaces, per read, FromOverlap::compute_bad_part (reference src/stack.rs:61-139) fused with
// editor::type_of_read (src/editor/mod.rs:85-100). The reference sorts the intervals and sweeps them with
// a min-heap of interval ends. The device computes the same bad-region list in closed form:
//
//   B[0..k) = begins sorted ascending, E[0..k) = ends sorted ascending (two independent sorts).
//   The heap sweep pops every end <= begin before it pushes (stack.rs:72-81), so just before begin i is
//   pushed the heap holds  d_i = i - #{E <= B_i}  ends, and just before end q is popped it holds
//   f_q = #{B < E_q} - q.  With threshold c = `-c`:
//     up-crossing   U : begin i with d_i == c      <=>  E[i-c-1] <= B_i <  E[i-c]      (depth c -> c+1)
//     down-crossing D : end   q with f_q == c + 1  <=>  B[q+c]   <  E_q <= B[q+c+1]    (depth c+1 -> c)
//   (out-of-range E[-1] = 0, E[>=k] = B[>=k] = +inf). With X_i = (B_i < E[i-c]) and Y_i = (E[i-c-1] <= B_i)
//   both tests need only those two comparison vectors: U at begin i = X_i & Y_i, D at end i-c = X_i & Y_{i+1}.
//   Crossings alternate U0 D0 U1 D1 ... and the cleaned gap list of stack.rs:107-138 is
//       [(0,U0) if U0 != 0] ++ [(D_t, U_t+1)] ++ [(D_last, len) if D_last != len]
//   or [(0,len) if len != 0] when depth never exceeds c (tests/device_model.py is the executable form,
//   fuzzed against the literal heap sweep in tests/test_device_model.py).
//   Classification (editor/mod.rs:85-100): bad_len = len + sum(U) - sum(D) in wrapping u32;
//   NotCovered iff (double)bad_len / (double)len > n (same IEEE divide, tested first); else Chimeric iff
//   there is an interior gap <=> #U >= 2; else NotBad.
//
// Kernels (all integer work; no tensor cores — there is no contraction on this path):
//   plan_kernel    tile boundaries (binary search on rowptr[r] + 8r), list of big rows, zeroing.
//   big_kernel     rows with k > 256: one CTA per row, 2k event keys bitonic-sorted in shared memory (or in
//                  a global slab beyond 16384 events); results parked in a side buffer.
//   fused_kernel   persistent CTAs pull tiles of consecutive rows; TMA bulk copy of the tile's interval
//                  slab into shared memory; rows binned by size class; sub-warp groups of G = 2..16 lanes
//                  sort one row each in registers (16 keys per lane per array, blocked layout: shuffles only
//                  on the log2(G) outermost merge levels); crossings detected against a skewed shared-memory
//                  copy of E; per-tile scan + decoupled look-back gives the global bad-region offsets; gap
//                  CSR, classes and the 2-bit bitmap are written once, in final position.
#include "pileup.cuh"

namespace yb {
namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr uint32_t INF = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t classify(uint32_t bad_len, uint32_t len, uint32_t n_up, double not_cov) {
    // editor/mod.rs:88: `bad_region_len as f64 / length as f64 > not_covered` (NaN compares false)
    const double ratio = (double)bad_len / (double)len;
    if (ratio > not_cov) return 2u;  // NotCovered is tested first
    return n_up >= 2u ? 1u : 0u;     // an interior gap exists iff there are >= 2 up-crossings
}

__device__ __forceinline__ void ce(uint32_t &a, uint32_t &b) {
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const uint32_t lane = lane_id();
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, v, off);
        if (lane >= (uint32_t)off) v += o;
    }
    return v;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}

__host__ __device__ inline uint64_t next_pow2_u64(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
constexpr int E = 16;                          // keys per lane per array in the register tier
constexpr uint32_t kSmallMaxK = 256;           // register tier: rows with k <= 256 (G = 16 lanes x 16 keys)
#ifndef YB_TILE_W
#define YB_TILE_W 1024
#endif
constexpr uint32_t kTileW = YB_TILE_W;         // tile = rows whose weight rowptr[r] + 8r falls in one window
constexpr uint32_t kReadW = 8;
constexpr uint32_t kMaxTileReads = kTileW / kReadW;
constexpr uint32_t kSlabCap = kTileW + kSmallMaxK + 64;  // intervals staged per tile (+ alignment slack per run)
#ifndef YB_FUSED_WARPS
#define YB_FUSED_WARPS 2
#endif
constexpr uint32_t kFusedWarps = YB_FUSED_WARPS;
constexpr uint32_t kFusedThreads = kFusedWarps * 32;
constexpr uint32_t kBigThreads = 512;
constexpr uint32_t kBigSmemEvents = 16384;     // big_kernel: 64 KB of u32 event keys in shared memory

// scratch carve-up
struct Work {
    uint32_t *tile_first;            // n_tiles + 1
    unsigned long long *tile_status; // n_tiles: decoupled look-back (flag << 62 | value)
    uint32_t *big_list;              // rows with k > kSmallMaxK
    uint32_t *big_off;               // their offset (in pairs) into big_gaps
    uint32_t *big_cnt;               // their bad-region count
    uint8_t *big_cls;                // their class
    uint32_t *big_slot;              // n_reads: row -> index in big_list (valid for big rows only)
    uint2 *big_gaps;                 // sum over big rows of (k + 1) pairs
    uint32_t *huge_keys;             // event keys of rows beyond the shared-memory tier
    uint32_t n_tiles;
};

__host__ __device__ inline uint32_t n_tiles_of(uint32_t n_reads, uint32_t n_iv) {
    const uint64_t total = (uint64_t)n_iv + (uint64_t)kReadW * n_reads;
    return (uint32_t)(total / kTileW) + 1u;
}

// ------------------------------------------------------------------------------------------------
// plan_kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) plan_kernel(DetectArgs a, Work w) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    // tile t starts at the first row r with rowptr[r] + 8 r >= t * kTileW
    for (uint32_t t = tid; t <= w.n_tiles; t += nthr) {
        const uint64_t target = (uint64_t)t * kTileW;
        uint32_t lo = 0, hi = a.n_reads;  // first r in [0, n_reads] with weight(r) >= target
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            const uint64_t wgt = (uint64_t)__ldg(a.rowptr + mid) + (uint64_t)kReadW * mid;
            if (wgt < target) lo = mid + 1; else hi = mid;
        }
        w.tile_first[t] = t == w.n_tiles ? a.n_reads : lo;
        if (t < w.n_tiles) w.tile_status[t] = 0ull;
    }
    const uint32_t n_words = (a.n_reads + 15u) >> 4;
    for (uint32_t i = tid; i < n_words; i += nthr) reinterpret_cast<uint32_t *>(a.bitmap)[i] = 0u;
    if (a.max_k > kSmallMaxK) {
        for (uint32_t r = tid; r < a.n_reads; r += nthr) {
            const uint32_t k = __ldg(a.rowptr + r + 1) - __ldg(a.rowptr + r);
            if (k > kSmallMaxK) {
                const uint32_t j = atomicAdd(a.counters + kCntBigList, 1u);
                w.big_list[j] = r;
                w.big_off[j] = atomicAdd(a.counters + kCntBigBump, k + 1u);
                w.big_slot[r] = j;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// big_kernel: one CTA per big row, event formulation (2k keys: begin 2b+1, end 2e; ends sort first at
// equal positions, stack.rs:72-81), bitonic network in shared memory or in a global slab.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void check_interval(const uint2 v, uint32_t len, uint32_t *counters) {
    if (!(v.x < v.y && v.y <= len)) atomicAdd(counters + kCntMalformed, 1u);
}

__device__ void cta_pileup(uint32_t *ev, uint32_t n_pow2, const uint2 *__restrict__ row, uint32_t k, uint32_t len,
                           uint32_t c, double not_cov, uint32_t *__restrict__ flat, uint8_t *__restrict__ cls_out,
                           uint32_t *__restrict__ cnt_out, uint32_t *sh /* 5 * 32 u32 */, uint32_t *counters) {
    const uint32_t tid = threadIdx.x, nthr = blockDim.x, n_ev = 2u * k;
    const uint32_t lane = lane_id(), wid = tid >> 5, nwarps = nthr >> 5;
    for (uint32_t i = tid; i < n_pow2; i += nthr) {
        uint32_t kk = INF;
        if (i < n_ev) {
            const uint2 v = __ldg(row + (i >> 1));
            if (i & 1u) check_interval(v, len, counters);
            kk = (i & 1u) ? v.y * 2u : v.x * 2u + 1u;
        }
        ev[i] = kk;
    }
    __syncthreads();
    // all-ascending bitonic network over ev[0..n_pow2)
    const uint32_t half_n = n_pow2 >> 1;
    for (uint32_t size = 2; size <= n_pow2; size <<= 1) {
        const uint32_t half = size >> 1;
        for (uint32_t p = tid; p < half_n; p += nthr) {
            const uint32_t blk = p / half, o = p - blk * half;
            const uint32_t lo = blk * size + o, hi = blk * size + size - 1u - o;
            const uint32_t x = ev[lo], y = ev[hi];
            if (x > y) {
                ev[lo] = y;
                ev[hi] = x;
            }
        }
        __syncthreads();
        for (uint32_t stride = size >> 2; stride > 0; stride >>= 1) {
            for (uint32_t p = tid; p < half_n; p += nthr) {
                const uint32_t lo = 2u * stride * (p / stride) + (p % stride), hi = lo + stride;
                const uint32_t x = ev[lo], y = ev[hi];
                if (x > y) {
                    ev[lo] = y;
                    ev[hi] = x;
                }
            }
            __syncthreads();
        }
    }
    // Each warp owns a contiguous chunk of the sorted events and walks it 32 events per round;
    // depth inside a round comes from two ballots (begins, ends) and popc.
    uint32_t *sh_delta = sh, *sh_cross = sh + 32, *sh_first = sh + 64, *sh_last = sh + 96, *sh_bad = sh + 128;
    uint32_t chunk = n_pow2 / nwarps;
    if (chunk < 32u) chunk = 32u;
    const uint32_t beg = min(wid * chunk, n_ev), end = min(beg + chunk, n_ev);
    const uint32_t le = (2u << lane) - 1u, lt = (1u << lane) - 1u;
    uint32_t dsum = 0;
    for (uint32_t i0 = beg; i0 < end; i0 += 32u) {
        const uint32_t i = i0 + lane;
        const bool real = i < end;
        const uint32_t kb = real ? (ev[i] & 1u) : 0u;
        const uint32_t bm = __ballot_sync(FULL, real && kb), em = __ballot_sync(FULL, real && !kb);
        dsum += __popc(bm) - __popc(em);
    }
    if (lane == 0) sh_delta[wid] = dsum;
    __syncthreads();
    uint32_t depth0 = 0;
    for (uint32_t q = 0; q < wid; ++q) depth0 += sh_delta[q];
    const uint32_t cu = c + 1u;
    uint32_t d0 = depth0, ncross = 0, badsum = 0, firstpos = 0, lastpos = 0;
    for (uint32_t i0 = beg; i0 < end; i0 += 32u) {
        const uint32_t i = i0 + lane;
        const bool real = i < end;
        const uint32_t kk = real ? ev[i] : 0u;
        const uint32_t kb = real ? (kk & 1u) : 0u, pos = kk >> 1;
        const uint32_t bm = __ballot_sync(FULL, real && kb), em = __ballot_sync(FULL, real && !kb);
        const uint32_t depth = d0 + __popc(bm & le) - __popc(em & le);
        const bool up = real && kb && depth == cu, down = real && !kb && depth == c;
        const uint32_t xm = __ballot_sync(FULL, up || down);
        if (xm) {
            const uint32_t f = __shfl_sync(FULL, pos, __ffs(xm) - 1);
            const uint32_t l = __shfl_sync(FULL, pos, 31 - __clz(xm));
            if (ncross == 0) firstpos = f;
            lastpos = l;
            ncross += __popc(xm);
        }
        badsum += up ? pos : (down ? 0u - pos : 0u);
        d0 += __popc(bm) - __popc(em);
    }
    badsum = warp_sum(badsum);
    if (lane == 0) {
        sh_cross[wid] = ncross;
        sh_first[wid] = firstpos;
        sh_last[wid] = lastpos;
        sh_bad[wid] = badsum;
    }
    __syncthreads();
    uint32_t X = 0, xbase = 0, U0 = 0, Dl = 0, bad = 0;
    for (uint32_t q = 0; q < nwarps; ++q) {
        const uint32_t n = sh_cross[q];
        if (q == wid) xbase = X;
        if (n) {
            if (X == 0) U0 = sh_first[q];
            Dl = sh_last[q];
        }
        X += n;
        bad += sh_bad[q];
    }
    uint32_t n_gaps, h = 0;
    if (X) {
        h = U0 != 0u;
        n_gaps = (X >> 1) - 1u + h + (Dl != len ? 1u : 0u);
        uint32_t x = xbase;
        d0 = depth0;
        // crossing number x lands at flat[x + 2h - 1]
        for (uint32_t i0 = beg; i0 < end; i0 += 32u) {
            const uint32_t i = i0 + lane;
            const bool real = i < end;
            const uint32_t kk = real ? ev[i] : 0u;
            const uint32_t kb = real ? (kk & 1u) : 0u, pos = kk >> 1;
            const uint32_t bm = __ballot_sync(FULL, real && kb), em = __ballot_sync(FULL, real && !kb);
            const uint32_t depth = d0 + __popc(bm & le) - __popc(em & le);
            const bool cross = real && ((kb && depth == cu) || (!kb && depth == c));
            const uint32_t xm = __ballot_sync(FULL, cross);
            if (cross) {
                const int idx = (int)(x + __popc(xm & lt) + 2u * h) - 1;
                if (idx >= 0) flat[idx] = pos;
            }
            x += __popc(xm);
            d0 += __popc(bm) - __popc(em);
        }
        if (tid == 0) {
            if (h) flat[0] = 0u;
            if (Dl != len) flat[X + 2u * h - 1u] = len;
        }
    } else {
        n_gaps = len != 0u;
        if (tid == 0 && n_gaps) {
            flat[0] = 0u;
            flat[1] = len;
        }
    }
    if (tid == 0) {
        *cls_out = (uint8_t)classify(len + bad, len, X >> 1, not_cov);
        *cnt_out = n_gaps;
    }
}

__global__ void __launch_bounds__(kBigThreads) big_kernel(DetectArgs a, Work w, uint32_t c, double not_cov) {
    extern __shared__ uint32_t ev_smem[];
    __shared__ uint32_t sh[160];
    __shared__ uint32_t sh_off;
    const uint32_t n_big = a.counters[kCntBigList];
    for (uint32_t j = blockIdx.x; j < n_big; j += gridDim.x) {
        const uint32_t r = w.big_list[j];
        const uint32_t s = a.rowptr[r], k = a.rowptr[r + 1] - s;
        const uint32_t n_pow2 = (uint32_t)next_pow2_u64(2ull * k);
        uint32_t *ev = ev_smem;
        if (n_pow2 > kBigSmemEvents) {  // keys live in a bump-allocated global slab
            if (threadIdx.x == 0) sh_off = atomicAdd(a.counters + kCntHugeBump, n_pow2);
            __syncthreads();
            ev = w.huge_keys + sh_off;
        }
        cta_pileup(ev, n_pow2, a.iv + s, k, a.len[r], c, not_cov, reinterpret_cast<uint32_t *>(w.big_gaps + w.big_off[j]),
                   w.big_cls + j, w.big_cnt + j, sh, a.counters);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// register tier: G lanes per row (G = 2..16, per lane at run time), E = 16 keys per lane per array,
// blocked layout (element = g*16 + t): shuffles only on the log2(G) outermost merge levels.
// ------------------------------------------------------------------------------------------------
// Batcher odd-even merge sort of the 16 keys a lane holds (63 compare-exchanges, no shuffles).
__device__ __forceinline__ void sort16(uint32_t (&k)[E]) {
#define CE(i, j) ce(k[i], k[j]);
    CE(0, 1) CE(2, 3) CE(0, 2) CE(1, 3) CE(1, 2) CE(4, 5) CE(6, 7) CE(4, 6) CE(5, 7) CE(5, 6) CE(0, 4) CE(2, 6)
    CE(2, 4) CE(1, 5) CE(3, 7) CE(3, 5) CE(1, 2) CE(3, 4) CE(5, 6) CE(8, 9) CE(10, 11) CE(8, 10) CE(9, 11)
    CE(9, 10) CE(12, 13) CE(14, 15) CE(12, 14) CE(13, 15) CE(13, 14) CE(8, 12) CE(10, 14) CE(10, 12) CE(9, 13)
    CE(11, 15) CE(11, 13) CE(9, 10) CE(11, 12) CE(13, 14) CE(0, 8) CE(4, 12) CE(4, 8) CE(2, 10) CE(6, 14)
    CE(6, 10) CE(2, 4) CE(6, 8) CE(10, 12) CE(1, 9) CE(5, 13) CE(5, 9) CE(3, 11) CE(7, 15) CE(7, 11) CE(3, 5)
    CE(7, 9) CE(11, 13) CE(1, 2) CE(3, 4) CE(5, 6) CE(7, 8) CE(9, 10) CE(11, 12) CE(13, 14)
#undef CE
}

// Sorts, for every group of G lanes, its 16*G keys (ascending in element order g*16 + t). G is a per-lane
// run-time value (groups of different sizes share a warp, larger groups on lower lanes); gmax is the
// warp's largest G. One rolled loop serves every group size and both arrays, so the code stays small
// enough for the instruction caches.
__device__ __forceinline__ void sort_group(uint32_t (&key)[E], uint32_t G, uint32_t gmax) {
    sort16(key);
    const uint32_t lane = lane_id();
#pragma unroll 1
    for (uint32_t ls = 2; ls <= gmax; ls <<= 1) {
        const bool on = ls <= G;
        {   // flip: element e pairs with e ^ (16*ls - 1): partner lane ^ (ls-1), slot 15 - t
            const bool keep_min = (lane & (ls >> 1)) == 0;
            uint32_t other[E];
#pragma unroll
            for (int t = 0; t < E; ++t) other[t] = __shfl_xor_sync(FULL, key[E - 1 - t], ls - 1);
            if (on) {
#pragma unroll
                for (int t = 0; t < E; ++t) key[t] = keep_min ? min(key[t], other[t]) : max(key[t], other[t]);
            }
        }
#pragma unroll 1
        for (uint32_t j = ls >> 2; j > 0; j >>= 1) {  // half-cleaners on the lane bits
            const bool keep_min = (lane & j) == 0;
#pragma unroll
            for (int t = 0; t < E; ++t) {
                const uint32_t o = __shfl_xor_sync(FULL, key[t], j);
                if (on) key[t] = keep_min ? min(key[t], o) : max(key[t], o);
            }
        }
        if (on) {
#pragma unroll
            for (int s = E >> 1; s > 0; s >>= 1) {  // half-cleaners on the slot bits
#pragma unroll
                for (int t = 0; t < E; ++t)
                    if ((t & s) == 0) ce(key[t], key[t | s]);
            }
        }
    }
}

constexpr uint32_t kScratchWords = 17u * 33u + 3u;  // 33 blocks of 16 keys at a 17-word pitch

struct alignas(16) WarpSmem {  // one per warp: a warp runs its tiles on its own, no CTA-wide barrier anywhere
    unsigned long long mbar;
    uint32_t row[kMaxTileReads + 1];   // rowptr values of the tile's rows (row i = read r0 + i)
    uint32_t len[kMaxTileReads];
    uint32_t meta[kMaxTileReads];      // n_gaps | h << 30 | tail << 31; big rows: n_gaps
    uint32_t goff[kMaxTileReads];      // exclusive scan of n_gaps inside the tile
    uint16_t soff[kMaxTileReads];      // slab offset (in intervals) of the row's data
    uint8_t order[kMaxTileReads];      // rows grouped by size class, largest first
    uint8_t cls[kMaxTileReads];
    uint32_t scr[kScratchWords];
};
static_assert(sizeof(WarpSmem) % 16 == 0, "slab must stay 16-byte aligned");
constexpr size_t kWarpSmemBytes = sizeof(WarpSmem) + sizeof(uint2) * kSlabCap;
static_assert(kMaxTileReads <= 128 && kMaxTileReads % 32 == 0, "row ids are u8; rows are walked 32 at a time");

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// One batch of rows: lane p of the batch belongs to the group of G lanes that owns row i_row (valid lanes
// only). Sorts the row's begins and ends, finds the crossings, leaves them in the row's slab slot and the
// row's bad-region count / flags / class in ws.meta / ws.cls.
__device__ __forceinline__ void process_batch(WarpSmem &ws, uint2 *slab, bool valid, uint32_t i_row, uint32_t G, uint32_t g,
                                              uint32_t gmax, uint32_t c, double not_cov, uint32_t &malformed) {
    const uint32_t lane = lane_id();
    const uint32_t k = valid ? ws.row[i_row + 1] - ws.row[i_row] : 0u;
    const uint32_t len = valid ? ws.len[i_row] : 0u;
    const uint32_t so = valid ? ws.soff[i_row] : 0u;
    const uint2 *row = slab + so;
    // striped load (conflict-free); the initial arrangement is irrelevant to the sort
    uint32_t B[E], En[E];
    bool bad_iv = false;
#pragma unroll
    for (int t = 0; t < E; ++t) {
        const uint32_t e = (uint32_t)t * G + g;
        uint2 v = make_uint2(INF, INF);
        if (e < k) {
            v = row[e];
            bad_iv |= !(v.x < v.y && v.y <= len);
        }
        B[t] = v.x;
        En[t] = v.y;
    }
    malformed += bad_iv;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {  // rolled: one copy of the network sorts begins, then ends
        sort_group(B, G, gmax);
#pragma unroll
        for (int t = 0; t < E; ++t) {
            const uint32_t x = B[t];
            B[t] = En[t];
            En[t] = x;
        }
    }
    // skewed copy of E (17-word pitch per 16-key block, lane l's block at 17 (l + 1)): conflict-free for
    // the blocked writers and for the shifted readers
    uint32_t *q = ws.scr + 17u * (lane + 1u);
    __syncwarp();
#pragma unroll
    for (int t = 0; t < E; ++t) q[t] = En[t];
    __syncwarp();
    // Ev[t] = E[16 g + t - c - 1], t = 0..16 (0 below the row's first end)
    uint32_t Ev[E + 1];
    if (c < 16u) {
        // 18 consecutive words starting in the previous lane's block; the pitch hole sits at index c + 1
        const uint32_t *wptr = q - 2u - c;
        uint32_t W[E + 2];
#pragma unroll
        for (int t = 0; t < E + 2; ++t) W[t] = wptr[t];
#pragma unroll
        for (int t = 0; t < E + 1; ++t) Ev[t] = (uint32_t)t <= c ? (g == 0u ? 0u : W[t]) : W[t + 1];
    } else {
        const uint32_t *q0 = q - 17u * g;  // the group's first block
#pragma unroll
        for (int t = 0; t < E + 1; ++t) {
            const int e = (int)(16u * g + t) - (int)min(c, 0x7FFFFFF0u) - 1;
            Ev[t] = e < 0 ? 0u : q0[e + (e >> 4)];
        }
    }
    uint32_t Bnext = __shfl_down_sync(FULL, B[0], 1);
    if (g == G - 1u) Bnext = INF;
    // X_t = B_t < Ev[t+1], Y_t = Ev[t] <= B_t; U = X_t & Y_t, D = X_t & Y_{t+1}
    uint32_t um = 0, dm = 0;
#pragma unroll
    for (int t = 0; t < E; ++t) {
        const bool x = B[t] < Ev[t + 1];
        const bool y = Ev[t] <= B[t];
        const bool y1 = Ev[t + 1] <= (t + 1 < E ? B[(t + 1) % E] : Bnext);
        if (x && y) um |= 1u << t;
        if (x && y1) dm |= 1u << t;
    }
    // ranks of this lane's crossings among the row's ups / downs (packed segmented scan over the group)
    const uint32_t mine = __popc(um) | (__popc(dm) << 16);
    uint32_t incl = mine;
#pragma unroll 1
    for (uint32_t off = 1; off < gmax; off <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, incl, off);
        if (g >= off) incl += o;
    }
    const uint32_t tot = __shfl_sync(FULL, incl, lane | (G - 1u));
    uint32_t ru = (incl - mine) & 0xFFFFu, rd = (incl - mine) >> 16;
    const uint32_t n_up = tot & 0xFFFFu;
    // crossings go back into the row's own slab slot (2k words, no longer needed): C[2j] = U_j, C[2j+1] = D_j
    uint32_t *C = reinterpret_cast<uint32_t *>(slab + so);
    uint32_t acc = 0;
#pragma unroll
    for (int t = 0; t < E; ++t) {
        if (um & (1u << t)) {
            C[2u * ru] = B[t];
            acc += B[t];
            ++ru;
        }
        if (dm & (1u << t)) {
            C[2u * rd + 1u] = Ev[t + 1];
            acc -= Ev[t + 1];
            ++rd;
        }
    }
#pragma unroll 1
    for (uint32_t off = gmax >> 1; off > 0; off >>= 1) {
        const uint32_t o = __shfl_xor_sync(FULL, acc, off);
        if (off < G) acc += o;
    }
    __syncwarp();
    if (valid && g == 0u) {
        uint32_t ng, h, tail;
        if (n_up) {
            h = C[0] != 0u;
            tail = C[2u * n_up - 1u] != len;
            ng = n_up - 1u + h + tail;
        } else {
            ng = h = tail = len != 0u;
        }
        ws.meta[i_row] = ng | (h << 30) | (tail << 31);
        ws.cls[i_row] = (uint8_t)classify(len + acc, len, n_up, not_cov);
    }
}

__global__ void __launch_bounds__(kFusedThreads) fused_kernel(DetectArgs a, Work w, uint32_t c, double not_cov) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    WarpSmem &ws = *reinterpret_cast<WarpSmem *>(smem_raw + wid * kWarpSmemBytes);
    uint2 *slab = reinterpret_cast<uint2 *>(smem_raw + wid * kWarpSmemBytes + sizeof(WarpSmem));
    if (lane == 0) {
        mbar_init(&ws.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t parity = 0, malformed = 0, hist0 = 0, hist1 = 0, hist2 = 0;
    const uint32_t lt = (1u << lane) - 1u;
    volatile unsigned long long *st = w.tile_status;
    const uint32_t n_warps = gridDim.x * kFusedWarps;
    // Static round-robin over tiles: every tile below the one a warp works on belongs to a resident warp
    // that is at or before it in its own sequence, so the look-back below cannot deadlock.
    for (uint32_t tile = blockIdx.x * kFusedWarps + wid; tile < w.n_tiles; tile += n_warps) {
        const uint32_t r0 = __ldg(w.tile_first + tile), r1 = __ldg(w.tile_first + tile + 1), R = r1 - r0;
        if (R == 0) {  // empty window: still a link of the look-back chain
            if (lane == 0) {
                unsigned long long v = 0;
                if (tile) {
                    do v = st[tile - 1];
                    while ((v >> 62) != 2ull);
                }
                st[tile] = (2ull << 62) | (v & 0xFFFFFFFFull);
            }
            continue;
        }
#pragma unroll 1
        for (uint32_t i = lane; i <= R; i += 32u) ws.row[i] = __ldg(a.rowptr + r0 + i);
#pragma unroll 1
        for (uint32_t i = lane; i < R; i += 32u) ws.len[i] = __ldg(a.len + r0 + i);
        __syncwarp();
        // ---- rows -> size classes (G = 2, 4, 8, 16 lanes); trivial rows (k <= c: depth never exceeds c)
        //      are finished here; big rows (k > 256) were finished by big_kernel ----
        uint32_t cnt0 = 0, cnt1 = 0, cnt2 = 0, cnt3 = 0;
        uint32_t my_cls[kMaxTileReads / 32], my_rank[kMaxTileReads / 32];
        bool any_big = false;
#pragma unroll
        for (uint32_t u = 0; u < kMaxTileReads / 32; ++u) {
            const uint32_t i = u * 32u + lane;
            uint32_t cl = 0xFFu;
            if (i < R) {
                const uint32_t k = ws.row[i + 1] - ws.row[i], len = ws.len[i];
                if (k > kSmallMaxK) {
                    cl = 0xFEu;
                } else if (k <= c) {
                    const uint32_t ng = len != 0u;
                    ws.meta[i] = ng | (ng << 30) | (ng << 31);
                    ws.cls[i] = (uint8_t)classify(len, len, 0u, not_cov);
                    if (k) {  // still validate the intervals of a row that is not sorted
                        const uint2 *gi = a.iv + ws.row[i];
                        bool bad = false;
#pragma unroll 1
                        for (uint32_t j = 0; j < k; ++j) {
                            const uint2 v = __ldg(gi + j);
                            bad |= !(v.x < v.y && v.y <= len);
                        }
                        malformed += bad;
                    }
                } else {
                    cl = k <= 32u ? 0u : (k <= 64u ? 1u : (k <= 128u ? 2u : 3u));
                }
            }
            any_big |= cl == 0xFEu;
            const uint32_t m0 = __ballot_sync(FULL, cl == 0u), m1 = __ballot_sync(FULL, cl == 1u);
            const uint32_t m2 = __ballot_sync(FULL, cl == 2u), m3 = __ballot_sync(FULL, cl == 3u);
            my_cls[u] = cl;
            my_rank[u] = cl == 0u ? cnt0 + __popc(m0 & lt)
                       : cl == 1u ? cnt1 + __popc(m1 & lt)
                       : cl == 2u ? cnt2 + __popc(m2 & lt) : cnt3 + __popc(m3 & lt);
            cnt0 += __popc(m0);
            cnt1 += __popc(m1);
            cnt2 += __popc(m2);
            cnt3 += __popc(m3);
        }
        const bool tile_has_big = __any_sync(FULL, any_big);
        // row order: class 3 (G = 16) first; lane position p of a row = lane base of its class + rank * G
        const uint32_t rb3 = 0, rb2 = cnt3, rb1 = rb2 + cnt2, rb0 = rb1 + cnt1;
        const uint32_t lb3 = 0, lb2 = 16u * cnt3, lb1 = lb2 + 8u * cnt2, lb0 = lb1 + 4u * cnt1, lanes_total = lb0 + 2u * cnt0;
#pragma unroll
        for (uint32_t u = 0; u < kMaxTileReads / 32; ++u) {
            const uint32_t i = u * 32u + lane, cl = my_cls[u];
            if (cl < 4u) ws.order[(cl == 0u ? rb0 : cl == 1u ? rb1 : cl == 2u ? rb2 : rb3) + my_rank[u]] = (uint8_t)i;
            if (i < R && !tile_has_big) ws.soff[i] = (uint16_t)(ws.row[i] - (ws.row[0] & ~1u));
        }
        // ---- stage the tile's interval slab: TMA bulk copies, one per run of non-big rows ----
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (!tile_has_big) {
                const uint32_t cs = ws.row[0] & ~1u, ce_ = (ws.row[R] + 1u) & ~1u, nb = (ce_ - cs) * 8u;
                mbar_expect_tx(&ws.mbar, nb);
                if (nb) tma_load_1d(slab, a.iv + cs, nb, &ws.mbar);
            } else {
                uint32_t bytes_total = 0;
                for (int pass = 0; pass < 2; ++pass) {  // expect_tx must precede the copies
                    uint32_t d = 0, i = 0;
                    while (i < R) {
                        while (i < R && ws.row[i + 1] - ws.row[i] > kSmallMaxK) {
                            ws.soff[i] = 0;
                            ++i;
                        }
                        if (i >= R) break;
                        uint32_t j = i;
                        const uint32_t cs = ws.row[i] & ~1u;
                        while (j < R && ws.row[j + 1] - ws.row[j] <= kSmallMaxK) {
                            ws.soff[j] = (uint16_t)(d + (ws.row[j] - cs));
                            ++j;
                        }
                        const uint32_t ce_ = (ws.row[j] + 1u) & ~1u, nb = (ce_ - cs) * 8u;
                        if (pass == 0) bytes_total += nb;
                        else if (nb) tma_load_1d(slab + d, a.iv + cs, nb, &ws.mbar);
                        d += ce_ - cs;
                        i = j;
                    }
                    if (pass == 0) mbar_expect_tx(&ws.mbar, bytes_total);
                }
            }
        }
        __syncwarp();
        mbar_wait(&ws.mbar, parity);
        parity ^= 1u;
        // ---- pass A: sort + crossings, 32 lanes of rows at a time ----
        for (uint32_t p0 = 0; p0 < lanes_total; p0 += 32u) {
            const uint32_t p = p0 + lane;
            const bool valid = p < lanes_total;
            const uint32_t cl = p < lb2 ? 3u : (p < lb1 ? 2u : (p < lb0 ? 1u : 0u));
            const uint32_t G = 2u << cl;
            const uint32_t rel = p - (cl == 3u ? lb3 : cl == 2u ? lb2 : cl == 1u ? lb1 : lb0);
            const uint32_t rank = rel >> (cl + 1u), g = rel & (G - 1u);
            const uint32_t i_row = valid ? ws.order[(cl == 0u ? rb0 : cl == 1u ? rb1 : cl == 2u ? rb2 : rb3) + rank] : 0u;
            const uint32_t gmax = __shfl_sync(FULL, G, 0);  // classes are laid out largest first
            process_batch(ws, slab, valid, i_row, G, g, gmax, c, not_cov, malformed);
        }
        if (tile_has_big) {
            for (uint32_t i = lane; i < R; i += 32u)
                if (ws.row[i + 1] - ws.row[i] > kSmallMaxK) {
                    const uint32_t j = w.big_slot[r0 + i];
                    ws.meta[i] = w.big_cnt[j];
                    ws.cls[i] = w.big_cls[j];
                }
        }
        __syncwarp();
        // ---- exclusive scan of the bad-region counts over the tile's rows (row order) ----
        constexpr uint32_t kPer = kMaxTileReads / 32;
        uint32_t v[kPer], s = 0;
#pragma unroll
        for (uint32_t u = 0; u < kPer; ++u) {
            const uint32_t i = lane * kPer + u;
            v[u] = i < R ? (ws.meta[i] & 0x3FFFFFFFu) : 0u;
            s += v[u];
        }
        const uint32_t incl = warp_incl_scan(s);
        const uint32_t tile_total = __shfl_sync(FULL, incl, 31);
        uint32_t pre = incl - s;
#pragma unroll
        for (uint32_t u = 0; u < kPer; ++u) {
            const uint32_t i = lane * kPer + u;
            if (i < R) ws.goff[i] = pre;
            pre += v[u];
        }
        // ---- decoupled look-back over the tiles before this one ----
        if (lane == 0) st[tile] = ((tile ? 1ull : 2ull) << 62) | tile_total;
        uint32_t excl = 0;
        if (tile) {
            int look = (int)tile - 1;
            for (;;) {
                const int idx = look - (int)lane;
                unsigned long long sv = (2ull << 62);
                if (idx >= 0) sv = st[idx];
                const uint32_t flag = (uint32_t)(sv >> 62);
                const uint32_t inval = __ballot_sync(FULL, flag == 0u);
                const uint32_t incl_m = __ballot_sync(FULL, flag == 2u);
                const uint32_t upto = incl_m ? ((2u << (__ffs(incl_m) - 1)) - 1u) : FULL;
                if (inval & upto) continue;  // a needed predecessor has not published yet
                excl += warp_sum(((1u << lane) & upto) ? (uint32_t)sv : 0u);
                if (incl_m) break;
                look -= 32;
            }
            if (lane == 0) st[tile] = (2ull << 62) | (unsigned long long)(excl + tile_total);
        }
        __syncwarp();
        // ---- pass B: results to HBM, once, in final position ----
        for (uint32_t i = lane; i < R; i += 32u) {
            const uint32_t m = ws.meta[i], len = ws.len[i], k = ws.row[i + 1] - ws.row[i];
            const uint32_t base = excl + ws.goff[i], cl = ws.cls[i];
            a.gap_ptr[r0 + i] = base;
            a.cls[r0 + i] = (uint8_t)cl;
            hist0 += cl == 0u;
            hist1 += cl == 1u;
            hist2 += cl == 2u;
            if (k > kSmallMaxK) {
                const uint32_t j = w.big_slot[r0 + i];
                const uint2 *src = w.big_gaps + w.big_off[j];
                for (uint32_t gq = 0; gq < m; ++gq) a.gaps[base + gq] = src[gq];
            } else {
                const uint32_t ng = m & 0x3FFFFFFFu, h = (m >> 30) & 1u, tail = m >> 31;
                const uint32_t *C = reinterpret_cast<const uint32_t *>(slab + ws.soff[i]);
                for (uint32_t gq = 0; gq < ng; ++gq) {
                    const uint32_t f0 = 2u * gq, f1 = f0 + 1u;
                    uint2 o;
                    o.x = (f0 == 0u && h) ? 0u : C[f0 + 1u - 2u * h];
                    o.y = (f1 == 2u * ng - 1u && tail) ? len : C[f1 + 1u - 2u * h];
                    a.gaps[base + gq] = o;
                }
            }
        }
        // 2-bit bitmap: word j covers reads 16 j .. 16 j + 15; words shared with a neighbour tile are OR-ed
        {
            const uint32_t w0 = r0 >> 4, w1 = (r1 - 1u) >> 4;
            for (uint32_t wj = w0 + lane; wj <= w1; wj += 32u) {
                const uint32_t lo = max(wj << 4, r0), hi = min((wj << 4) + 16u, r1);
                uint32_t bits = 0;
#pragma unroll 1
                for (uint32_t r = lo; r < hi; ++r) bits |= (uint32_t)ws.cls[r - r0] << (2u * (r & 15u));
                uint32_t *dst = reinterpret_cast<uint32_t *>(a.bitmap) + wj;
                if (hi - lo == 16u) *dst = bits;
                else if (bits) atomicOr(dst, bits);
            }
        }
        if (r1 == a.n_reads && lane == 0) a.gap_ptr[a.n_reads] = excl + tile_total;
        __syncwarp();
    }
    hist0 = warp_sum(hist0);
    hist1 = warp_sum(hist1);
    hist2 = warp_sum(hist2);
    malformed = warp_sum(malformed);
    if (lane == 0) {
        if (hist0) atomicAdd(a.counters + kCntNotBad, hist0);
        if (hist1) atomicAdd(a.counters + kCntChimeric, hist1);
        if (hist2) atomicAdd(a.counters + kCntNotCovered, hist2);
        if (malformed) atomicAdd(a.counters + kCntMalformed, malformed);
    }
}

constexpr size_t kFusedSmemBytes = kWarpSmemBytes * kFusedWarps;

// FromReport path: bad regions are given, only type_of_read (editor/mod.rs:85-100) runs. One thread
// takes 16 consecutive reads so it owns one 32-bit word of the 2-bit bitmap.
__global__ void __launch_bounds__(256) classify_kernel(const uint32_t *__restrict__ len, const uint32_t *__restrict__ gap_ptr,
                                                        const uint2 *__restrict__ gaps, uint32_t n, double not_cov,
                                                        uint8_t *__restrict__ cls, uint8_t *__restrict__ bitmap,
                                                        uint32_t *counters) {
    const uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) * 16u;
    if (base >= n) return;
    uint32_t bits = 0;
    for (uint32_t i = 0; i < 16u && base + i < n; ++i) {
        const uint32_t r = base + i, l = len[r];
        uint32_t bad = 0, interior = 0;
        for (uint32_t g = gap_ptr[r]; g < gap_ptr[r + 1]; ++g) {
            const uint2 v = gaps[g];
            bad += v.y - v.x;
            interior |= (v.x != 0u && v.y != l) ? 1u : 0u;
        }
        const uint32_t cl = classify(bad, l, interior ? 2u : 0u, not_cov);
        cls[r] = (uint8_t)cl;
        bits |= cl << (2u * i);
        atomicAdd(counters + kCntNotBad + cl, 1u);
    }
    reinterpret_cast<uint32_t *>(bitmap)[base >> 4] = bits;
}

Work carve(const DetectArgs &a, uint64_t huge_keys, uint64_t n_big, uint64_t big_pairs, size_t *total) {
    Work w;
    size_t off = 0;
    char *base = static_cast<char *>(a.scratch);
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off += align256(bytes);
        return p;
    };
    w.n_tiles = n_tiles_of(a.n_reads, a.n_iv);
    w.tile_first = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * ((size_t)w.n_tiles + 2)));
    w.tile_status = reinterpret_cast<unsigned long long *>(take(sizeof(unsigned long long) * ((size_t)w.n_tiles + 1)));
    w.big_list = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * (n_big + 1)));
    w.big_off = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * (n_big + 1)));
    w.big_cnt = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * (n_big + 1)));
    w.big_cls = reinterpret_cast<uint8_t *>(take(n_big + 1));
    w.big_slot = reinterpret_cast<uint32_t *>(take(n_big ? sizeof(uint32_t) * ((size_t)a.n_reads + 1) : 4));
    w.big_gaps = reinterpret_cast<uint2 *>(take(sizeof(uint2) * (big_pairs + 1)));
    w.huge_keys = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * (huge_keys + 1)));
    *total = off;
    return w;
}

}  // namespace

uint64_t huge_keys_for_row(uint64_t k) {
    const uint64_t p = next_pow2_u64(2 * k);
    return (k > kSmallMaxK && p > kBigSmemEvents) ? p : 0;
}
uint64_t big_pairs_for_row(uint64_t k) { return k > kSmallMaxK ? k + 1 : 0; }

size_t detect_scratch_bytes(uint32_t n_reads, uint32_t n_iv, const RowStats &rs) {
    DetectArgs a{};
    a.n_reads = n_reads;
    a.n_iv = n_iv;
    size_t total = 0;
    carve(a, rs.huge_keys, rs.n_big, rs.big_pairs, &total);
    return total;
}

int launch_detect(const DetectArgs &a, uint32_t coverage, double not_coverage, cudaStream_t stream) {
    static int n_sm = 0, fused_occ = 0;
    if (!n_sm) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return -1;
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kBigSmemEvents * sizeof(uint32_t))) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemBytes) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fused_occ, fused_kernel, kFusedThreads, kFusedSmemBytes) != cudaSuccess) return -1;
        if (fused_occ < 1) return -1;
    }
    int launches = 0;
    if (cudaMemsetAsync(a.counters, 0, kNumCounters * sizeof(uint32_t), stream) != cudaSuccess) return -1;
    if (a.n_reads == 0) {
        if (cudaMemsetAsync(a.gap_ptr, 0, sizeof(uint32_t), stream) != cudaSuccess) return -1;
        return 0;
    }
    size_t total = 0;
    Work w = carve(a, a.rows.huge_keys, a.rows.n_big, a.rows.big_pairs, &total);
    if (total > a.scratch_bytes) return -1;
    const uint32_t c = coverage;
    const uint32_t plan_items = a.rows.n_big ? a.n_reads : (w.n_tiles > ((a.n_reads + 15u) >> 4) ? w.n_tiles : ((a.n_reads + 15u) >> 4));
    uint32_t plan_blocks = (plan_items + 255u) / 256u;
    if (plan_blocks > (uint32_t)n_sm * 8u) plan_blocks = (uint32_t)n_sm * 8u;
    if (plan_blocks == 0) plan_blocks = 1;
    plan_kernel<<<plan_blocks, 256, 0, stream>>>(a, w);
    ++launches;
    if (a.rows.n_big) {
        uint32_t grid = (uint32_t)n_sm * 2u;
        if (grid > a.rows.n_big) grid = (uint32_t)a.rows.n_big;
        big_kernel<<<grid, kBigThreads, kBigSmemEvents * sizeof(uint32_t), stream>>>(a, w, c, not_coverage);
        ++launches;
    }
    uint32_t grid = (uint32_t)(n_sm * fused_occ);
    if (grid > w.n_tiles) grid = w.n_tiles;
    fused_kernel<<<grid, kFusedThreads, kFusedSmemBytes, stream>>>(a, w, c, not_coverage);
    ++launches;
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

int launch_classify(const uint32_t *len, const uint32_t *gap_ptr, const uint2 *gaps, uint32_t n_reads, double not_coverage,
                    uint8_t *cls, uint8_t *bitmap, uint32_t *counters, cudaStream_t stream) {
    if (cudaMemsetAsync(counters, 0, kNumCounters * sizeof(uint32_t), stream) != cudaSuccess) return -1;
    if (n_reads == 0) return 0;
    const uint32_t threads = (n_reads + 15) / 16;
    classify_kernel<<<(threads + 255) / 256, 256, 0, stream>>>(len, gap_ptr, gaps, n_reads, not_coverage, cls, bitmap, counters);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace yb
