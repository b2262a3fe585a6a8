// detect.cu — sm_100a kernels of the detect hot path (v4).
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
//   (out-of-range E[-1] = 0, E[>=k] = B[>=k] = +inf). With V1_i = (E[i-c-1] <= B_i) and V0_i = (E[i-c] <= B_i)
//   both tests need only those two comparison vectors: U at begin i = V1_i & !V0_i, D at end i-c =
//   !V0_i & V1_{i+1}. Crossings alternate U0 D0 U1 D1 ... and the cleaned gap list of stack.rs:107-138 is
//       [(0,U0) if U0 != 0] ++ [(D_t, U_t+1)] ++ [(D_last, len) if D_last != len]
//   or [(0,len) if len != 0] when depth never exceeds c (tests/device_model.py is the executable form,
//   fuzzed against the literal heap sweep in tests/test_device_model.py).
//   Classification (editor/mod.rs:85-100): bad_len = len + sum(U) - sum(D) in wrapping u32;
//   NotCovered iff (double)bad_len / (double)len > n (same IEEE divide, tested first); else Chimeric iff
//   there is an interior gap <=> #U >= 2; else NotBad.
//
// Kernels (all integer work; no tensor cores — there is no contraction on this path):
//   plan_kernel      tile descriptors (binary search on rowptr[r] + 8r), list of big rows.
//   big_kernel       rows with k > 256: one CTA per row, 2k event keys bitonic-sorted in shared memory (or in
//                    a global slab beyond 16384 events); results parked in a side buffer.
//   fused_kernel<1>  the fast pass. Warps pull tiles of consecutive rows from an atomic counter; one TMA
//                    bulk copy (cp.async.bulk, SASS UBLKCP) stages the tile's interval slab in shared memory
//                    while the rows are binned by size; sub-warp groups of G = 1..16 lanes sort one row each
//                    in registers as PACKED u16x2 keys (begin | end << 16): one VIMNMX.U16x2 moves a begin
//                    and an end through the same network, so both sorts cost one; crossings come from two
//                    u32 compares per slot against a skewed shared-memory copy; the tile's bad regions go to
//                    a bump-allocated staging segment (one atomic per tile, no inter-tile waiting).
//   fused_kernel<0>  the same pass with two u32 key arrays, for the tiles that hold a read longer than
//                    65534 bases or a big row.
//   scan_tiles_kernel / finalize_kernel / bitmap_kernel
//                    exclusive scan of the per-tile totals, segment copy staging -> ordered bad-region CSR
//                    (+ gap_ptr fix-up), 2-bit class bitmap and class histogram.
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
constexpr uint32_t kSlowFlag = 0x80000000u;    // tile_desc.y: the tile goes to the generic (u32) pass

// scratch carve-up
struct Work {
    uint4 *tile_desc;                // n_tiles + 1: {first row, rows | kSlowFlag, slab start (even), slab intervals (even)}
    uint32_t *tile_base;             // n_tiles: where the tile's bad regions sit in `stage` (pairs)
    uint32_t *tile_total;            // n_tiles: how many
    uint32_t *tile_off;              // n_tiles + 1: exclusive scan of tile_total
    uint2 *stage;                    // n_iv + n_reads pairs: bad regions in tile-completion order
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
__device__ __forceinline__ uint32_t tile_boundary(const DetectArgs &a, uint32_t t) {
    // first row r in [0, n_reads] with rowptr[r] + 8 r >= t * kTileW
    const uint64_t target = (uint64_t)t * kTileW;
    uint32_t lo = 0, hi = a.n_reads;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const uint64_t wgt = (uint64_t)__ldg(a.rowptr + mid) + (uint64_t)kReadW * mid;
        if (wgt < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) plan_kernel(DetectArgs a, Work w, int check_rows) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (uint32_t t = tid; t < w.n_tiles; t += nthr) {
        const uint32_t r0 = tile_boundary(a, t), r1 = t + 1 == w.n_tiles ? a.n_reads : tile_boundary(a, t + 1);
        const uint32_t p0 = __ldg(a.rowptr + r0), p1 = __ldg(a.rowptr + r1);
        const uint32_t cs = p0 & ~1u, ce_ = (p1 + 1u) & ~1u;
        uint32_t slow = ce_ - cs > kSlabCap ? kSlowFlag : 0u;
        if (check_rows) {  // a read too long for 16-bit positions, or a big row, sends the tile to the generic pass
            for (uint32_t r = r0; r < r1 && !slow; ++r) {
                const uint32_t k = __ldg(a.rowptr + r + 1) - __ldg(a.rowptr + r);
                if (k > kSmallMaxK || __ldg(a.len + r) > kPackedMaxLen) slow = kSlowFlag;
            }
        }
        w.tile_desc[t] = make_uint4(r0, (r1 - r0) | slow, cs, ce_ - cs);
    }
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
// register tier: G lanes per row (G = 1..16, per lane at run time), E = 16 keys per lane,
// blocked layout (element = g*16 + t): shuffles only on the log2(G) outermost merge levels.
// PK: a key register holds begin | end << 16 and the network runs on both halves at once (VIMNMX.U16x2).
// ------------------------------------------------------------------------------------------------
template <bool PK> __device__ __forceinline__ uint32_t kmin(uint32_t a, uint32_t b) { return PK ? __vminu2(a, b) : min(a, b); }
template <bool PK> __device__ __forceinline__ uint32_t kmax(uint32_t a, uint32_t b) { return PK ? __vmaxu2(a, b) : max(a, b); }
template <bool PK> __device__ __forceinline__ void ce(uint32_t &a, uint32_t &b) {
    const uint32_t lo = kmin<PK>(a, b), hi = kmax<PK>(a, b);
    a = lo;
    b = hi;
}

// Batcher odd-even merge sort of the 16 keys a lane holds (63 compare-exchanges, no shuffles).
template <bool PK> __device__ __forceinline__ void sort16(uint32_t (&k)[E]) {
#define CE(i, j) ce<PK>(k[i], k[j]);
    CE(0, 1) CE(2, 3) CE(0, 2) CE(1, 3) CE(1, 2) CE(4, 5) CE(6, 7) CE(4, 6) CE(5, 7) CE(5, 6) CE(0, 4) CE(2, 6)
    CE(2, 4) CE(1, 5) CE(3, 7) CE(3, 5) CE(1, 2) CE(3, 4) CE(5, 6) CE(8, 9) CE(10, 11) CE(8, 10) CE(9, 11)
    CE(9, 10) CE(12, 13) CE(14, 15) CE(12, 14) CE(13, 15) CE(13, 14) CE(8, 12) CE(10, 14) CE(10, 12) CE(9, 13)
    CE(11, 15) CE(11, 13) CE(9, 10) CE(11, 12) CE(13, 14) CE(0, 8) CE(4, 12) CE(4, 8) CE(2, 10) CE(6, 14)
    CE(6, 10) CE(2, 4) CE(6, 8) CE(10, 12) CE(1, 9) CE(5, 13) CE(5, 9) CE(3, 11) CE(7, 15) CE(7, 11) CE(3, 5)
    CE(7, 9) CE(11, 13) CE(1, 2) CE(3, 4) CE(5, 6) CE(7, 8) CE(9, 10) CE(11, 12) CE(13, 14)
#undef CE
}

// Sorts, for every group of G lanes, its 16*G keys (ascending in element order g*16 + t). G is a per-lane
// run-time power of two (groups of different sizes share a warp, larger groups on lower lanes, every group
// aligned to its size); gmax is the warp's largest G. One rolled loop serves every group size. A lane whose
// group is smaller than the current level keeps its (sorted) keys: its exchanges are predicated off and the
// in-lane half-cleaners leave a sorted sequence as it is.
template <bool PK> __device__ __forceinline__ void sort_group(uint32_t (&key)[E], uint32_t G, uint32_t gmax) {
    sort16<PK>(key);
    const uint32_t lane = lane_id();
#pragma unroll 1
    for (uint32_t ls = 2; ls <= gmax; ls <<= 1) {
        const bool on = ls <= G;
        {   // flip: element e pairs with e ^ (16*ls - 1): partner lane ^ (ls-1), slot 15 - t
            const bool lo_half = (lane & (ls >> 1)) == 0;
            const bool pmin = on && lo_half, pmax = on && !lo_half;
            uint32_t other[E];
#pragma unroll
            for (int t = 0; t < E; ++t) other[t] = __shfl_xor_sync(FULL, key[E - 1 - t], ls - 1);
#pragma unroll
            for (int t = 0; t < E; ++t) {
                if (pmin) key[t] = kmin<PK>(key[t], other[t]);
                if (pmax) key[t] = kmax<PK>(key[t], other[t]);
            }
        }
#pragma unroll 1
        for (uint32_t j = ls >> 2; j > 0; j >>= 1) {  // half-cleaners on the lane bits
            const bool lo_half = (lane & j) == 0;
            const bool pmin = on && lo_half, pmax = on && !lo_half;
#pragma unroll
            for (int t = 0; t < E; ++t) {
                const uint32_t o = __shfl_xor_sync(FULL, key[t], j);
                if (pmin) key[t] = kmin<PK>(key[t], o);
                if (pmax) key[t] = kmax<PK>(key[t], o);
            }
        }
#pragma unroll
        for (int s = E >> 1; s > 0; s >>= 1) {  // half-cleaners on the slot bits
#pragma unroll
            for (int t = 0; t < E; ++t)
                if ((t & s) == 0) ce<PK>(key[t], key[t | s]);
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
    uint16_t nup[kMaxTileReads];       // up-crossings of the row; kRowDone: finished without the sort
    uint8_t order[kMaxTileReads];      // rows grouped by size class, largest first
    uint8_t cls[kMaxTileReads];
    uint32_t scr[kScratchWords];
};
constexpr uint16_t kRowDone = 0xFFFFu;
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
// TMA prefetch of a global range into L2 (the next tile's slab, while this tile is being sorted).
__device__ __forceinline__ void tma_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// One batch of rows: lane p of the batch belongs to the group of G lanes that owns row i_row (valid lanes
// only). Sorts the row's begins and ends, finds the crossings and leaves them, U0 D0 U1 D1 ..., in the row's
// own slab slot; the row's up-crossing count goes to ws.nup.
template <bool PK>
__device__ __forceinline__ void process_batch(WarpSmem &ws, uint2 *slab, bool valid, uint32_t i_row, uint32_t G, uint32_t g,
                                              uint32_t gmax, uint32_t c, uint32_t &malformed) {
    const uint32_t lane = lane_id();
    const uint32_t k = valid ? ws.row[i_row + 1] - ws.row[i_row] : 0u;
    const uint32_t len = valid ? ws.len[i_row] : 0u;
    const uint32_t so = valid ? ws.soff[i_row] : 0u;
    const uint2 *row = slab + so;
    // striped load (conflict-free); the initial arrangement is irrelevant to the sort
    uint32_t K0[E];             // PK: begin | end << 16; else begins
    uint32_t K1[PK ? 1 : E];    // else ends
    bool bad_iv = false;
#pragma unroll
    for (int t = 0; t < E; ++t) {
        const uint32_t e = (uint32_t)t * G + g;
        uint2 v = make_uint2(INF, INF);
        if (e < k) {
            v = row[e];
            bad_iv |= !(v.x < v.y && v.y <= len);
        }
        if (PK) {
            K0[t] = __byte_perm(v.x, v.y, 0x5410);
        } else {
            K0[t] = v.x;
            K1[PK ? 0 : t] = v.y;
        }
    }
    malformed += bad_iv;
    if (PK) {
        sort_group<PK>(K0, G, gmax);
    } else {
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {  // rolled: one copy of the network sorts begins, then ends
            sort_group<PK>(K0, G, gmax);
#pragma unroll
            for (int t = 0; t < E; ++t) {
                const uint32_t x = K0[t];
                K0[t] = K1[PK ? 0 : t];
                K1[PK ? 0 : t] = x;
            }
        }
    }
    // skewed copy of the sorted ends (PK: of the packed keys; the compares below only look at the end half):
    // 17-word pitch per 16-key block, lane l's block at 17 (l + 1): conflict-free for the blocked writers and
    // for the shifted readers
    uint32_t *q = ws.scr + 17u * (lane + 1u);
    __syncwarp();
#pragma unroll
    for (int t = 0; t < E; ++t) q[t] = PK ? K0[t] : K1[PK ? 0 : t];
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
    uint32_t Knext = __shfl_down_sync(FULL, K0[0], 1);
    if (g == G - 1u) Knext = INF;
    // PK: (end_j <= begin_i)  <=>  key_j <= (begin_i << 16 | 0xFFFF) as plain u32
    uint32_t um = 0, dm = 0;
    {
        bool v1 = Ev[0] <= (PK ? __byte_perm(K0[0], FULL, 0x1044) : K0[0]);
#pragma unroll
        for (int t = 0; t < E; ++t) {
            const uint32_t qt = PK ? __byte_perm(K0[t], FULL, 0x1044) : K0[t];
            const uint32_t kn = t + 1 < E ? K0[(t + 1) % E] : Knext;
            const uint32_t qn = PK ? __byte_perm(kn, FULL, 0x1044) : kn;
            const bool v0 = Ev[t + 1] <= qt;
            const bool v1n = Ev[t + 1] <= qn;
            if (v1 && !v0) um |= 1u << t;
            if (!v0 && v1n) dm |= 1u << t;
            v1 = v1n;
        }
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
    // crossings go back into the row's own slab slot (2k words, no longer needed): C[2j] = U_j, C[2j+1] = D_j
    uint32_t *C = reinterpret_cast<uint32_t *>(slab + so);
    if (um | dm) {
#pragma unroll
        for (int t = 0; t < E; ++t) {
            if (um & (1u << t)) {
                C[2u * ru] = PK ? (K0[t] & 0xFFFFu) : K0[t];
                ++ru;
            }
            if (dm & (1u << t)) {
                C[2u * rd + 1u] = PK ? (Ev[t + 1] >> 16) : Ev[t + 1];
                ++rd;
            }
        }
    }
    if (valid && g == 0u) ws.nup[i_row] = (uint16_t)(tot & 0xFFFFu);
}

// Everything a warp does for one tile. PK tiles hold only rows with k <= 256 and len <= kPackedMaxLen and
// their slab is one contiguous range (plan_kernel checked), so the TMA copy is issued before the rows are
// even looked at; the generic pass splits the slab around big rows first.
template <bool PK>
__device__ __forceinline__ void process_tile(const DetectArgs &a, const Work &w, WarpSmem &ws, uint2 *slab, uint32_t tile,
                                             const uint4 d, uint32_t c, double not_cov, uint32_t &parity,
                                             uint32_t &malformed, const uint4 *next_desc, bool have_next) {
    const uint32_t lane = lane_id();
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t r0 = d.x, R = d.y & ~kSlowFlag;
    if (R == 0) {  // empty window (inside a big row)
        if (lane == 0) {
            w.tile_base[tile] = 0u;
            w.tile_total[tile] = 0u;
        }
        return;
    }
    if (PK && lane == 0) {
        // generic-proxy accesses of the previous tile (crossings written into the slab) before the async-proxy writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&ws.mbar, d.w * 8u);
        if (d.w) tma_load_1d(slab, a.iv + d.z, d.w * 8u, &ws.mbar);
    }
#pragma unroll 1
    for (uint32_t i = lane; i <= R; i += 32u) ws.row[i] = __ldg(a.rowptr + r0 + i);
#pragma unroll 1
    for (uint32_t i = lane; i < R; i += 32u) ws.len[i] = __ldg(a.len + r0 + i);
    __syncwarp();
    // ---- rows -> size classes (G = 1, 2, 4, 8, 16 lanes); trivial rows (k <= c: depth never exceeds c)
    //      are finished here; big rows (k > 256) were finished by big_kernel ----
    uint32_t cnt0 = 0, cnt1 = 0, cnt2 = 0, cnt3 = 0, cnt4 = 0;
    uint32_t my_cls[kMaxTileReads / 32], my_rank[kMaxTileReads / 32];
    bool any_big = false;
#pragma unroll
    for (uint32_t u = 0; u < kMaxTileReads / 32; ++u) {
        const uint32_t i = u * 32u + lane;
        uint32_t cl = 0xFFu;
        if (u * 32u < R) {
            if (i < R) {
                const uint32_t k = ws.row[i + 1] - ws.row[i], len = ws.len[i];
                if (!PK && k > kSmallMaxK) {
                    cl = 0xFEu;
                    ws.nup[i] = kRowDone;
                } else if (k <= c) {
                    const uint32_t ng = len != 0u;
                    ws.meta[i] = ng | (ng << 30) | (ng << 31);
                    ws.cls[i] = (uint8_t)classify(len, len, 0u, not_cov);
                    ws.nup[i] = kRowDone;
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
                    cl = k <= 16u ? 0u : (k <= 32u ? 1u : (k <= 64u ? 2u : (k <= 128u ? 3u : 4u)));
                }
            }
            any_big |= cl == 0xFEu;
            const uint32_t m0 = __ballot_sync(FULL, cl == 0u), m1 = __ballot_sync(FULL, cl == 1u);
            const uint32_t m2 = __ballot_sync(FULL, cl == 2u), m3 = __ballot_sync(FULL, cl == 3u);
            const uint32_t m4 = __ballot_sync(FULL, cl == 4u);
            my_rank[u] = cl == 0u ? cnt0 + __popc(m0 & lt)
                       : cl == 1u ? cnt1 + __popc(m1 & lt)
                       : cl == 2u ? cnt2 + __popc(m2 & lt)
                       : cl == 3u ? cnt3 + __popc(m3 & lt) : cnt4 + __popc(m4 & lt);
            cnt0 += __popc(m0);
            cnt1 += __popc(m1);
            cnt2 += __popc(m2);
            cnt3 += __popc(m3);
            cnt4 += __popc(m4);
        }
        my_cls[u] = cl;
    }
    const bool tile_has_big = !PK && __any_sync(FULL, any_big);
    // row order: class 4 (G = 16) first; lane position p of a row = lane base of its class + rank * G
    const uint32_t rb4 = 0, rb3 = cnt4, rb2 = rb3 + cnt3, rb1 = rb2 + cnt2, rb0 = rb1 + cnt1;
    const uint32_t lb3 = 16u * cnt4, lb2 = lb3 + 8u * cnt3, lb1 = lb2 + 4u * cnt2, lb0 = lb1 + 2u * cnt1;
    const uint32_t lanes_total = lb0 + cnt0;
#pragma unroll
    for (uint32_t u = 0; u < kMaxTileReads / 32; ++u) {
        const uint32_t i = u * 32u + lane, cl = my_cls[u];
        if (cl < 5u) ws.order[(cl == 0u ? rb0 : cl == 1u ? rb1 : cl == 2u ? rb2 : cl == 3u ? rb3 : rb4) + my_rank[u]] = (uint8_t)i;
        if (i < R && !tile_has_big) ws.soff[i] = (uint16_t)(ws.row[i] - (ws.row[0] & ~1u));
    }
    if (!PK) {
        // ---- generic pass: stage the slab now, one TMA bulk copy per run of non-big rows ----
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (!tile_has_big) {
                const uint32_t cs = ws.row[0] & ~1u, ce_ = (ws.row[R] + 1u) & ~1u, nb = (ce_ - cs) * 8u;
                mbar_expect_tx(&ws.mbar, nb);
                if (nb) tma_load_1d(slab, a.iv + cs, nb, &ws.mbar);
            } else {
                uint32_t bytes_total = 0;
                for (int pass = 0; pass < 2; ++pass) {  // expect_tx must precede the copies
                    uint32_t dd = 0, i = 0;
                    while (i < R) {
                        while (i < R && ws.row[i + 1] - ws.row[i] > kSmallMaxK) {
                            ws.soff[i] = 0;
                            ++i;
                        }
                        if (i >= R) break;
                        uint32_t j = i;
                        const uint32_t cs = ws.row[i] & ~1u;
                        while (j < R && ws.row[j + 1] - ws.row[j] <= kSmallMaxK) {
                            ws.soff[j] = (uint16_t)(dd + (ws.row[j] - cs));
                            ++j;
                        }
                        const uint32_t ce_ = (ws.row[j] + 1u) & ~1u, nb = (ce_ - cs) * 8u;
                        if (pass == 0) bytes_total += nb;
                        else if (nb) tma_load_1d(slab + dd, a.iv + cs, nb, &ws.mbar);
                        dd += ce_ - cs;
                        i = j;
                    }
                    if (pass == 0) mbar_expect_tx(&ws.mbar, bytes_total);
                }
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
        const uint32_t cl = p < lb3 ? 4u : (p < lb2 ? 3u : (p < lb1 ? 2u : (p < lb0 ? 1u : 0u)));
        const uint32_t G = 1u << cl;
        const uint32_t rel = p - (cl == 4u ? 0u : cl == 3u ? lb3 : cl == 2u ? lb2 : cl == 1u ? lb1 : lb0);
        const uint32_t rank = rel >> cl, g = rel & (G - 1u);
        const uint32_t i_row = valid ? ws.order[(cl == 0u ? rb0 : cl == 1u ? rb1 : cl == 2u ? rb2 : cl == 3u ? rb3 : rb4) + rank] : 0u;
        const uint32_t gmax = __shfl_sync(FULL, G, 0);  // classes are laid out largest first
        process_batch<PK>(ws, slab, valid, i_row, G, g, gmax, c, malformed);
    }
    // the next tile's slab: HBM -> L2 while this tile finishes
    if (PK && have_next && lane == 0) {
        const uint4 dn = *next_desc;
        if (!(dn.y & kSlowFlag) && dn.w) tma_prefetch_l2(a.iv + dn.z, dn.w * 8u);
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
    // ---- per row (one lane each): bad-region count, flags, class; exclusive scan inside the tile ----
    uint32_t carry = 0;
#pragma unroll 1
    for (uint32_t i0 = 0; i0 < R; i0 += 32u) {
        const uint32_t i = i0 + lane;
        uint32_t ng = 0;
        if (i < R) {
            const uint32_t n_up = ws.nup[i];
            if (n_up != kRowDone) {
                const uint32_t len = ws.len[i];
                const uint32_t *C = reinterpret_cast<const uint32_t *>(slab + ws.soff[i]);
                uint32_t h, tail, bad = len;
                if (n_up) {
                    h = C[0] != 0u;
                    tail = C[2u * n_up - 1u] != len;
                    ng = n_up - 1u + h + tail;
#pragma unroll 1
                    for (uint32_t j = 0; j < n_up; ++j) bad += C[2u * j] - C[2u * j + 1u];
                } else {
                    ng = h = tail = len != 0u;
                }
                ws.meta[i] = ng | (h << 30) | (tail << 31);
                ws.cls[i] = (uint8_t)classify(bad, len, n_up, not_cov);
            } else {
                ng = ws.meta[i] & 0x3FFFFFFFu;
            }
        }
        const uint32_t incl = warp_incl_scan(ng);
        if (i < R) ws.goff[i] = carry + incl - ng;
        carry += __shfl_sync(FULL, incl, 31);
    }
    const uint32_t tile_total = carry;
    uint32_t base = 0;
    if (lane == 0) {
        if (tile_total) base = atomicAdd(a.counters + kCntStage, tile_total);
        w.tile_base[tile] = base;
        w.tile_total[tile] = tile_total;
    }
    base = __shfl_sync(FULL, base, 0);
    __syncwarp();
    // ---- pass B: classes, in-tile offsets and the tile's bad regions (staging segment) ----
    for (uint32_t i = lane; i < R; i += 32u) {
        const uint32_t m = ws.meta[i], len = ws.len[i], k = ws.row[i + 1] - ws.row[i];
        const uint32_t at = base + ws.goff[i];
        a.gap_ptr[r0 + i] = ws.goff[i];
        a.cls[r0 + i] = ws.cls[i];
        if (!PK && k > kSmallMaxK) {
            const uint32_t j = w.big_slot[r0 + i];
            const uint2 *src = w.big_gaps + w.big_off[j];
            for (uint32_t gq = 0; gq < m; ++gq) w.stage[at + gq] = src[gq];
        } else {
            const uint32_t ng = m & 0x3FFFFFFFu, h = (m >> 30) & 1u, tail = m >> 31;
            const uint32_t *C = reinterpret_cast<const uint32_t *>(slab + ws.soff[i]);
            for (uint32_t gq = 0; gq < ng; ++gq) {
                const uint32_t f0 = 2u * gq, f1 = f0 + 1u;
                uint2 o;
                o.x = (f0 == 0u && h) ? 0u : C[f0 + 1u - 2u * h];
                o.y = (f1 == 2u * ng - 1u && tail) ? len : C[f1 + 1u - 2u * h];
                w.stage[at + gq] = o;
            }
        }
    }
    __syncwarp();
}

template <bool PK>
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
    uint32_t parity = 0, malformed = 0;
    if (PK) {
        // dynamic schedule; the next tile's index and descriptor are fetched while this tile is processed
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(a.counters + kCntTile, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        uint4 d = make_uint4(0, 0, 0, 0);
        if (tile < w.n_tiles) d = __ldg(w.tile_desc + tile);
        while (tile < w.n_tiles) {
            uint32_t nxt = 0;
            if (lane == 0) nxt = atomicAdd(a.counters + kCntTile, 1u);
            nxt = __shfl_sync(FULL, nxt, 0);
            const bool have_next = nxt < w.n_tiles;
            if (!(d.y & kSlowFlag))
                process_tile<true>(a, w, ws, slab, tile, d, c, not_cov, parity, malformed, w.tile_desc + nxt, have_next);
            tile = nxt;
            if (have_next) d = __ldg(w.tile_desc + tile);
        }
    } else {
        const uint32_t n_warps = gridDim.x * kFusedWarps;
        for (uint32_t tile = blockIdx.x * kFusedWarps + wid; tile < w.n_tiles; tile += n_warps) {
            const uint4 d = __ldg(w.tile_desc + tile);
            if (d.y & kSlowFlag) process_tile<false>(a, w, ws, slab, tile, d, c, not_cov, parity, malformed, nullptr, false);
        }
    }
    malformed = warp_sum(malformed);
    if (lane == 0 && malformed) atomicAdd(a.counters + kCntMalformed, malformed);
}

constexpr size_t kFusedSmemBytes = kWarpSmemBytes * kFusedWarps;

// ------------------------------------------------------------------------------------------------
// ordering pass: per-tile totals -> tile offsets -> ordered bad-region CSR, bitmap, histogram
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) scan_tiles_kernel(DetectArgs a, Work w) {
    __shared__ uint32_t sh[32];
    const uint32_t tid = threadIdx.x, n = w.n_tiles;
    const uint32_t per = (n + 1023u) / 1024u, beg = min(tid * per, n), end = min(beg + per, n);
    uint32_t s = 0;
    for (uint32_t i = beg; i < end; ++i) s += w.tile_total[i];
    const uint32_t incl = warp_incl_scan(s);
    if ((tid & 31u) == 31u) sh[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32u) {
        const uint32_t v = sh[tid], iv = warp_incl_scan(v);
        sh[tid] = iv - v;
        if (tid == 31u) {
            w.tile_off[n] = iv;
            a.gap_ptr[a.n_reads] = iv;
        }
    }
    __syncthreads();
    uint32_t run = sh[tid >> 5] + incl - s;
    for (uint32_t i = beg; i < end; ++i) {
        w.tile_off[i] = run;
        run += w.tile_total[i];
    }
}

// One warp per tile: gap_ptr += tile offset; staging segment -> final position.
__global__ void __launch_bounds__(256) finalize_kernel(DetectArgs a, Work w) {
    const uint32_t lane = lane_id();
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t tile = warp; tile < w.n_tiles; tile += n_warps) {
        const uint4 d = __ldg(w.tile_desc + tile);
        const uint32_t r0 = d.x, R = d.y & ~kSlowFlag;
        const uint32_t off = w.tile_off[tile], base = w.tile_base[tile], tot = w.tile_total[tile];
        for (uint32_t i = lane; i < R; i += 32u) a.gap_ptr[r0 + i] += off;
        for (uint32_t j = lane; j < tot; j += 32u) a.gaps[off + j] = w.stage[base + j];
    }
}

// One thread per 16 reads = one 32-bit word of the 2-bit bitmap; class histogram.
__global__ void __launch_bounds__(256) bitmap_kernel(const uint8_t *__restrict__ cls, uint32_t n, uint8_t *__restrict__ bitmap,
                                                      uint32_t *counters) {
    const uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) * 16u;
    uint32_t h0 = 0, h1 = 0, h2 = 0;
    if (base < n) {
        uint32_t wv[4];
        if (base + 16u <= n) {
            const uint4 v = *reinterpret_cast<const uint4 *>(cls + base);
            wv[0] = v.x; wv[1] = v.y; wv[2] = v.z; wv[3] = v.w;
        } else {
            wv[0] = wv[1] = wv[2] = wv[3] = 0u;
            for (uint32_t i = 0; base + i < n; ++i) wv[i >> 2] |= (uint32_t)cls[base + i] << (8u * (i & 3u));
            h0 -= 16u - (n - base);  // the zero padding is not NotBad
        }
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t cl = (wv[i >> 2] >> (8 * (i & 3))) & 3u;
            bits |= cl << (2 * i);
            h0 += cl == 0u;
            h1 += cl == 1u;
            h2 += cl == 2u;
        }
        reinterpret_cast<uint32_t *>(bitmap)[base >> 4] = bits;
    }
    h0 = warp_sum(h0);
    h1 = warp_sum(h1);
    h2 = warp_sum(h2);
    if (lane_id() == 0) {
        if (h0) atomicAdd(counters + kCntNotBad, h0);
        if (h1) atomicAdd(counters + kCntChimeric, h1);
        if (h2) atomicAdd(counters + kCntNotCovered, h2);
    }
}

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
    w.tile_desc = reinterpret_cast<uint4 *>(take(sizeof(uint4) * ((size_t)w.n_tiles + 1)));
    w.tile_base = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * ((size_t)w.n_tiles + 1)));
    w.tile_total = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * ((size_t)w.n_tiles + 1)));
    w.tile_off = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * ((size_t)w.n_tiles + 2)));
    w.stage = reinterpret_cast<uint2 *>(take(sizeof(uint2) * ((size_t)a.n_iv + a.n_reads + 1)));
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
    static int n_sm = 0, occ_fast = 0, occ_slow = 0;
    if (!n_sm) {
        int dev = 0, sm = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return -1;
        if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kBigSmemEvents * sizeof(uint32_t))) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemBytes) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemBytes) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_fast, fused_kernel<true>, kFusedThreads, kFusedSmemBytes) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_slow, fused_kernel<false>, kFusedThreads, kFusedSmemBytes) != cudaSuccess) return -1;
        if (occ_fast < 1 || occ_slow < 1) return -1;
        n_sm = sm;
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
    const bool any_slow = a.rows.n_big || a.rows.n_wide;
    const uint32_t plan_items = a.rows.n_big ? (a.n_reads > w.n_tiles ? a.n_reads : w.n_tiles) : w.n_tiles;
    uint32_t plan_blocks = (plan_items + 255u) / 256u;
    if (plan_blocks > (uint32_t)n_sm * 8u) plan_blocks = (uint32_t)n_sm * 8u;
    plan_kernel<<<plan_blocks, 256, 0, stream>>>(a, w, any_slow ? 1 : 0);
    ++launches;
    if (a.rows.n_big) {
        uint32_t grid = (uint32_t)n_sm * 2u;
        if (grid > a.rows.n_big) grid = (uint32_t)a.rows.n_big;
        big_kernel<<<grid, kBigThreads, kBigSmemEvents * sizeof(uint32_t), stream>>>(a, w, c, not_coverage);
        ++launches;
    }
    {
        uint32_t grid = (uint32_t)(n_sm * occ_fast);
        const uint32_t want = (w.n_tiles + kFusedWarps - 1u) / kFusedWarps;
        if (grid > want) grid = want;
        fused_kernel<true><<<grid, kFusedThreads, kFusedSmemBytes, stream>>>(a, w, c, not_coverage);
        ++launches;
    }
    if (any_slow) {
        uint32_t grid = (uint32_t)(n_sm * occ_slow);
        const uint32_t want = (w.n_tiles + kFusedWarps - 1u) / kFusedWarps;
        if (grid > want) grid = want;
        fused_kernel<false><<<grid, kFusedThreads, kFusedSmemBytes, stream>>>(a, w, c, not_coverage);
        ++launches;
    }
    scan_tiles_kernel<<<1, 1024, 0, stream>>>(a, w);
    ++launches;
    {
        uint32_t blocks = (w.n_tiles + 7u) / 8u;
        if (blocks > (uint32_t)n_sm * 16u) blocks = (uint32_t)n_sm * 16u;
        finalize_kernel<<<blocks, 256, 0, stream>>>(a, w);
        ++launches;
    }
    {
        const uint32_t threads = (a.n_reads + 15u) / 16u;
        bitmap_kernel<<<(threads + 255u) / 256u, 256, 0, stream>>>(a.cls, a.n_reads, a.bitmap, a.counters);
        ++launches;
    }
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
