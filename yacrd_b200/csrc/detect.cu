// detect.cu — sm_100a kernels of the detect hot path.
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
//   Classification (editor/mod.rs:85-100) runs on the final bad-region list exactly as the reference does:
//   bad_len = sum(end - begin) in wrapping u32; NotCovered iff (double)bad_len / (double)len > n (same IEEE
//   divide, tested first); else Chimeric iff some region has begin != 0 && end != len; else NotBad.
//
// Kernels (all integer work; no tensor cores: there is no contraction on this path). DESIGN.md section 3 has the details.
//   per upload
//     row_stats_kernel  size-class histogram, big-row scratch needs, input sanity (from rowptr / len only)
//     scatter_kernel    every row -> its size class (G = 1,2,3,4,5,8,16 lanes x 32 keys; packed or wide) and a 16-byte
//                       worklist record {row, first interval, k | class, len}; rows with k > 512 -> scan / big lists
//   per detect step (no memset: the counters alternate between two sets, see pileup.cuh)
//     sort_kernel<VAL>  one persistent CTA per SM; warps take batches of floor(32 / G) rows of ONE class from a global
//                       cursor (regtier.cuh): record prefetch by cp.async, one TMA bulk copy per row (UBLKCP) onto the
//                       warp's mbarrier, keys sorted in registers as PACKED u16x2 (begin | end << 16: one VIMNMX.U16x2
//                       moves a begin and an end through the same network) or as two u32 arrays when the read is longer
//                       than 65534 bases, crossings by carry-chain compares against a transposed shared copy, the row's
//                       position list written to a bump-allocated staging buffer, meta[row] = {first pair, count}.
//                       VAL (first step after an upload) also tests 0 <= begin < end <= len; rows that fail go to
//     literal_kernel    the reference's own algorithm (sort + min-heap sweep, stack.rs:61-139), one thread per row
//     bigscan_kernel    (side stream) rows with 512 < k <= 65534 on reads shorter than 393216: begins / ends counted per
//                       position in shared memory, occupied positions compacted and scanned; no sort
//     big_kernel        (side stream) the remaining big rows: bitonic sort in shared memory / global scratch
//     order_kernel      one pass: parts of 1024 rows in ticket order publish their totals and sum the earlier ones,
//                       staged regions -> ordered bad-region CSR, classification, 2-bit bitmap (to every peer's gather
//                       buffer when peers are bound), class histogram; the last part closes the step
//     peer_wait_kernel  consumer side of the fused all-gather
//   classify_kernel     FromReport path: type_of_read over given bad regions
#include "pileup.cuh"
#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace yb {
namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr uint32_t INF = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// The same decision, with the double-precision division (a long instruction sequence) only where single precision
// cannot tell: f = bad / len in float is within 2^-20 relative of the exact quotient, so outside that band around
// not_cov the comparison is already decided. NaN / infinity on either side compare false twice and take the exact path.
struct NotCovBand {
    float lo, hi;
};
__device__ __forceinline__ NotCovBand not_cov_band(double not_cov) {
    const float f = (float)not_cov, w = fabsf(f) * 4e-6f + 1e-30f;
    return {f - w, f + w};
}
__device__ __forceinline__ uint32_t classify(uint32_t bad_len, uint32_t len, uint32_t n_up, double not_cov);
__device__ __forceinline__ uint32_t classify_fast(uint32_t bad_len, uint32_t len, uint32_t n_up, double not_cov, NotCovBand band) {
    const float f = __fdividef((float)bad_len, (float)len);
    if (f > band.hi) return 2u;
    if (f < band.lo) return n_up >= 2u ? 1u : 0u;
    return classify(bad_len, len, n_up, not_cov);
}

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

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
constexpr int E = 16;                          // keys per lane per array in the register tier
constexpr uint32_t kSmallMaxK = kRegisterTierMaxK;
#ifndef YB_SORT_WARPS
#define YB_SORT_WARPS (YB_KE == 32 ? 16 : 24)
#endif
constexpr uint32_t kSortWarps = YB_SORT_WARPS;  // warps of sort_kernel's one CTA per SM (they only share the batch counter)
constexpr uint32_t kSortThreads = kSortWarps * 32;
#ifndef YB_BUF_INTERVALS
#define YB_BUF_INTERVALS (32 * (YB_KE + 2))
#endif
constexpr uint32_t kBufIntervals = YB_BUF_INTERVALS;  // row slots of a warp's slab buffer: a batch of class G holds min(floor(32 / G),
                                                      // floor(kBufIntervals / (kE G + 2))) rows; 32 (kE + 2) = 32 rows of the G = 1 class
constexpr uint32_t kScatterRows = 1024;        // rows per CTA of scatter_kernel
#ifndef YB_ORDER_THREADS
#define YB_ORDER_THREADS 256
#endif
#ifndef YB_ORDER_ROWS
#define YB_ORDER_ROWS 4
#endif
constexpr uint32_t kOrderThreads = YB_ORDER_THREADS, kOrderRows = YB_ORDER_ROWS, kPartRows = kOrderThreads * kOrderRows;  // rows per CTA of order_kernel
static_assert(kOrderRows == 4, "order_kernel: a thread's 4 rows travel as one 16-byte word");
#ifndef YB_ORDER_UNROLL
#define YB_ORDER_UNROLL 4
#endif
constexpr uint32_t kOrderUnroll = YB_ORDER_UNROLL;
constexpr uint32_t kOrderMap = 1024;  // regions of a warp's 128 rows whose row is looked up in the warp's byte map  // regions a thread moves per turn (loads in flight)
#ifndef YB_FIXED_ROUNDS
#define YB_FIXED_ROUNDS 3
#endif
constexpr uint32_t kFixedRounds = YB_FIXED_ROUNDS;  // sort_kernel: batches per warp dealt by position (1 or 3), the rest by the cursor
static_assert(kFixedRounds == 1 || kFixedRounds == 3, "one or three fixed rounds");
constexpr uint32_t kStageChunk = 2048;       // pairs a warp reserves in the staging buffer per atomic (an L2 round trip the warp waits for)
constexpr uint32_t kRecValid = 0x80000000u;    // worklist record .z = k | class << 16 | kRecValid

// Host-built table of the size classes (sizes are known from the row pointers at upload time).
struct ClassTab {
    uint32_t entry_base[kNumClasses];     // where the class's records start in the worklist
    uint32_t count[kNumClasses];          // rows in the class
    uint32_t order[kNumClasses];          // classes in processing order (largest groups first, long reads before packed rows)
    uint32_t item_base[kNumClasses + 1];  // batches before the q-th class in processing order
    uint32_t lanes[kNumClasses];          // G: lanes per row
    uint32_t rpb[kNumClasses];            // rows per batch
    uint32_t inv[kNumClasses];            // ceil(65536 / G): x / G == (x * inv) >> 16 for x < 2048
};

// scratch carve-up
struct Work {
    uint4 *recs;                     // n_reads worklist records {row, first interval, k | class << 16 | valid, len}
    uint2 *meta;                     // n_reads: {where the row's bad regions sit in `stage` (pairs), how many}
    uint2 *stage;                    // bad regions in batch-completion order (warps reserve chunks with one atomic)
    uint32_t stage_cap;              // pairs
    unsigned long long *part_desc;   // n_parts: step tag << 32 | bad regions of the part (order_kernel publishes, later parts sum)
    uint32_t n_parts;
    uint32_t *lit_list;              // rows holding a malformed interval, found by a validating step (they take the literal heap sweep)
    uint32_t *big_list;              // rows with k > kSmallMaxK that are sorted by a CTA
    uint32_t *scan_list;             // rows with k > kSmallMaxK that are scanned by position (big_row_scans)
    uint32_t *huge_keys;             // event keys of rows beyond the shared-memory tier
};

// ------------------------------------------------------------------------------------------------
// scatter_kernel: rows -> worklist records grouped by size class (CTA-aggregated cursors)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScatterRows) scatter_kernel(DetectArgs a, Work w, ClassTab tab) {
    __shared__ uint32_t s_cnt[kNumClasses], s_base[kNumClasses];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, r = blockIdx.x * kScatterRows + tid;
    if (tid < (uint32_t)kNumClasses) s_cnt[tid] = 0u;
    __syncthreads();
    int cls = -2;  // 0 .. kNumClasses-1: lane-group classes; -1: big row
    uint32_t p0 = 0, k = 0, len = 0;
    if (r < a.n_reads) {
        p0 = __ldg(a.rowptr + r);
        k = __ldg(a.rowptr + r + 1) - p0;
        len = __ldg(a.len + r);
        cls = class_of_row(k, len);
        if (cls < 0) {
            if (big_row_scans(k, len)) {  // heavy rows (several windows, or many intervals) to the front, the others to the back
                if (len >= kScanWindow || k >= 2048u) w.scan_list[atomicAdd(a.counters + kCntScanList, 1u)] = r;
                else w.scan_list[(uint32_t)a.rows.n_scan - 1u - atomicAdd(a.counters + kCntScanListBack, 1u)] = r;
            } else {
                w.big_list[atomicAdd(a.counters + kCntBigList, 1u)] = r;
            }
        }
    }
    const uint32_t peers = __match_any_sync(FULL, cls);
    const uint32_t leader = __ffs(peers) - 1u, rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t wbase = 0;
    if (cls >= 0 && lane == leader) wbase = atomicAdd(&s_cnt[cls], (uint32_t)__popc(peers));
    wbase = __shfl_sync(FULL, wbase, leader);
    __syncthreads();
    if (tid < (uint32_t)kNumClasses && s_cnt[tid]) s_base[tid] = atomicAdd(a.counters + kCntClassCursor + tid, s_cnt[tid]);
    __syncthreads();
    if (cls >= 0) w.recs[tab.entry_base[cls] + s_base[cls] + wbase + rank] = make_uint4(r, p0, k | ((uint32_t)cls << 16) | kRecValid, len);
}

// ------------------------------------------------------------------------------------------------
// register tier: G lanes per row (one G per batch), E = 16 keys per lane, blocked layout (element =
// g*16 + t): shuffles only on the ceil(log2 G) outermost merge levels. G need not be a power of two: the
// network is the one for next_pow2(G) lanes whose missing top lanes hold +inf, and every exchange with a
// missing lane is a no-op, so it is simply predicated off.
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

// Sorts, for every group of G consecutive lanes, its 16*G keys (ascending in element order g*16 + t).
// `g` is the lane's index inside its group; lanes outside any group pass g = 0, in_group = false.
template <bool PK> __device__ __forceinline__ void sort_group(uint32_t (&key)[E], uint32_t G, uint32_t g, bool in_group) {
    sort16<PK>(key);
    const uint32_t lane = lane_id();
#pragma unroll 1
    for (uint32_t ls = 2; ls < 2u * G; ls <<= 1) {
        {   // flip: element e pairs with e ^ (16*ls - 1): partner lane g ^ (ls-1), slot 15 - t
            const uint32_t partner = g ^ (ls - 1u);
            const bool ex = in_group && partner < G, lo_half = (g & (ls >> 1)) == 0;
            const bool pmin = ex && lo_half, pmax = ex && !lo_half;
            const uint32_t src = ex ? lane + partner - g : lane;
            uint32_t other[E];
#pragma unroll
            for (int t = 0; t < E; ++t) other[t] = __shfl_sync(FULL, key[E - 1 - t], src);
#pragma unroll
            for (int t = 0; t < E; ++t) {
                if (pmin) key[t] = kmin<PK>(key[t], other[t]);
                if (pmax) key[t] = kmax<PK>(key[t], other[t]);
            }
        }
#pragma unroll 1
        for (uint32_t j = ls >> 2; j > 0; j >>= 1) {  // half-cleaners on the lane bits
            const uint32_t partner = g ^ j;
            const bool ex = in_group && partner < G, lo_half = (g & j) == 0;
            const bool pmin = ex && lo_half, pmax = ex && !lo_half;
            const uint32_t src = ex ? lane + partner - g : lane;
#pragma unroll
            for (int t = 0; t < E; ++t) {
                const uint32_t o = __shfl_sync(FULL, key[t], src);
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

// ------------------------------------------------------------------------------------------------
// CTA tier (big_kernel): rows with k > 512, one CTA per row. The same closed form as the register tier on arrays
// that live in shared memory (or, beyond kCtaMaxSmemWords, in a global slab): 512-key chunks are sorted by a warp in
// registers (sort_group, G = 32), larger strides are compare-exchanged in place, and every level is finished by
// warps merging 512-key blocks in registers again, so a level costs log2(size / 512) passes over the array
// instead of log2(size). PK: one array of begin | end << 16; else two u32 arrays (begins, ends).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kCtaThreads = 256;
constexpr uint32_t kCtaMaxSmemWords = 32768u + 2048u;  // 136 KB: k <= 32768 packed, k <= 16384 wide

__host__ __device__ inline uint32_t cta_idx(uint32_t e) { return e + (e >> 4); }  // 17-word pitch per 16 keys
__host__ __device__ inline uint64_t cta_words(uint64_t k, bool wide) {
    uint64_t K = 512;
    while (K < k) K <<= 1;
    return (wide ? 2ull : 1ull) * (K + (K >> 4));
}

// bitonic merge of the 512 keys a warp holds (blocked, 16 per lane) once they form a bitonic sequence
template <bool PK> __device__ __forceinline__ void merge512(uint32_t (&key)[E]) {
    const uint32_t lane = lane_id();
#pragma unroll 1
    for (uint32_t j = 16; j > 0; j >>= 1) {
        const bool lo_half = (lane & j) == 0;
#pragma unroll
        for (int t = 0; t < E; ++t) {
            const uint32_t o = __shfl_xor_sync(FULL, key[t], j);
            key[t] = lo_half ? kmin<PK>(key[t], o) : kmax<PK>(key[t], o);
        }
    }
#pragma unroll
    for (int s = E >> 1; s > 0; s >>= 1) {
#pragma unroll
        for (int t = 0; t < E; ++t)
            if ((t & s) == 0) ce<PK>(key[t], key[t | s]);
    }
}

template <bool PK> __device__ __forceinline__ void cta_ce(uint32_t *keys, uint32_t lo, uint32_t hi) {
    const uint32_t x = keys[cta_idx(lo)], y = keys[cta_idx(hi)];
    keys[cta_idx(lo)] = kmin<PK>(x, y);
    keys[cta_idx(hi)] = kmax<PK>(x, y);
}

// Finishes the sort of `arrays` arrays of K keys whose 512-key chunks are already sorted (K a power of two >= 512);
// array q sits at keys + q * (K + K / 16).
template <bool PK> __device__ void cta_merge_levels(uint32_t *keys, uint32_t K, uint32_t arrays) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5, nwarps = kCtaThreads / 32;
    const uint32_t pitch = K + (K >> 4);
    for (uint32_t size = 1024, lg = 10; size <= K; size <<= 1, ++lg) {
        for (uint32_t q = 0; q < arrays; ++q) {
            uint32_t *a = keys + q * pitch;
            for (uint32_t p = tid; p < (K >> 1); p += kCtaThreads) {  // flip: e <-> e ^ (size - 1)
                const uint32_t blk = p >> (lg - 1), o = p & ((size >> 1) - 1u);
                cta_ce<PK>(a, (blk << lg) + o, (blk << lg) + size - 1u - o);
            }
        }
        __syncthreads();
        for (uint32_t stride = size >> 2, ls = lg - 2; stride >= 512u; stride >>= 1, --ls) {
            for (uint32_t q = 0; q < arrays; ++q) {
                uint32_t *a = keys + q * pitch;
                for (uint32_t p = tid; p < (K >> 1); p += kCtaThreads) {
                    const uint32_t lo = ((p >> ls) << (ls + 1)) | (p & (stride - 1u));
                    cta_ce<PK>(a, lo, lo + stride);
                }
            }
            __syncthreads();
        }
        for (uint32_t q = 0; q < arrays; ++q) {  // strides 256 .. 1: 512-key blocks, in registers
            uint32_t *a = keys + q * pitch;
            for (uint32_t blk = wid; blk < (K >> 9); blk += nwarps) {
                uint32_t *b = a + cta_idx(blk * 512u) + 17u * lane;
                uint32_t key[E];
#pragma unroll
                for (int t = 0; t < E; ++t) key[t] = b[t];
                merge512<PK>(key);
#pragma unroll
                for (int t = 0; t < E; ++t) b[t] = key[t];
            }
        }
        __syncthreads();
    }
}

// Block-wide exclusive scan of one value per thread (kCtaThreads threads); *total gets the sum. sh: 8 words.
__device__ __forceinline__ uint32_t cta_excl_scan(uint32_t v, uint32_t *sh, uint32_t *total) {
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    const uint32_t incl = warp_incl_scan(v);
    __syncthreads();
    if (lane == 31u) sh[wid] = incl;
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (uint32_t q = 0; q < kCtaThreads / 32; ++q) {
        const uint32_t x = sh[q];
        before += q < wid ? x : 0u;
        tot += x;
    }
    *total = tot;
    return before + incl - v;
}

template <bool PK>
__device__ void cta_row(const DetectArgs &a, const Work &w, uint32_t *cnt, uint32_t *keys, uint32_t r, uint32_t c, uint32_t *sh /* 16 u32 */) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5, nwarps = kCtaThreads / 32;
    const uint32_t s = a.rowptr[r], k = a.rowptr[r + 1] - s, len = a.len[r];
    const uint2 *row = a.iv + s;
    uint32_t K = 512;
    while (K < k) K <<= 1;
    const uint32_t pitch = K + (K >> 4);
    uint32_t *kB = keys, *kE = keys + (PK ? 0u : pitch);
    // ---- 512-key chunks: load, validate, sort in registers, store blocked ----
    bool bad_iv = false;
    for (uint32_t ch = wid; ch < (K >> 9); ch += nwarps) {
        uint32_t K0[E], K1[PK ? 1 : E];
#pragma unroll
        for (int t = 0; t < E; ++t) {
            const uint32_t e = ch * 512u + (uint32_t)t * 32u + lane;
            uint2 v = make_uint2(INF, INF);
            if (e < k) {
                v = __ldg(row + e);
                bad_iv |= !(v.x < v.y && v.y <= len);
            }
            if (PK) {
                K0[t] = __byte_perm(v.x, v.y, 0x5410);
            } else {
                K0[t] = v.x;
                K1[PK ? 0 : t] = v.y;
            }
        }
        sort_group<PK>(K0, 32u, lane, true);
        uint32_t *b = kB + cta_idx(ch * 512u) + 17u * lane;
#pragma unroll
        for (int t = 0; t < E; ++t) b[t] = K0[t];
        if (!PK) {
            uint32_t Kx[E];
#pragma unroll
            for (int t = 0; t < E; ++t) Kx[t] = K1[PK ? 0 : t];
            sort_group<PK>(Kx, 32u, lane, true);
            uint32_t *eb = kE + cta_idx(ch * 512u) + 17u * lane;
#pragma unroll
            for (int t = 0; t < E; ++t) eb[t] = Kx[t];
        }
    }
    if (a.validate && bad_iv) sh[12] = 1u;  // (cleared by thread 0 at the end of the row)
    __syncthreads();
    cta_merge_levels<PK>(keys, K, PK ? 1u : 2u);
    // ---- crossings: thread t owns the contiguous slots [i0, i1) ----
    const uint32_t BINF = PK ? 0xFFFFu : INF;
    const uint32_t cc = min(c, K);  // beyond k every threshold behaves the same
    auto Bv = [&](uint32_t i) -> uint32_t { return i >= K ? BINF : (PK ? (kB[cta_idx(i)] & 0xFFFFu) : kB[cta_idx(i)]); };
    auto Ev = [&](uint32_t i, uint32_t back) -> uint32_t {  // E[i - back], 0 below the first end
        if (i < back) return 0u;
        const uint32_t j = i - back;
        return PK ? (kE[cta_idx(j)] >> 16) : kE[cta_idx(j)];
    };
    const uint32_t per = K / kCtaThreads, i0 = tid * per, i1 = i0 + per;
    uint32_t nu = 0, nd = 0, firstU = 0, lastD = 0;
    for (uint32_t i = i0; i < i1; ++i) {
        const uint32_t bi = Bv(i), bn = Bv(i + 1u), e1 = Ev(i, cc + 1u), e0 = Ev(i, cc);
        const bool v1 = e1 <= bi, v0 = e0 <= bi, v1n = e0 <= bn;
        if (v1 && !v0) {
            if (nu == 0) firstU = bi;
            ++nu;
        }
        if (!v0 && v1n) {
            lastD = e0;
            ++nd;
        }
    }
    uint32_t n_up, n_down;
    const uint32_t ru0 = cta_excl_scan(nu, sh, &n_up);
    const uint32_t rd0 = cta_excl_scan(nd, sh, &n_down);
    __syncthreads();
    if (nu && ru0 == 0u) sh[8] = firstU;
    if (nd && rd0 + nd == n_down) sh[9] = lastD;
    if (tid == 0) sh[10] = atomicAdd(cnt + kCntStage, k + 1u);  // room for the row's k + 1 possible bad regions
    __syncthreads();
    const uint32_t U0 = sh[8], Dl = sh[9], at = sh[10];
    uint32_t ng, h = 0, tail = 0;
    if (n_up) {
        h = U0 != 0u;
        tail = Dl != len;
        ng = n_up - 1u + h + tail;
    } else {
        ng = h = tail = len != 0u;
    }
    const bool fits = (uint64_t)at + k + 1u <= w.stage_cap;
    if (!fits) ng = 0;
    uint32_t *F = reinterpret_cast<uint32_t *>(w.stage + at);
    if (fits && n_up) {
        // flat layout [0 if h] U0 D0 U1 D1 ... [len if tail]: U_j sits at 2j - 1 + 2h, D_j at 2j + 2h
        uint32_t ru = ru0, rd = rd0;
        for (uint32_t i = i0; i < i1; ++i) {
            const uint32_t bi = Bv(i), bn = Bv(i + 1u), e1 = Ev(i, cc + 1u), e0 = Ev(i, cc);
            const bool v1 = e1 <= bi, v0 = e0 <= bi, v1n = e0 <= bn;
            if (v1 && !v0) {
                const int f = 2 * (int)ru - 1 + 2 * (int)h;
                if (f >= 0) F[f] = bi;
                ++ru;
            }
            if (!v0 && v1n) {
                const uint32_t f = 2u * rd + 2u * h;
                if (f < 2u * ng) F[f] = e0;
                ++rd;
            }
        }
    }
    if (tid == 0) {
        if (fits) {
            if (h) F[0] = 0u;
            if (tail) F[2u * ng - 1u] = len;
        } else {
            atomicAdd(cnt + kCntStageOverflow, 1u);
        }
        if (sh[12]) {  // validating step, malformed interval: literal_kernel computes the row (see process_batch_t)
            w.lit_list[atomicAdd(a.counters + kCntLiteralList, 1u)] = r;
            atomicAdd(a.counters + kCntMalformedIv, 1u);  // (at least one; the exact count is not kept on this path)
            sh[12] = 0u;
        } else {
            w.meta[r] = make_uint2(at, ng);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kCtaThreads) big_kernel(DetectArgs a, Work w, uint32_t c, uint32_t smem_words) {
    extern __shared__ __align__(16) uint32_t cta_smem[];
    __shared__ uint32_t sh[16];
    if (threadIdx.x == 0) sh[12] = 0u;
    __syncthreads();
    const uint32_t n_big = min((uint32_t)(a.rows.n_big - a.rows.n_scan), __ldcg(a.counters + kCntBigList));  // (rows redone by literal_kernel are not listed)
    const uint32_t ep = __ldcg(a.counters + kCntEpoch);
    uint32_t *cnt = a.counters + (ep & 1u) * kNumCounters;
    for (uint32_t j = blockIdx.x; j < n_big; j += gridDim.x) {
        const uint32_t r = w.big_list[j];
        const uint32_t k = a.rowptr[r + 1] - a.rowptr[r];
        const bool wide = a.len[r] > kPackedMaxLen;
        const uint64_t words = cta_words(k, wide);
        uint32_t *keys = cta_smem;
        if (words > smem_words) {  // keys live in a bump-allocated global slab
            if (threadIdx.x == 0) sh[11] = atomicAdd(cnt + kCntHugeBump, (uint32_t)words);
            __syncthreads();
            keys = w.huge_keys + sh[11];
        }
        if (wide) cta_row<false>(a, w, cnt, keys, r, c, sh);
        else cta_row<true>(a, w, cnt, keys, r, c, sh);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA tier, scan path (bigscan_kernel): a row with k > 512 whose read is short enough is not sorted. This is the heap
// sweep of stack.rs:71-105 seen from the positions: cnt[x] counts the begins (low half) and the ends (high half) at
// every position x of a window of the read (shared-memory atomics), a block scan gives the depth in front of every
// thread's stretch of positions, and a walk over the stretch finds the crossings: at x the ends pop first
// (stack.rs:72-81: a down-crossing when the depth falls from above c to c or less), then the begins push
// (stack.rs:83-90: an up-crossing when it rises from c or less to above c) - a zero-length region (x, x) when both
// happen. Work is O(len + k) instead of O(k log^2 k): a row of 5000 intervals takes a few microseconds of one CTA, two
// CTAs per SM hide each other's load latencies. The crossings go to the staging buffer as the same pair list the
// register tier writes (k + 1 pairs reserved up front: the ranks are only known window by window).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kScanThreads = kScanWindow / 32;  // one thread per 32 positions of a window
static_assert(kScanThreads <= 1024 && kScanThreads % 32 == 0, "bigscan_kernel's block scan handles up to 32 warps");
constexpr size_t kScanSmemBytes = (sizeof(uint32_t) + sizeof(uint16_t)) * kScanWindow;  // counters + the list of occupied positions

__global__ void __launch_bounds__(kScanThreads, 2) bigscan_kernel(DetectArgs a, Work w, uint32_t c) {
    // cnt_x[x]: begins | ends << 16 at position x of the window; occ[t]: which of positions 32 t .. 32 t + 31 hold anything;
    // lst[]: the occupied positions, compacted, so that every thread walks an equal share of them (the events of a read
    // crowd at its two ends: a thread per stretch of positions would leave one thread with all the work). Only occupied
    // positions are ever looked at again (a read of 14 K bases holds a few thousand), and they are cleared on the way out,
    // so nothing is zeroed per row.
    extern __shared__ __align__(16) uint32_t cnt_x[];
    uint16_t *lst = reinterpret_cast<uint16_t *>(cnt_x + kScanWindow);
    __shared__ uint32_t occ[kScanThreads], s_w[32], s_v[4];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t ep = __ldcg(a.counters + kCntEpoch);
    uint32_t *cnt = a.counters + (ep & 1u) * kNumCounters;
    const uint32_t n_scan = (uint32_t)a.rows.n_scan;  // (scatter_kernel filled the list from both ends: heavy rows first)
    __shared__ uint32_t s_nextj;
    const uint32_t cc = min(c, 0x7FFFFFF0u);
    for (uint32_t x = tid; x < kScanWindow; x += kScanThreads) cnt_x[x] = 0u;
    occ[tid] = 0u;
    if (tid < 32u) s_w[tid] = 0u;
    auto block_excl = [&](uint32_t v, uint32_t *total) {  // exclusive scan of one value per thread
        const uint32_t incl = warp_incl_scan(v);
        __syncthreads();
        if (lane == 31u) s_w[wid] = incl;
        __syncthreads();
        const uint32_t x = s_w[lane];  // (warps beyond the CTA's hold 0)
        const uint32_t wi = warp_incl_scan(x);
        *total = __shfl_sync(FULL, wi, 31);
        return __shfl_sync(FULL, wi - x, wid) + incl - v;
    };
    // The next row's size, place and first intervals travel while this row is scanned (registers): a row is a chain of
    // dependent loads otherwise (list entry -> row pointers -> intervals), paid by the whole CTA.
    constexpr uint32_t NPF = 6;  // intervals per thread held ahead (rows of up to 3072 intervals entirely)
    uint32_t r_n = 0, s_n = 0, k_n = 0, len_n = 0;
    uint2 pf[NPF];
    auto fetch_meta = [&](uint32_t jn) {
        if (jn < n_scan) {
            r_n = w.scan_list[jn];
            s_n = a.rowptr[r_n];
            k_n = a.rowptr[r_n + 1] - s_n;
            len_n = a.len[r_n];
        }
    };
    auto fetch_ivs = [&](uint32_t jn) {
        if (jn < n_scan) {
#pragma unroll
            for (uint32_t u = 0; u < NPF; ++u)
                if (tid + u * kScanThreads < k_n) pf[u] = __ldg(a.iv + s_n + tid + u * kScanThreads);
        }
    };
    fetch_meta(blockIdx.x);
    fetch_ivs(blockIdx.x);
    // rows are dealt in list order from a cursor (the first one per CTA by its index): rows cost anything between one
    // window of 600 intervals and several windows of 5000, and a fixed deal left SMs idle for a fifth of the kernel
    for (uint32_t j = blockIdx.x, jn = 0; j < n_scan; j = jn) {
        const uint32_t r = r_n, s = s_n, k = k_n, len = len_n;
        const uint32_t n_pos = len + 1u;  // positions 0 .. len
        if (tid == 0) {
            s_nextj = gridDim.x + atomicAdd(cnt + kCntScanTicket, 1u);
            s_v[2] = atomicAdd(cnt + kCntStage, k + 1u);  // the pair list P[q] = (D_{q-1}, U_q), q = 0 .. n_up <= k
            s_v[3] = 0u;                                  // validating step: the row holds a malformed interval
        }
        uint32_t nbad = 0;
        uint32_t depth = 0, ups = 0, downs = 0;  // carried from window to window (the same in every thread)
        for (uint32_t lo = 0; lo < n_pos; lo += kScanWindow) {
            const uint32_t m = min(kScanWindow, n_pos - lo);
            __syncthreads();  // the previous window is cleared
            jn = s_nextj;
            auto count = [&](const uint2 v) {
                if (a.validate && lo == 0u) nbad += !(v.x < v.y && v.y <= len);
                const uint32_t xb = v.x - lo, xe = v.y - lo;
                if (xb < m) {
                    atomicAdd(&cnt_x[xb], 1u);
                    if (!(occ[xb >> 5] & (1u << (xb & 31u)))) atomicOr(&occ[xb >> 5], 1u << (xb & 31u));  // (a stale read costs an OR)
                }
                if (xe < m) {
                    atomicAdd(&cnt_x[xe], 0x10000u);
                    if (!(occ[xe >> 5] & (1u << (xe & 31u)))) atomicOr(&occ[xe >> 5], 1u << (xe & 31u));
                }
            };
#pragma unroll
            for (uint32_t u = 0; u < NPF; ++u)
                if (tid + u * kScanThreads < k) count(pf[u]);
            for (uint32_t i = tid + NPF * kScanThreads; i < k; i += kScanThreads) count(__ldg(a.iv + s + i));
            if (nbad) s_v[3] = 1u;
            if (lo + kScanWindow >= n_pos) {  // last window: the registers are free for the next row
                fetch_meta(jn);
            }
            __syncthreads();
            // compact the occupied positions (ascending), then thread t takes entries [e0, e1)
            uint32_t n_occ;
            {
                const uint32_t bits = occ[tid];
                uint32_t at = block_excl(__popc(bits), &n_occ);
                for (uint32_t bm = bits; bm; bm &= bm - 1u) lst[at++] = (uint16_t)(32u * tid + __ffs(bm) - 1);
                occ[tid] = 0u;
            }
            __syncthreads();
            const uint32_t per = (n_occ + kScanThreads - 1u) / kScanThreads, e0 = min(tid * per, n_occ), e1 = min(e0 + per, n_occ);
            uint32_t net = 0;
            for (uint32_t e = e0; e < e1; ++e) {
                const uint32_t q = cnt_x[lst[e]];
                net += (q & 0xFFFFu) - (q >> 16);  // (wrapping u32: the sums are exact)
            }
            uint32_t net_all;
            const uint32_t d0 = depth + block_excl(net, &net_all);  // depth in front of entry e0
            auto walk = [&](auto on_up, auto on_down) {
                uint32_t d = d0;
                for (uint32_t e = e0; e < e1; ++e) {
                    const uint32_t x = lst[e], q = cnt_x[x];
                    const uint32_t ne = q >> 16, nb = q & 0xFFFFu;
                    if (d > cc && d - ne <= cc) on_down(lo + x);  // (no end here: ne = 0, never true)
                    d -= ne;
                    if (d <= cc && d + nb > cc) on_up(lo + x);    // (no begin here: nb = 0, never true)
                    d += nb;
                }
            };
            uint32_t nu = 0, nd = 0, first_u = 0, last_d = 0;
            walk([&](uint32_t x) { if (!nu) first_u = x; ++nu; }, [&](uint32_t x) { last_d = x; ++nd; });
            uint32_t tot_ud;
            const uint32_t r0 = block_excl(nu | (nd << 16), &tot_ud);
            const uint32_t ru0 = ups + (r0 & 0xFFFFu), rd0 = downs + (r0 >> 16);
            if (nu && ru0 == 0u) s_v[0] = first_u;                          // first up-crossing of the row
            if (nd && (r0 >> 16) + nd == (tot_ud >> 16)) s_v[1] = last_d;   // last down-crossing so far
            const uint32_t base = s_v[2];  // (written before the first barrier of the row)
            if ((nu | nd) && (uint64_t)base + k + 1u <= w.stage_cap) {
                uint32_t *P = reinterpret_cast<uint32_t *>(w.stage + base);
                uint32_t ru = ru0, rd = rd0;
                walk([&](uint32_t x) { P[2u * ru++ + 1u] = x; }, [&](uint32_t x) { P[2u * rd++ + 2u] = x; });
            }
            for (uint32_t e = e0; e < e1; ++e) cnt_x[lst[e]] = 0u;  // leave the window clean
            if (lo + kScanWindow >= n_pos) fetch_ivs(jn);
            depth += net_all;
            ups += tot_ud & 0xFFFFu;
            downs += tot_ud >> 16;
        }
        if (nbad) atomicAdd(a.counters + kCntMalformedIv, nbad);
        __syncthreads();
        if (tid == 0) {
            const uint32_t base = s_v[2], n_up = ups, U0 = s_v[0], Dl = s_v[1];
            if (s_v[3]) {  // not this kernel's business (see process_batch_t): literal_kernel computes the row
                w.lit_list[atomicAdd(a.counters + kCntLiteralList, 1u)] = r;
            } else if ((uint64_t)base + k + 1u <= w.stage_cap) {
                uint32_t *P = reinterpret_cast<uint32_t *>(w.stage + base);
                P[0] = 0u;
                P[2u * n_up + 1u] = len;
                const uint32_t q0 = n_up ? (U0 == 0u) : 0u;
                const uint32_t ng = n_up ? n_up + (Dl != len) - q0 : (len != 0u);
                w.meta[r] = make_uint2(base + q0, ng);
            } else {
                atomicAdd(cnt + kCntStageOverflow, 1u);
                w.meta[r] = make_uint2(0u, 0u);
            }
        }
        __syncthreads();  // s_v is reused by the next row
    }
}

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

// m = 2 m + (ev <= q): the compare is the carry of q - ev, pushed in with an add-with-carry (SASS: IADD3 + IMAD.X)
__device__ __forceinline__ uint32_t push_le(uint32_t m, uint32_t ev, uint32_t q) {
    uint32_t r;
    asm("{\n.reg .u32 t;\nsub.cc.u32 t, %1, %2;\naddc.u32 %0, %3, %3;\n}" : "=r"(r) : "r"(q), "r"(ev), "r"(m));
    return r;
}

#include "regtier.cuh"

// The class a warp's batch indices currently fall in, held in registers: indices drawn by a warp only grow, so the class
// table (kernel parameter, i.e. constant bank with a run-time index: a chain of slow loads) is read once per class, not
// once per batch.
struct ClassCursor {
    uint32_t q, lo, hi;  // position in processing order; batches [lo, hi) belong to it
    uint32_t cls, G, rpb, inv, ebase, count;
};
__device__ __forceinline__ void cursor_load(const ClassTab &tab, ClassCursor &cu) {
    cu.lo = tab.item_base[cu.q];
    cu.hi = tab.item_base[cu.q + 1];
    cu.cls = tab.order[cu.q];
    cu.G = tab.lanes[cu.cls];
    cu.rpb = tab.rpb[cu.cls];
    cu.inv = tab.inv[cu.cls];
    cu.ebase = tab.entry_base[cu.cls];
    cu.count = tab.count[cu.cls];
}

// Starts the copy of this lane's record of batch `item` (the record of the row its group sorts) into the warp's
// shared-memory slot: cp.async, global -> shared without a register in between, so nothing has to stay live (or be
// spilled, which would wait for the load) across the batch being sorted. Lanes without a row get a zero record. Beside
// it goes the lane's place in the batch for issue_batch: where its row's slab starts in the buffer, and (top bit)
// whether this lane is the first of its group.
__device__ __forceinline__ void fetch_rec(const Work &w, const ClassTab &tab, ClassCursor &cu, uint4 *slot, uint32_t *geo_slot, uint32_t item,
                                          uint32_t n_items, uint32_t &cls, uint32_t lane) {
    cls = 0;
    const uint4 *src = nullptr;
    uint32_t geo = 0;
    if (item < n_items) {
        while (item >= cu.hi) {
            ++cu.q;
            cursor_load(tab, cu);
        }
        cls = cu.cls;
        const uint32_t j = (lane * cu.inv) >> 16, g = lane - j * cu.G;  // lane / G, lane % G
        if (j < cu.rpb) {
            const uint32_t e = (item - cu.lo) * cu.rpb + j;
            if (e < cu.count) src = w.recs + cu.ebase + e;
            geo = j * ((uint32_t)kE * cu.G + 2u) | (g == 0u ? 0x80000000u : 0u);
        }
    }
    geo_slot[lane] = geo;
    if (src) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(slot + lane)), "l"(src) : "memory");
    else slot[lane] = make_uint4(0, 0, 0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void fetch_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// TMA copies of the batch's row slabs into `buf` (one per row, issued by the group's first lane).
__device__ __forceinline__ void issue_batch(const DetectArgs &a, uint2 *buf, unsigned long long *bar, const uint4 rec, const uint32_t geo,
                                            uint32_t lane) {
    uint32_t bytes = 0, cs = 0;
    if ((geo >> 31) && (rec.z & kRecValid)) {
        cs = rec.y & ~1u;
        bytes = (((rec.y + (rec.z & 0xFFFFu) + 1u) & ~1u) - cs) * 8u;
    }
    const uint32_t total = __reduce_add_sync(FULL, bytes);
    if (lane == 0) mbar_expect_tx(bar, total);
    __syncwarp();
    if (bytes) tma_load_1d(buf + (geo & 0x7FFFFFFFu), a.iv + cs, bytes, bar);
}

template <bool VAL>
__global__ void __launch_bounds__(kSortThreads, 1) sort_kernel(DetectArgs a, Work w, ClassTab tab, uint32_t c, PipeMul pm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
#ifdef YB_TRACE_CTA
    unsigned long long tr_t0, tr_t1 = 0, tr_t2;
    uint32_t tr_n = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_t0));
#endif
    WarpSmem &ws = *reinterpret_cast<WarpSmem *>(smem_raw + wid * kWarpSmemBytes);
    uint2 *buf = reinterpret_cast<uint2 *>(smem_raw + wid * kWarpSmemBytes + sizeof(WarpSmem));
    if (lane == 0) {
        mbar_init(&ws.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t ep = __ldcg(a.counters + kCntEpoch);
    uint32_t *cnt = a.counters + (ep & 1u) * kNumCounters;
    const uint32_t n_items = tab.item_base[kNumClasses];
    ClassCursor cu;
    cu.q = 0;
    cursor_load(tab, cu);
    // Batches are dealt from one global cursor, in processing order: SMs do not all run at the same speed (a few per
    // cent, measured), and a fixed deal leaves the slowest CTA working alone at the end. The cursor is a double: ptxas
    // turns an integer atomicAdd of one lane into its warp-aggregated form, whose shuffle needs the result at once, i.e. a
    // stall of an L2 round trip per batch; the f64 add (ATOMG.E.ADD.F64) is left alone, and its result is first touched
    // when the batch's keys are in registers (refill, below).
    double *dyn_cursor = reinterpret_cast<double *>(cnt + kCntDynTicket);
    auto draw_raw = [&]() {  // the warp's next batch (lane 0 holds it; indices drawn by a warp only grow)
        double d = 0.0;
        if (lane == 0) d = atomicAdd(dyn_cursor, 1.0);
        return d;
    };
    // A warp's first three batches are fixed: batch w * grid + b of each of the first three rounds for warp w of CTA b,
    // so that the heaviest batches (the processing order starts with them) are spread evenly over the SMs and no
    // start-up atomics are needed; the cursor deals everything behind them. (Drawing the three from the cursor back to
    // back gave a warp three NEIGHBOURING heavy batches: +22 us on a 1/8 shard of C3.)
    const uint32_t n_round = gridDim.x * kSortWarps, first = wid * gridDim.x + blockIdx.x;
    const double n_fixed = (double)kFixedRounds * (double)n_round;
    auto to_item = [&](double d) {
        d += n_fixed;
        return d < (double)n_items ? __double2uint_rz(d) : n_items;
    };
    uint32_t item = min(first, n_items), item1, item2;
    if (kFixedRounds == 3u) {
        item1 = min(n_round + first, n_items), item2 = min(2u * n_round + first, n_items);
    } else {  // (one fixed round: the other two from the cursor, one after the other)
        item1 = __shfl_sync(FULL, to_item(draw_raw()), 0);
        item2 = __shfl_sync(FULL, to_item(draw_raw()), 0);
    }
    // software pipeline: the record of batch i+2 is on its way to shared memory, the slab copies of batch i+1 are issued
    // as soon as the keys of batch i are in registers (one slab buffer) and land while batch i is sorted
    uint32_t cls0, cls1, cls2, s = 1;  // ws.rec[s]: record of batch i+1; ws.rec[s ^ 1]: of batch i+2
    fetch_rec(w, tab, cu, ws.rec[0], ws.geo[0], item, n_items, cls0, lane);
    fetch_rec(w, tab, cu, ws.rec[1], ws.geo[1], item1, n_items, cls1, lane);
    fetch_wait();
    uint4 rec0 = ws.rec[0][lane];
    if (item < n_items) issue_batch(a, buf, &ws.mbar, rec0, ws.geo[0][lane], lane);
    fetch_rec(w, tab, cu, ws.rec[0], ws.geo[0], item2, n_items, cls2, lane);
    uint32_t parity = 0;
    uint2 chunk = make_uint2(0, 0);  // [next free pair, end) of the warp's staging chunk
    while (item < n_items) {
        const double raw3 = draw_raw();  // first touched in refill, consumed at the end of this iteration
        // the loop's state goes to shared memory for the time the batch is sorted and comes back after (see WarpSmem::st)
        if (lane == 0) {
            ws.st[0] = item1, ws.st[1] = item2, ws.st[2] = cls1, ws.st[3] = cls2, ws.st[4] = s, ws.st[5] = parity;
            ws.st[6] = cu.q, ws.st[7] = cu.lo, ws.st[8] = cu.hi, ws.st[9] = cu.cls, ws.st[10] = cu.G, ws.st[11] = cu.rpb;
            ws.st[12] = cu.inv, ws.st[13] = cu.ebase, ws.st[14] = cu.count;
        }
        mbar_wait(&ws.mbar, parity);
        fetch_wait();  // (issued a batch ago)
        auto refill = [&]() {
            // this batch's generic-proxy reads of the slab before the async-proxy writes of the next batch's copies
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (item1 < n_items) issue_batch(a, buf, &ws.mbar, ws.rec[s][lane], ws.geo[s][lane], lane);
            if (lane == 0) ws.st[15] = to_item(raw3);  // (the keys are in registers: the draw's L2 round trip is over, or nearly)
        };
#define YB_CASE(gi)                                                                                                        \
    case gi: process_batch_t<class_lanes_c(gi), true, VAL>(a, w, cnt, ws, buf, rec0, c, chunk, pm, lane, refill); break;        \
    case gi + kNumG: process_batch_t<class_lanes_c(gi), false, VAL>(a, w, cnt, ws, buf, rec0, c, chunk, pm, lane, refill); break;
        switch (cls0) {  // classes 0 .. kNumG-1: packed rows, G = class_lanes(class); kNumG .. : the same sizes for long reads
            YB_CASE(0) YB_CASE(1) YB_CASE(2) YB_CASE(3) YB_CASE(4) YB_CASE(5) YB_CASE(6)
#if YB_KE == 16
            YB_CASE(7) YB_CASE(8) YB_CASE(9)
#endif
            default: break;
        }
#undef YB_CASE
        __syncwarp();
#ifdef YB_TRACE_CTA
        if (tr_n++ == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_t1));
#endif
        s = ws.st[4];
        rec0 = ws.rec[s][lane];
        item = ws.st[0], item1 = ws.st[1], cls0 = ws.st[2], cls1 = ws.st[3], parity = ws.st[5];
        cu.q = ws.st[6], cu.lo = ws.st[7], cu.hi = ws.st[8], cu.cls = ws.st[9], cu.G = ws.st[10], cu.rpb = ws.st[11];
        cu.inv = ws.st[12], cu.ebase = ws.st[13], cu.count = ws.st[14];
        item2 = ws.st[15];
        fetch_rec(w, tab, cu, ws.rec[s], ws.geo[s], item2, n_items, cls2, lane);
        s ^= 1u;
        parity ^= 1u;
    }
#ifdef YB_TRACE_CTA
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_t2));
    if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 73 || blockIdx.x == 147))
        printf("TRACE val=%d cta %u warp %u t0 %llu first %llu end %llu batches %u\n", (int)VAL, blockIdx.x, wid, tr_t0, tr_t1 ? tr_t1 - tr_t0 : 0ull,
               tr_t2 - tr_t0, tr_n);
#endif
}

constexpr size_t kSortSmemBytes = kWarpSmemBytes * kSortWarps;

// ------------------------------------------------------------------------------------------------
// ordering pass (order_kernel, the second and last kernel of a detect step): the sorting kernels left, per row, a
// staging offset and a count. A CTA takes one part of 2048 rows (in ticket order), 4 consecutive rows per thread: it
// publishes the part's total in one 64-bit word tagged with the step number (never reset) as soon as its counts are
// loaded, sums the totals of ALL parts before it (a thousand words at most; it only ever waits for parts that are
// already running and whose total does not depend on anybody: no scan kernel, no look-back chain, and no RED per row
// in the sorting kernels, which serialised in L2 when every warp of the GPU worked on the same stretch of rows), turns
// counts into offsets (thread-local prefix + block scan), moves the staged regions to their final place, classifies
// (editor/mod.rs:85-100) and writes the 2-bit bitmap (to every peer's gather buffer when the all-gather is fused in).
// The last CTA to finish closes the step: it zeroes the other counter set, bumps the step number and, with peers, tells
// every rank that this rank's slot is complete.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

#ifndef YB_ORDER_MIN_CTAS
#define YB_ORDER_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(kOrderThreads, YB_ORDER_MIN_CTAS) order_kernel(DetectArgs a, Work w, double not_cov) {
    constexpr uint32_t R = kOrderRows, NW = kOrderThreads / 32;
    __shared__ uint32_t s_warp[NW], s_pre[NW], s_hist[NW], s_last, s_part;
    __shared__ __align__(16) uint32_t s_off[kPartRows], s_src[kPartRows], s_len[kPartRows], s_bad[kPartRows];
    __shared__ __align__(16) uint8_t s_int[kPartRows], s_map[NW][kOrderMap];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t ep = __ldcg(a.counters + kCntEpoch);
    uint32_t *cnt = a.counters + (ep & 1u) * kNumCounters;
    if (tid == 0) s_part = atomicAdd(cnt + kCntTicket, 1u);  // parts start in ticket order: a part only waits for parts that run
    uint32_t peer_step = 0;
    if (a.n_peers) {
        // every rank has finished step peer_step - 1 (and, in its stream order, whatever read the gather buffer of step
        // peer_step - 2, which this step overwrites) before this rank stores into the peers' buffers. Bounded wait.
        peer_step = __ldcg(a.peer_flag[a.rank] + 31);
        if (tid < a.n_peers) {
            const uint32_t *mine = a.peer_flag[a.rank] + tid;
            uint32_t seen = 0;
            for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
                if ((int32_t)(seen - peer_step) >= 0) break;
                __nanosleep(200);
            }
            if ((int32_t)(seen - peer_step) < 0) atomicAdd(cnt + kCntPeerTimeout, 1u);
        }
    }
    __syncthreads();
    const uint32_t part = s_part, r0 = part * kPartRows + tid * R;
    uint2 m[R];
    uint32_t l[R];
    const bool full = r0 + R <= a.n_reads;
    if (full) {
        const uint4 *mp = reinterpret_cast<const uint4 *>(w.meta + r0);
#pragma unroll
        for (uint32_t i = 0; i < R / 2; ++i) {
            const uint4 x = mp[i];
            m[2 * i] = make_uint2(x.x, x.y);
            m[2 * i + 1] = make_uint2(x.z, x.w);
        }
        const uint4 x = __ldg(reinterpret_cast<const uint4 *>(a.len + r0));
        l[0] = x.x, l[1] = x.y, l[2] = x.z, l[3] = x.w;
    } else {
#pragma unroll
        for (uint32_t i = 0; i < R; ++i) {
            const bool live = r0 + i < a.n_reads;
            m[i] = live ? w.meta[r0 + i] : make_uint2(0u, 0u);
            l[i] = live ? __ldg(a.len + r0 + i) : 0u;
        }
    }
    uint32_t mine = 0;
#pragma unroll
    for (uint32_t i = 0; i < R; ++i) mine += m[i].y;
    const uint32_t incl = warp_incl_scan(mine);
    if (lane == 31u) s_warp[wid] = incl;
    {  // the rows' places in the part, for the threads that will move their regions
        uint32_t o = incl - mine;  // within the warp; the warps before this one are added by the reader (s_wbase)
        *reinterpret_cast<uint4 *>(s_src + tid * R) = make_uint4(m[0].x, m[1].x, m[2].x, m[3].x);
        *reinterpret_cast<uint4 *>(s_len + tid * R) = make_uint4(l[0], l[1], l[2], l[3]);
        *reinterpret_cast<uint4 *>(s_bad + tid * R) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint32_t *>(s_int + tid * R) = 0u;
        uint4 oo;
        oo.x = o, o += m[0].y;
        oo.y = o, o += m[1].y;
        oo.z = o, o += m[2].y;
        oo.w = o;
        *reinterpret_cast<uint4 *>(s_off + tid * R) = oo;
    }
    __syncthreads();
    const unsigned long long tag = (unsigned long long)(ep + 1u) << 32;
    uint32_t wbase = 0, tot = 0;  // regions of the warps before this one, of the part
#pragma unroll
    for (uint32_t q = 0; q < NW; ++q) {
        wbase += q < wid ? s_warp[q] : 0u;
        tot += s_warp[q];
    }
    if (tid == 0) st_desc(w.part_desc + part, tag | tot);  // the part's total, for the parts behind this one
    uint32_t pre = 0;  // totals of the parts before this one (they hold earlier tickets: running or done)
    for (uint32_t i0 = tid; i0 < part; i0 += 8u * kOrderThreads) {  // eight loads in flight per thread, the stragglers polled
        unsigned long long d[8];
#pragma unroll
        for (uint32_t u = 0; u < 8u; ++u) d[u] = i0 + u * kOrderThreads < part ? ld_desc(w.part_desc + i0 + u * kOrderThreads) : tag;
#pragma unroll
        for (uint32_t u = 0; u < 8u; ++u) {
            while ((d[u] >> 32) != (tag >> 32)) d[u] = ld_desc(w.part_desc + i0 + u * kOrderThreads);
            pre += (uint32_t)d[u];
        }
    }
    pre = __reduce_add_sync(FULL, pre);
    if (lane == 0u) s_pre[wid] = pre;
    __syncthreads();
    uint32_t gp0 = 0;  // where the part's regions start in the output
#pragma unroll
    for (uint32_t q = 0; q < NW; ++q) gp0 += s_pre[q];
    if (tid == kOrderThreads - 1u && part == w.n_parts - 1u) a.gap_ptr[a.n_reads] = gp0 + tot;
    // The regions move from the staging buffer to their final place warp by warp (a warp owns 128 consecutive rows and so
    // a contiguous stretch of the output): one region per lane and turn, consecutive lanes writing consecutive regions
    // whatever row they belong to, every load of a turn in flight together. (A thread walking its own rows' regions
    // waits a DRAM latency per region of its longest row with the rest of its warp idle.) The row of a region comes
    // from a byte map the warp fills first (row owners write their rows' entries: two stores for most rows), or, when
    // the warp's rows have more regions than the map holds, from a bisection over the rows' offsets.
    {
        const uint32_t wrow = wid * 32u * R, wtot = s_warp[wid], gpw = gp0 + wbase;
        const uint32_t *so = s_off + wrow, *ssrc = s_src + wrow, *slen = s_len + wrow;
        uint32_t *sbad = s_bad + wrow;
        uint8_t *sint = s_int + wrow, *mp = s_map[wid];
        const bool mapped = wtot <= kOrderMap;
        if (mapped) {
            uint32_t o = incl - mine, longrows = 0;
#pragma unroll
            for (uint32_t i = 0; i < R; ++i) {
                const uint32_t n = m[i].y, rl = lane * R + i;
                if (n > 0u) mp[o] = (uint8_t)rl;
                if (n > 1u) mp[o + 1u] = (uint8_t)rl;
                if (n > 2u) {
                    if (n <= 34u) {
                        for (uint32_t gq = 2; gq < n; ++gq) mp[o + gq] = (uint8_t)rl;
                    } else {
                        longrows |= 1u << i;
                    }
                }
                o += n;
            }
            if (__any_sync(FULL, longrows != 0u)) {  // rows with many regions: the whole warp fills their entries
                uint32_t oo = incl - mine;
#pragma unroll
                for (uint32_t i = 0; i < R; ++i) {
                    uint32_t todo = __ballot_sync(FULL, (longrows >> i) & 1u);
                    while (todo) {
                        const uint32_t sl = __ffs(todo) - 1u;
                        todo &= todo - 1u;
                        const uint32_t ob = __shfl_sync(FULL, oo, sl), nb = __shfl_sync(FULL, m[i].y, sl);
                        for (uint32_t gq = 2u + lane; gq < nb; gq += 32u) mp[ob + gq] = (uint8_t)(sl * R + i);
                    }
                    oo += m[i].y;
                }
            }
            __syncwarp();
        }
        for (uint32_t e0 = lane; e0 < wtot; e0 += kOrderUnroll * 32u) {
            uint32_t row[kOrderUnroll];
            uint2 v[kOrderUnroll];
#pragma unroll
            for (uint32_t u = 0; u < kOrderUnroll; ++u) {
                const uint32_t e = e0 + u * 32u;
                row[u] = 0;
                if (e < wtot) {
                    uint32_t lo = 0;
                    if (mapped) {
                        lo = mp[e];
                    } else {  // last row whose offset is <= e
#pragma unroll
                        for (uint32_t h = 16u * R; h > 0u; h >>= 1)
                            if (so[lo + h] <= e) lo += h;
                    }
                    row[u] = lo;
                    v[u] = w.stage[ssrc[lo] + (e - so[lo])];
                }
            }
#pragma unroll
            for (uint32_t u = 0; u < kOrderUnroll; ++u) {
                const uint32_t e = e0 + u * 32u;
                if (e < wtot) {
                    a.gaps[gpw + e] = v[u];
                    atomicAdd(sbad + row[u], v[u].y - v[u].x);
                    if (v[u].x != 0u && v[u].y != slen[row[u]]) sint[row[u]] = 1;
                }
            }
        }
        __syncwarp();
    }
    uint32_t gp = gp0 + wbase + incl - mine;
    const NotCovBand band = not_cov_band(not_cov);
    uint32_t off[R], cl[R], h1 = 0, h2 = 0, bits = 0;
#pragma unroll
    for (uint32_t i = 0; i < R; ++i) {
        off[i] = gp;
        gp += m[i].y;
        cl[i] = r0 + i < a.n_reads ? classify_fast(s_bad[tid * R + i], l[i], s_int[tid * R + i] ? 2u : 0u, not_cov, band) : 0u;
        h1 += cl[i] == 1u;
        h2 += cl[i] == 2u;
        bits |= cl[i] << (2u * i);
    }
    if (full) {
        *reinterpret_cast<uint4 *>(a.gap_ptr + r0) = make_uint4(off[0], off[1], off[2], off[3]);
        *reinterpret_cast<uint32_t *>(a.cls + r0) = cl[0] | cl[1] << 8 | cl[2] << 16 | cl[3] << 24;
    } else {
#pragma unroll
        for (uint32_t i = 0; i < R; ++i)
            if (r0 + i < a.n_reads) {
                a.gap_ptr[r0 + i] = off[i];
                a.cls[r0 + i] = (uint8_t)cl[i];
            }
    }
    // four threads (16 rows) make one 32-bit word of the 2-bit bitmap
    uint32_t word = bits << (8u * (lane & 3u));
    word |= __shfl_xor_sync(FULL, word, 1);
    word |= __shfl_xor_sync(FULL, word, 2);
    if ((lane & 3u) == 0u && r0 < a.n_reads) {
        if (a.n_peers == 0u) {
            reinterpret_cast<uint32_t *>(a.bitmap)[r0 >> 4] = word;
        } else {  // all-gather fused into the epilogue: the word goes to this rank's slot on every rank (NVLink stores);
                  // two slots per rank, alternating with the step, so a reader of step s never sees stores of step s + 1
            const size_t po = (size_t)(peer_step & 1u) * a.peer_parity_bytes;
            for (uint32_t p = 0; p < a.n_peers; ++p) reinterpret_cast<uint32_t *>(a.peer_slot[p] + po)[r0 >> 4] = word;
            // (made visible system-wide by ONE fence of thread 0 behind the CTA barrier below, the grid-sync pattern:
            // a fence per storing thread kept every CTA alive for 64 NVLink round trips)
        }
    }
    // class histogram: one RED per class per CTA, spread over kHistSlots copies (same-address atomics serialise in L2)
    const uint32_t hh = __reduce_add_sync(FULL, h1 | (h2 << 16));
    if (lane == 0) s_hist[wid] = hh;
    __syncthreads();
    if (tid == 0) {
        uint32_t hs = 0;
#pragma unroll
        for (uint32_t q = 0; q < NW; ++q) hs += s_hist[q];
        const uint32_t n_live = min(kPartRows, a.n_reads - part * kPartRows);
        uint32_t *slot = cnt + kCntHist + 3u * (part % kHistSlots);
        const uint32_t c1 = hs & 0xFFFFu, c2 = hs >> 16;
        if (n_live - c1 - c2) atomicAdd(slot + 0, n_live - c1 - c2);
        if (c1) atomicAdd(slot + 1, c1);
        if (c2) atomicAdd(slot + 2, c2);
        if (a.n_peers) __threadfence_system();  // the CTA's peer stores (ordered before this thread by the barrier) before kCntDone
        s_last = atomicAdd(cnt + kCntDone, 1u) == w.n_parts - 1u;
    }
    __syncthreads();
    if (s_last) {  // the step is complete: next step's counter set, step number, peers
        uint32_t *other = a.counters + ((ep + 1u) & 1u) * kNumCounters;
        for (uint32_t i = tid; i < kNumCounters; i += kOrderThreads) other[i] = 0u;
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            a.counters[kCntEpoch] = ep + 1u;
            if (a.validate) {  // what this step's validation found, for the host; the running counts start over
                a.counters[kCntLiteralLast] = a.counters[kCntLiteralList];
                a.counters[kCntMalformedLast] = a.counters[kCntMalformedIv];
                a.counters[kCntLiteralList] = 0u;
                a.counters[kCntMalformedIv] = 0u;
            }
        }
        if (a.n_peers) {
            __threadfence_system();  // every part's peer stores (fenced by its thread 0 before kCntDone) before the flags
            if (tid < a.n_peers) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_flag[tid] + a.rank), "r"(peer_step + 1u) : "memory");
            if (tid == 0) a.peer_flag[a.rank][31] = peer_step + 1u;
        }
    }
}

// Consumer side of the fused all-gather: returns (on the stream) once every rank's slot of the LAST finished step is
// complete in this rank's gather buffer. Bounded wait: a missing rank must not hang the GPU.
__global__ void __launch_bounds__(32) peer_wait_kernel(DetectArgs a) {
    const uint32_t p = threadIdx.x;
    if (p >= a.n_peers) return;
    const uint32_t step = __ldcg(a.peer_flag[a.rank] + 31);
    const uint32_t *mine = a.peer_flag[a.rank] + p;
    uint32_t seen = 0;
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        if ((int32_t)(seen - step) >= 0) break;
        __nanosleep(100);
    }
    if ((int32_t)(seen - step) < 0) atomicAdd(a.counters + kCntPeerTimeoutWait, 1u);
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

// ------------------------------------------------------------------------------------------------
// row_stats_kernel (upload time): size-class histogram, big-row scratch needs, input sanity
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) row_stats_kernel(const uint32_t *__restrict__ rowptr, const uint32_t *__restrict__ len,
                                                          uint32_t n_reads, DevRowStats *out) {
    __shared__ uint32_t s_cnt[kNumClasses + 1], s_max, s_bad[3];
    (void)0;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, r = blockIdx.x * 1024u + tid;
    if (tid <= (uint32_t)kNumClasses) s_cnt[tid] = 0u;
    if (tid < 3u) s_bad[tid] = 0u;
    if (tid == 0) s_max = 0u;
    __syncthreads();
    int cls = -2;
    uint32_t k = 0;
    if (r < n_reads) {
        const uint32_t p0 = __ldg(rowptr + r), p1 = __ldg(rowptr + r + 1), l = __ldg(len + r);
        if (p1 < p0) {
            atomicAdd(&s_bad[0], 1u);
        } else {
            k = p1 - p0;
            cls = class_of_row(k, l);
            if (cls < 0) {  // big row: rare, straight to the global sums
                cls = kNumClasses;
                atomicAdd(&out->big_pairs, (unsigned long long)k + 1ull);
                if (big_row_scans(k, l)) {
                    atomicAdd(&out->n_scan, 1u);
                    atomicMax(&out->max_len_scan, l);
                } else {
                    atomicMax(&out->max_k_sort, k);
                    const unsigned long long hk = cta_words(k, l > kPackedMaxLen);
                    if (hk > kCtaMaxSmemWords) atomicAdd(&out->huge_keys, hk);
                }
            }
            if (l > kPackedMaxLen) atomicAdd(&s_bad[2], 1u);
        }
        if (l > kMaxLength) atomicAdd(&s_bad[1], 1u);
    }
    const uint32_t peers = __match_any_sync(FULL, cls);
    if (cls >= 0 && lane == (uint32_t)__ffs(peers) - 1u) atomicAdd(&s_cnt[cls], (uint32_t)__popc(peers));
    uint32_t mk = k;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mk = max(mk, __shfl_xor_sync(FULL, mk, off));
    if (lane == 0 && mk) atomicMax(&s_max, mk);
    __syncthreads();
    if (tid < (uint32_t)kNumClasses && s_cnt[tid]) atomicAdd(&out->class_count[tid], s_cnt[tid]);
    if (tid == 0) {
        if (s_cnt[kNumClasses]) atomicAdd(&out->n_big, s_cnt[kNumClasses]);
        if (s_max) atomicMax(&out->max_k, s_max);
        if (s_bad[0]) atomicAdd(&out->bad_rowptr, s_bad[0]);
        if (s_bad[1]) atomicAdd(&out->bad_len, s_bad[1]);
        if (s_bad[2]) atomicAdd(&out->n_wide, s_bad[2]);
    }
}

// literal_kernel: the reference's algorithm itself (stack.rs:61-139: sort, min-heap sweep of the ends, head / tail
// regions, merge of regions that share a begin), one thread per row, for the rows the validating kernels listed. It runs after
// the sorting kernels and replaces what they staged for those rows. The row is sorted in place in the device copy of the
// interval buffer (heapsort); the heap of ends and the regions live in a segment of the staging buffer.
__device__ __forceinline__ bool lit_less(uint2 x, uint2 y) { return x.x != y.x ? x.x < y.x : x.y < y.y; }
__device__ void lit_heapsort(uint2 *v, uint32_t n) {  // ovls.sort_unstable() (stack.rs:66)
    auto sift = [&](uint32_t i, uint32_t m) {
        for (;;) {
            uint32_t big = i;
            const uint32_t lc = 2u * i + 1u, rc = lc + 1u;
            if (lc < m && lit_less(v[big], v[lc])) big = lc;
            if (rc < m && lit_less(v[big], v[rc])) big = rc;
            if (big == i) return;
            const uint2 t = v[i];
            v[i] = v[big];
            v[big] = t;
            i = big;
        }
    };
    for (uint32_t i = n / 2u; i-- > 0u;) sift(i, n);
    for (uint32_t m = n; m > 1u; --m) {
        const uint2 t = v[0];
        v[0] = v[m - 1u];
        v[m - 1u] = t;
        sift(0u, m - 1u);
    }
}
__device__ __forceinline__ void lit_push(uint32_t *h, uint32_t &n, uint32_t x) {  // BinaryHeap<Reverse<u32>>::push
    uint32_t i = n++;
    while (i > 0u) {
        const uint32_t p = (i - 1u) >> 1;
        if (h[p] <= x) break;
        h[i] = h[p];
        i = p;
    }
    h[i] = x;
}
__device__ __forceinline__ void lit_pop(uint32_t *h, uint32_t &n) {
    const uint32_t x = h[--n];
    uint32_t i = 0;
    for (;;) {
        uint32_t ch = 2u * i + 1u;
        if (ch >= n) break;
        if (ch + 1u < n && h[ch + 1u] < h[ch]) ++ch;
        if (x <= h[ch]) break;
        h[i] = h[ch];
        i = ch;
    }
    if (n) h[i] = x;
}

__global__ void __launch_bounds__(64) literal_kernel(DetectArgs a, Work w, uint32_t coverage) {
    const uint32_t n_lit = __ldcg(a.counters + kCntLiteralList);  // what the validating kernels of this step listed
    const uint32_t ep = __ldcg(a.counters + kCntEpoch);
    uint32_t *cnt = a.counters + (ep & 1u) * kNumCounters;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_lit; j += gridDim.x * blockDim.x) {
        const uint32_t r = w.lit_list[j];
        const uint32_t s = a.rowptr[r], k = a.rowptr[r + 1] - s, len = a.len[r];
        uint2 *ovls = const_cast<uint2 *>(a.iv) + s;
        lit_heapsort(ovls, k);
        // staging segment: slot 0 for the head region, then at most k regions, then the heap (k words)
        const uint32_t pairs = k + 2u + (k + 1u) / 2u;
        const uint32_t at = atomicAdd(cnt + kCntStage, pairs);
        if ((uint64_t)at + pairs > w.stage_cap) {
            atomicAdd(cnt + kCntStageOverflow, 1u);
            w.meta[r] = make_uint2(0u, 0u);
            continue;
        }
        uint2 *raw = w.stage + at + 1u;
        uint32_t *heap = reinterpret_cast<uint32_t *>(w.stage + at + k + 2u);
        uint32_t n_raw = 0, hn = 0, first_covered = 0, last_covered = 0;  // stack.rs:62-69
        for (uint32_t i = 0; i < k; ++i) {                               // stack.rs:71
            const uint2 iv = ovls[i];
            while (hn > 0u) {                                            // stack.rs:72
                const uint32_t head = heap[0];
                if (head > iv.x) break;                                  // stack.rs:73-75
                if (hn > coverage) last_covered = head;                  // stack.rs:77-79
                lit_pop(heap, hn);                                       // stack.rs:80
            }
            if (hn <= coverage) {                                        // stack.rs:83
                if (last_covered != 0u) raw[n_raw++] = make_uint2(last_covered, iv.x);  // stack.rs:84-85
                else first_covered = iv.x;                               // stack.rs:87
            }
            lit_push(heap, hn, iv.y);                                    // stack.rs:90
        }
        while (hn > coverage) {                                          // stack.rs:93
            last_covered = heap[0];
            if (last_covered >= len) break;                              // stack.rs:101-103
            lit_pop(heap, hn);
        }
        uint2 *g = raw;
        if (first_covered != 0u) {                                       // stack.rs:107-109
            --g;
            g[0] = make_uint2(0u, first_covered);
            ++n_raw;
        }
        if (last_covered != len) g[n_raw++] = make_uint2(last_covered, len);  // stack.rs:111-113
        uint32_t n_clean = 0;
        if (n_raw) {                                                     // stack.rs:119-138, in place
            uint2 cur = g[0];
            for (uint32_t i = 0; i + 1u < n_raw; ++i) {
                const uint2 g1 = g[i], g2 = g[i + 1u];
                if (g1.x == g2.x) {
                    cur = make_uint2(g1.x, max(g1.y, g2.y));
                } else {
                    g[n_clean++] = cur;
                    cur = g2;
                }
            }
            g[n_clean++] = cur;
        }
        w.meta[r] = make_uint2((uint32_t)(g - w.stage), n_clean);
    }
}

Work carve(const DetectArgs &a, uint64_t huge_keys, uint64_t n_big, size_t *total) {
    Work w;
    size_t off = 0;
    char *base = static_cast<char *>(a.scratch);
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off += align256(bytes);
        return p;
    };
    w.n_parts = (a.n_reads + kPartRows - 1u) / kPartRows;
    w.recs = reinterpret_cast<uint4 *>(take(sizeof(uint4) * ((size_t)a.n_reads + 1)));
    w.meta = reinterpret_cast<uint2 *>(take(sizeof(uint2) * ((size_t)a.n_reads + 1)));
    // worst case of the bad regions (k + 1 per row) + half of it for chunk remainders (a remainder is only dropped
    // for a batch smaller than a quarter chunk, and a fresh chunk always holds at least two such batches)
    // (a row redone by literal_kernel takes 1.5 k + 3 pairs there and nothing in the sorting kernels)
    const uint64_t cap = (uint64_t)a.n_iv + a.n_reads + ((uint64_t)a.n_iv + a.n_reads) / 2 + 2ull * a.n_reads + 4096ull * kStageChunk;  // + one open chunk per resident warp
    w.stage_cap = cap > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)cap;
    w.stage = reinterpret_cast<uint2 *>(take(sizeof(uint2) * ((size_t)w.stage_cap + 1)));
    w.part_desc = reinterpret_cast<unsigned long long *>(take(sizeof(unsigned long long) * ((size_t)w.n_parts + 8)));
    w.big_list = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * (n_big + 1)));
    w.scan_list = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * (n_big + 1)));
    w.huge_keys = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * (huge_keys + 1)));
    w.lit_list = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * ((size_t)a.n_reads + 1)));
    *total = off;
    return w;
}

// The step's class table: where every class's records sit in the worklist and how its batches are numbered.
ClassTab make_plan(const DetectArgs &a) {
    ClassTab tab{};
    uint32_t at = 0;
    for (int cl = 0; cl < kNumClasses; ++cl) {
        tab.entry_base[cl] = at;
        tab.count[cl] = a.rows.class_count[cl];
        at += tab.count[cl];
    }
    // processing order: largest groups first (the cheap batches last make a fine-grained tail); the few long reads (their
    // code is cold and they refill late) go behind the first packed class, in the middle of everybody's work
    int seq[kNumClasses], n_seq = 0;
    seq[n_seq++] = kNumG - 1;
    for (int gi = kNumG - 1; gi >= 0; --gi) seq[n_seq++] = gi + kNumG;
    for (int gi = kNumG - 2; gi >= 0; --gi) seq[n_seq++] = gi;
    uint32_t items = 0;
    for (int q = 0; q < kNumClasses; ++q) {
        const int cl = seq[q], gi = cl % kNumG;
        const uint32_t G = class_lanes(gi), rpb = std::min(32u / G, kBufIntervals / ((uint32_t)kE * G + 2u));
        tab.lanes[cl] = G;
        tab.rpb[cl] = rpb;
        tab.inv[cl] = (65536u + G - 1u) / G;
        tab.order[q] = (uint32_t)cl;
        tab.item_base[q] = items;
        items += (tab.count[cl] + rpb - 1u) / rpb;
    }
    tab.item_base[kNumClasses] = items;
    return tab;
}

// Per-device launch constants (cudaFuncSetAttribute applies to the current device only; one context per GPU and per
// thread is a supported way to use the library).
struct DevCfg {
    std::once_flag once;
    int ok = 0, n_sm = 0, occ_sort = 0;
};
DevCfg g_dev[64];

const DevCfg *dev_cfg() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    DevCfg &d = g_dev[dev];
    std::call_once(d.once, [&]() {
        int sm = 0, occ = 0;
        if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return;
        if (cudaFuncSetAttribute(big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kCtaMaxSmemWords * sizeof(uint32_t))) != cudaSuccess) return;
        if (cudaFuncSetAttribute(bigscan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanSmemBytes) != cudaSuccess) return;
        if (cudaFuncSetAttribute(sort_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmemBytes) != cudaSuccess) return;
        if (cudaFuncSetAttribute(sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmemBytes) != cudaSuccess) return;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sort_kernel<true>, kSortThreads, kSortSmemBytes) != cudaSuccess || occ < 1) return;
        d.n_sm = sm;
        d.occ_sort = occ;
        d.ok = 1;
    });
    if (!d.ok) {
        cudaGetLastError();
        return nullptr;
    }
    return &d;
}

}  // namespace

// The same statistics from the host copy of the row pointers (their differences only: a chunk's pointers need not start
// at 0). The streamed path uses it: a chunk's upload then has no device round trip in it, and the H2D engine no bubble.
void host_row_stats(const uint32_t *rowptr, const uint32_t *len, uint32_t n_reads, DevRowStats *out) {
    *out = DevRowStats{};
    int cls_of_slabs[kRegisterTierMaxK / kE + 1];  // class_of_row by ceil(k / kE), for the rows of the register tier
    for (uint32_t q = 0; q <= kRegisterTierMaxK / kE; ++q) cls_of_slabs[q] = class_of_row(q * kE, 0);
    for (uint32_t r = 0; r < n_reads; ++r) {
        const uint32_t p0 = rowptr[r], p1 = rowptr[r + 1], l = len[r];
        if (l > kMaxLength) ++out->bad_len;
        if (p1 < p0) {
            ++out->bad_rowptr;
            continue;
        }
        const uint32_t k = p1 - p0;
        const int cls = k > kRegisterTierMaxK ? -1 : cls_of_slabs[(k + kE - 1) / kE] + (l > kPackedMaxLen ? kNumG : 0);
        if (cls < 0) {
            ++out->n_big;
            out->big_pairs += (unsigned long long)k + 1ull;
            if (big_row_scans(k, l)) {
                ++out->n_scan;
                out->max_len_scan = std::max(out->max_len_scan, l);
            } else {
                out->max_k_sort = std::max(out->max_k_sort, k);
                const unsigned long long hk = cta_words(k, l > kPackedMaxLen);
                if (hk > kCtaMaxSmemWords) out->huge_keys += hk;
            }
        } else {
            ++out->class_count[cls];
        }
        if (l > kPackedMaxLen) ++out->n_wide;
        out->max_k = std::max(out->max_k, k);
    }
}

int launch_row_stats(const uint32_t *rowptr, const uint32_t *len, uint32_t n_reads, DevRowStats *out, cudaStream_t stream) {
    if (cudaMemsetAsync(out, 0, sizeof(DevRowStats), stream) != cudaSuccess) return -1;
    if (n_reads == 0) return 0;
    row_stats_kernel<<<(n_reads + 1023u) / 1024u, 1024, 0, stream>>>(rowptr, len, n_reads, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_upload_kernels(const DetectArgs &a, cudaStream_t stream) {
    // a fresh CSR: both counter sets, the step number and the upload cursors start from zero
    if (cudaMemsetAsync(a.counters, 0, kCounterWords * sizeof(uint32_t), stream) != cudaSuccess) return -1;
    if (a.n_reads == 0) return 0;
    const DevCfg *dc = dev_cfg();
    if (!dc) return -1;
    size_t total = 0;
    Work w = carve(a, a.rows.huge_keys, a.rows.n_big, &total);
    if (total > a.scratch_bytes) return -1;
    if (cudaMemsetAsync(w.part_desc, 0, sizeof(unsigned long long) * (size_t)w.n_parts, stream) != cudaSuccess) return -1;
    scatter_kernel<<<(a.n_reads + kScatterRows - 1u) / kScatterRows, kScatterRows, 0, stream>>>(a, w, make_plan(a));
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

uint64_t huge_keys_for_row(uint64_t k) {
    const uint64_t p = cta_words(k, true);
    return (k > kSmallMaxK && p > kCtaMaxSmemWords) ? p : 0;
}

size_t detect_scratch_bytes(uint32_t n_reads, uint32_t n_iv, const RowStats &rs) {
    DetectArgs a{};
    a.n_reads = n_reads;
    a.n_iv = n_iv;
    size_t total = 0;
    carve(a, rs.huge_keys, rs.n_big, &total);
    return total;
}

int launch_detect(const DetectArgs &a, uint32_t coverage, double not_coverage, cudaStream_t stream) {
    if (a.n_reads == 0) return 0;
    const DevCfg *dc = dev_cfg();
    if (!dc) return -1;
    int launches = 0;
    size_t total = 0;
    Work w = carve(a, a.rows.huge_keys, a.rows.n_big, &total);
    if (total > a.scratch_bytes) return -1;
    const PipeMul pm = {1u, 0xFFFFFFFFu, 65536u};
    bool forked = false;
    if (a.rows.n_big) {
        // the few rows with more than 512 intervals (CTA tier) run on a side stream, beside the register tier (both only
        // append to the staging buffer)
        forked = a.side_stream && a.ev_fork && a.ev_join && cudaEventRecord(a.ev_fork, stream) == cudaSuccess &&
                 cudaStreamWaitEvent(a.side_stream, a.ev_fork, 0) == cudaSuccess;
    }
    {
        // the register tier goes first: its persistent CTAs own a whole SM each and its first batches per warp are dealt
        // by position, so it wants every SM from the start; the CTA tier's kernels deal their rows from a cursor and
        // fill the SMs as the register tier leaves them
        const ClassTab tab = make_plan(a);
        const uint32_t items = tab.item_base[kNumClasses];
        if (items) {
            uint32_t grid = (uint32_t)dc->n_sm;
            if (grid > items) grid = items;
            if (a.validate) sort_kernel<true><<<grid, kSortThreads, kSortSmemBytes, stream>>>(a, w, tab, coverage, pm);
            else sort_kernel<false><<<grid, kSortThreads, kSortSmemBytes, stream>>>(a, w, tab, coverage, pm);
            ++launches;
        }
    }
    if (a.rows.n_scan) {  // short enough reads: begins and ends counted per position, depth scanned window by window
        uint32_t grid = 2u * (uint32_t)dc->n_sm;
        if (grid > a.rows.n_scan) grid = (uint32_t)a.rows.n_scan;
        bigscan_kernel<<<grid, kScanThreads, kScanSmemBytes, forked ? a.side_stream : stream>>>(a, w, coverage);
        ++launches;
    }
    if (a.rows.n_big > a.rows.n_scan) {
        // the others are sorted: shared memory for the largest row (wide if any read is); rows beyond kCtaMaxSmemWords
        // sort in a global slab
        uint64_t words = cta_words(a.rows.max_k_sort, a.rows.n_wide != 0);
        if (words > kCtaMaxSmemWords) words = kCtaMaxSmemWords;
        uint32_t per_sm = (uint32_t)((220u * 1024u) / (words * 4u + 1024u));
        if (per_sm < 1u) per_sm = 1u;
        if (per_sm > 8u) per_sm = 8u;
        uint32_t grid = (uint32_t)dc->n_sm * per_sm;
        if (grid > a.rows.n_big - a.rows.n_scan) grid = (uint32_t)(a.rows.n_big - a.rows.n_scan);
        big_kernel<<<grid, kCtaThreads, words * sizeof(uint32_t), forked ? a.side_stream : stream>>>(a, w, coverage, (uint32_t)words);
        ++launches;
    }
    if (forked && cudaEventRecord(a.ev_join, a.side_stream) != cudaSuccess) return -1;
    if (forked && cudaStreamWaitEvent(stream, a.ev_join, 0) != cudaSuccess) return -1;
    if (a.validate) {  // rows with a malformed interval: the reference's heap sweep itself, over what the kernels above listed
        literal_kernel<<<(uint32_t)dc->n_sm, 64, 0, stream>>>(a, w, coverage);
        ++launches;
    }
    order_kernel<<<w.n_parts, kOrderThreads, 0, stream>>>(a, w, not_coverage);
    ++launches;
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

int launch_peer_wait(const DetectArgs &a, cudaStream_t stream) {
    if (!a.n_peers) return 0;
    peer_wait_kernel<<<1, 32, 0, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_classify(const uint32_t *len, const uint32_t *gap_ptr, const uint2 *gaps, uint32_t n_reads, double not_coverage,
                    uint8_t *cls, uint8_t *bitmap, uint32_t *counters, cudaStream_t stream) {
    if (cudaMemsetAsync(counters, 0, kCounterWords * sizeof(uint32_t), stream) != cudaSuccess) return -1;
    if (n_reads == 0) return 0;
    const uint32_t threads = (n_reads + 15) / 16;
    classify_kernel<<<(threads + 255) / 256, 256, 0, stream>>>(len, gap_ptr, gaps, n_reads, not_coverage, cls, bitmap, counters);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace yb
