// detect.cu — sm_100a kernels of the detect hot path (v5).
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
// Kernels (all integer work; no tensor cores — there is no contraction on this path):
//   scatter_kernel   every row -> its size class (G = 1,2,3,4,5,6,8,10,16,32 lanes x 16 keys; packed or wide) and
//                    a 16-byte worklist record {row, first interval, k, len}; rows with k > 512 -> big list.
//   big_kernel       rows with k > 512: one CTA per row; 512-key chunks sorted by warps in registers, larger strides
//                    in shared memory (packed keys again; a global slab beyond 32768 intervals).
//   sort_kernel      persistent warps walk the worklist in batches of floor(32 / G) rows of ONE class, so a
//                    batch fills the warp with equal-sized lane groups. Each row's interval slab is pulled into
//                    shared memory by its own TMA bulk copy (cp.async.bulk, SASS UBLKCP; double-buffered: the
//                    copies of batch i+2 are issued when batch i is done). A group sorts its row in registers as
//                    PACKED u16x2 keys (begin | end << 16) — one VIMNMX.U16x2 moves a begin and an end through
//                    the same network, so both sorts cost one — or as two u32 arrays when the read is longer
//                    than 65534 bases. Crossings come from two carry-chain compares per slot against a
//                    transposed shared-memory copy; the batch's bad regions go to a bump-allocated staging
//                    segment (one atomic per batch, no waiting between warps).
//   order_kernel     single pass over the rows: scan of the per-row counts (decoupled look-back over cheap,
//                    uniform parts), staging -> ordered bad-region CSR, classification, 2-bit bitmap, histogram.
#include "pileup.cuh"
#include "sortnets.cuh"
#include <algorithm>
#include <cstdlib>

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

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
constexpr int E = 16;                          // keys per lane per array in the register tier
constexpr uint32_t kSmallMaxK = kRegisterTierMaxK;
#ifndef YB_SORT_WARPS
#define YB_SORT_WARPS 1
#endif
constexpr uint32_t kSortWarps = YB_SORT_WARPS;  // warps per CTA of sort_kernel (warps never synchronise with each other)
constexpr uint32_t kSortThreads = kSortWarps * 32;
#ifndef YB_SORT_MIN_CTAS
#define YB_SORT_MIN_CTAS (20 / YB_SORT_WARPS)
#endif
#ifndef YB_BUF_INTERVALS
#define YB_BUF_INTERVALS 528
#endif
constexpr uint32_t kBufIntervals = YB_BUF_INTERVALS;  // row slots of a slab buffer. A batch holds min(floor(32 / G), floor(kBufIntervals /
                                                      // (16 G + 2))) rows: 528 fits every class but G = 1 (29 rows instead of 32) and
                                                      // G = 2 (15 instead of 16), and lets 20 instead of 18 warps share an SM
constexpr uint32_t kScatterRows = 1024;        // rows per CTA of scatter_kernel
constexpr uint32_t kPartShift = 8, kPartRows = 1u << kPartShift;  // rows per CTA of order_kernel
constexpr uint32_t kStageChunk = 1024;       // pairs a warp reserves in the staging buffer per atomic
constexpr uint32_t kRecValid = 0x80000000u;    // worklist record .z = k | class << 16 | kRecValid

// Host-built table of the size classes (sizes are known from the row pointers at freeze time).
struct ClassTab {
    uint32_t entry_base[kNumClasses];     // where the class's records start in the worklist
    uint32_t count[kNumClasses];          // rows in the class
    uint32_t order[kNumClasses];          // classes in processing order (largest groups first)
    uint32_t item_base[kNumClasses + 1];  // batches before the q-th class in processing order
    uint32_t lanes[kNumClasses];          // G: lanes per row
    uint32_t rpb[kNumClasses];            // rows per batch = 32 / G
    uint32_t inv[kNumClasses];            // ceil(65536 / G): x / G == (x * inv) >> 16 for x < 2048
};

// Row-per-lane tier: where each slot class's records sit in the worklist and how its batches (32 rows) are numbered.
// Classes are processed largest first: the small kernel walks N = 64, 56, ..., 8, the mid kernel N = 128, ..., 72.
struct RLTab {
    uint32_t entry_base[kNumRL];
    uint32_t count[kNumRL];
    uint32_t item_base_small[kNumRL / 2 + 1];  // batches before the q-th class of the small kernel
    uint32_t item_base_mid[kNumRL / 2 + 1];
};

// scratch carve-up
struct Work {
    uint4 *recs;                     // n_reads worklist records {row, first interval, k | class << 16 | valid, len}
    uint32_t *soff;                  // n_reads: where the row's bad regions sit in `stage` (pairs)
    uint2 *stage;                    // bad regions in batch-completion order (warps reserve chunks with one atomic)
    uint32_t stage_cap;              // pairs
    uint32_t *part_total;            // bad regions of every part of kPartRows rows (RED by the sorting kernels)
    uint32_t *part_prefix;           // exclusive scan of part_total
    uint32_t n_parts;
    uint32_t *big_list;              // rows with k > kSmallMaxK
    uint32_t *huge_keys;             // event keys of rows beyond the shared-memory tier
};

// ------------------------------------------------------------------------------------------------
// scatter_kernel: rows -> worklist records grouped by size class (CTA-aggregated cursors)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScatterRows) scatter_kernel(DetectArgs a, Work w, ClassTab tab, RLTab rl, uint32_t c, uint32_t rl_max) {
    __shared__ uint32_t s_cnt[kNumAllClasses], s_base[kNumAllClasses];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, r = blockIdx.x * kScatterRows + tid;
    if (tid < (uint32_t)kNumAllClasses) s_cnt[tid] = 0u;
    if (tid < kScatterRows / kPartRows && blockIdx.x * (kScatterRows / kPartRows) + tid < w.n_parts)
        w.part_total[blockIdx.x * (kScatterRows / kPartRows) + tid] = 0u;
    __syncthreads();
    int cls = -2;  // 0 .. kNumClasses-1: lane-group classes; kNumClasses + q: row-per-lane slot class q; -1: big row
    uint32_t p0 = 0, k = 0, len = 0;
    if (r < a.n_reads) {
        p0 = __ldg(a.rowptr + r);
        k = __ldg(a.rowptr + r + 1) - p0;
        len = __ldg(a.len + r);
        const int q = rl_class_of_row(k, len, c, rl_max);
        cls = q >= 0 ? kNumClasses + q : class_of_row(k, len);
        if (cls < 0) {
            const uint32_t j = atomicAdd(a.counters + kCntBigList, 1u);
            w.big_list[j] = r;
        }
    }
    const uint32_t peers = __match_any_sync(FULL, cls);
    const uint32_t leader = __ffs(peers) - 1u, rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t wbase = 0;
    if (cls >= 0 && lane == leader) wbase = atomicAdd(&s_cnt[cls], (uint32_t)__popc(peers));
    wbase = __shfl_sync(FULL, wbase, leader);
    __syncthreads();
    if (tid < (uint32_t)kNumAllClasses && s_cnt[tid]) s_base[tid] = atomicAdd(a.counters + kCntClassCursor + tid, s_cnt[tid]);
    __syncthreads();
    if (cls >= 0) {
        const uint32_t eb = cls < kNumClasses ? tab.entry_base[cls] : rl.entry_base[cls - kNumClasses];
        w.recs[eb + s_base[cls] + wbase + rank] = make_uint4(r, p0, k | ((uint32_t)(cls < kNumClasses ? cls : 0) << 16) | kRecValid, len);
    }
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
__device__ void cta_row(const DetectArgs &a, const Work &w, uint32_t *keys, uint32_t r, uint32_t c, uint32_t *sh /* 16 u32 */) {
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
    if (__any_sync(FULL, bad_iv) && lane == 0) atomicAdd(a.counters + kCntMalformed, 1u);
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
    if (tid == 0) sh[10] = atomicAdd(a.counters + kCntStage, k + 1u);  // room for the row's k + 1 possible bad regions
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
            atomicAdd(a.counters + kCntStageOverflow, 1u);
        }
        a.gap_ptr[r] = ng;  // count for now; order_kernel turns it into the offset
        w.soff[r] = at;
        if (ng) atomicAdd(w.part_total + (r >> kPartShift), ng);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kCtaThreads) big_kernel(DetectArgs a, Work w, uint32_t c, uint32_t smem_words) {
    extern __shared__ __align__(16) uint32_t cta_smem[];
    __shared__ uint32_t sh[16];
    const uint32_t n_big = (uint32_t)a.rows.n_big;  // the host knows it from the row statistics
    for (uint32_t j = blockIdx.x; j < n_big; j += gridDim.x) {
        const uint32_t r = w.big_list[j];
        const uint32_t k = a.rowptr[r + 1] - a.rowptr[r];
        const bool wide = a.len[r] > kPackedMaxLen;
        const uint64_t words = cta_words(k, wide);
        uint32_t *keys = cta_smem;
        if (words > smem_words) {  // keys live in a bump-allocated global slab
            if (threadIdx.x == 0) sh[11] = atomicAdd(a.counters + kCntHugeBump, (uint32_t)words);
            __syncthreads();
            keys = w.huge_keys + sh[11];
        }
        if (wide) cta_row<false>(a, w, keys, r, c, sh);
        else cta_row<true>(a, w, keys, r, c, sh);
    }
}

constexpr uint32_t kScrPitch = 33;
constexpr uint32_t kScrWords = 16u * kScrPitch + 4u;  // T[t][lane] at scr[1 + 33 t + lane]; scr[0] stands for lane -1

struct alignas(16) WarpSmem {  // one per warp: a warp runs on its own, no CTA-wide barrier anywhere
    unsigned long long mbar[2];
    uint32_t scr[kScrWords];
};
static_assert(sizeof(WarpSmem) % 16 == 0, "slabs must stay 16-byte aligned");
constexpr size_t kWarpSmemBytes = sizeof(WarpSmem) + 2 * sizeof(uint2) * kBufIntervals;

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

// Lane geometry of a batch of class G: rows_per_batch groups of G consecutive lanes.
struct LaneGeo {
    uint32_t G, rpb, j, g;  // lanes per row, rows per batch, this lane's row slot and index in the group
    bool in_group;
};
__device__ __forceinline__ LaneGeo lane_geo(const ClassTab &tab, uint32_t cls) {
    LaneGeo x;
    x.G = tab.lanes[cls];
    x.rpb = tab.rpb[cls];
    const uint32_t lane = lane_id();
    x.j = (lane * tab.inv[cls]) >> 16;  // lane / G
    x.g = lane - x.j * x.G;
    x.in_group = x.j < x.rpb;
    if (!x.in_group) x.g = 0;
    return x;
}

// One batch: every in-group lane holds its row's record (all G lanes of a group hold the same one). Sorts the
// row's begins and ends, finds the crossings (U0 D0 U1 D1 ... written over the row's slab slot), turns them
// into bad regions and appends the batch's regions to the staging buffer.
template <bool PK>
__device__ __forceinline__ void process_batch(const DetectArgs &a, const Work &w, WarpSmem &ws, uint2 *buf, const LaneGeo geo,
                                              const uint4 rec, uint32_t c, uint2 &chunk) {
    const uint32_t lane = lane_id();
    const uint32_t G = geo.G, g = geo.g;
    const bool valid = geo.in_group && (rec.z & kRecValid);
    const uint32_t k = valid ? (rec.z & 0xFFFFu) : 0u, len = rec.w;
    uint2 *slot = buf + geo.j * (16u * G + 2u) + (rec.y & 1u);  // the row's data starts here
    // striped load (conflict-free); the initial arrangement is irrelevant to the sort
    uint32_t K0[E];             // PK: begin | end << 16; else begins
    uint32_t K1[PK ? 1 : E];    // else ends
    // (validity 0 <= b < e <= len is tested once per upload by launch_validate, not at every detect step)
    const uint2 *lane_iv = slot + g;           // this lane's elements: g, g + G, g + 2G, ...
    const uint32_t left = k > g ? k - g : 0u;  // element t exists iff t * G < left
    auto load16 = [&](const uint32_t GG) {     // GG: the batch's G as a compile-time constant where it is a common one
#pragma unroll
        for (int t = 0; t < E; ++t) {
            uint2 v = make_uint2(INF, INF);
            if ((uint32_t)t * GG < left) {
                v = lane_iv[(uint32_t)t * GG];
            }
            if (PK) {
                K0[t] = __byte_perm(v.x, v.y, 0x5410);
            } else {
                K0[t] = v.x;
                K1[PK ? 0 : t] = v.y;
            }
        }
    };
    switch (G) {  // constant strides turn the address arithmetic into immediates
        case 1: load16(1); break;
        case 2: load16(2); break;
        case 3: load16(3); break;
        case 4: load16(4); break;
        case 5: load16(5); break;
        case 6: load16(6); break;
        case 8: load16(8); break;
        default: load16(G); break;
    }
    if (PK) {
        sort_group<PK>(K0, G, g, geo.in_group);
    } else {
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {  // rolled: one copy of the network sorts begins, then ends
            sort_group<PK>(K0, G, g, geo.in_group);
#pragma unroll
            for (int t = 0; t < E; ++t) {
                const uint32_t x = K0[t];
                K0[t] = K1[PK ? 0 : t];
                K1[PK ? 0 : t] = x;
            }
        }
    }
    // transposed copy of the sorted ends (PK: of the packed keys; the compares only look at the end half):
    // T[t][lane]; element 16 l + t - c - 1 is then T[(t - c - 1) & 15][l + ((t - c - 1) >> 4)], a warp-uniform
    // offset from the lane's own column: conflict-free writes and reads, no per-element index arithmetic
    uint32_t *T = ws.scr + 1u + lane;
    __syncwarp();
#pragma unroll
    for (int t = 0; t < E; ++t) T[kScrPitch * t] = PK ? K0[t] : K1[PK ? 0 : t];
    __syncwarp();
    // V1_t = (E[16g + t - c - 1] <= B_t), t = 0..16;  V0_t = (E[16g + t - c] <= B_t), t = 0..15.
    // PK: (end_j <= begin_i)  <=>  key_j <= (begin_i << 16 | 0xFFFF) as plain u32.
    // Built most-significant-first: m1 bit (16 - t) = V1_t, m0 bit (15 - t) = V0_t.
    const uint32_t cc = min(c, 16u * 32u + 16u);  // beyond k every threshold behaves the same
    uint32_t m1 = 0, m0 = 0;
    {
        uint32_t Knext = __shfl_down_sync(FULL, K0[0], 1);
        if (g == G - 1u) Knext = INF;
#pragma unroll
        for (int t = 0; t <= E; ++t) {
            const int jr = t - (int)cc - 1;  // uniform
            int col = jr >> 4;
            if (cc >= 16u) col = max(col, -(int)lane - 1);  // stay inside scr; those elements are forced below
            const uint32_t ev = T[(int)kScrPitch * (jr & 15) + col];
            const uint32_t kt = t < E ? K0[t % E] : Knext;
            const uint32_t q = PK ? __byte_perm(kt, FULL, 0x1044) : kt;
            m1 = push_le(m1, ev, q);
            if (t > 0) {
                const uint32_t kp = K0[(t - 1) % E];
                const uint32_t qp = PK ? __byte_perm(kp, FULL, 0x1044) : kp;
                m0 = push_le(m0, ev, qp);
            }
        }
        // elements below the row's first end are 0 (E[-1] = 0): V1_t true for 16g + t <= c, V0_t for 16g + t < c
        const int z = (int)cc - 16 * (int)g;
        if (z >= 0) {
            const uint32_t zz = min((uint32_t)z, 16u);
            m1 |= ((2u << zz) - 1u) << (16u - zz);
            m0 |= ((1u << zz) - 1u) << (16u - zz);
        }
    }
    // bit (15 - t): U at begin t = V1_t & !V0_t; D at end t = !V0_t & V1_{t+1}
    uint32_t um = (m1 >> 1) & ~m0 & 0xFFFFu, dm = m1 & ~m0 & 0xFFFFu;
    if (!valid) um = dm = 0;
    // ranks of this lane's crossings among the row's ups / downs (packed segmented scan over the group)
    const uint32_t mine = __popc(um) | (__popc(dm) << 16);
    uint32_t incl = mine;
#pragma unroll 1
    for (uint32_t off = 1; off < G; off <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, incl, off);
        if (g >= off) incl += o;
    }
    const uint32_t tot = __shfl_sync(FULL, incl, min(lane - g + G - 1u, 31u));
    uint32_t ru = (incl - mine) & 0xFFFFu, rd = (incl - mine) >> 16;
    // crossings go back into the row's own slab slot (2k words, no longer needed): C[2j] = U_j, C[2j+1] = D_j
    uint32_t *C = reinterpret_cast<uint32_t *>(slot);
    if (PK) {
        while (um) {  // sparse: a row has a handful of crossings
            const int t = __clz(um) - 16;
            um &= ~(0x8000u >> t);
            C[2u * ru++] = T[(int)kScrPitch * t] & 0xFFFFu;
        }
        while (dm) {
            const int t = __clz(dm) - 16;
            dm &= ~(0x8000u >> t);
            const int jr = t - (int)cc;
            C[2u * rd++ + 1u] = T[(int)kScrPitch * (jr & 15) + (jr >> 4)] >> 16;
        }
    } else if (um | dm) {
#pragma unroll
        for (int t = 0; t < E; ++t) {
            if (um & (0x8000u >> t)) {
                C[2u * ru] = K0[t];
                ++ru;
            }
            if (dm & (0x8000u >> t)) {
                const int jr = t - (int)cc;
                C[2u * rd + 1u] = T[(int)kScrPitch * (jr & 15) + (jr >> 4)];
                ++rd;
            }
        }
    }
    __syncwarp();
    // ---- bad regions of the row (every lane of the group derives the same numbers) ----
    const uint32_t n_up = tot & 0xFFFFu;
    uint32_t ng = 0, h = 0, tail = 0;
    if (valid) {
        if (n_up) {
            h = C[0] != 0u;
            tail = C[2u * n_up - 1u] != len;
            ng = n_up - 1u + h + tail;
        } else {
            ng = h = tail = len != 0u;
        }
    }
    // staging: the warp owns a chunk of the staging buffer and refills it with one atomic when it runs out
    const uint32_t inc = warp_incl_scan(g == 0u ? ng : 0u);
    const uint32_t total = __shfl_sync(FULL, inc, 31);
    uint32_t base;
    if (total <= chunk.y - chunk.x) {
        base = chunk.x;
        chunk.x += total;
    } else {
        const bool direct = total >= kStageChunk / 4u;  // a large batch takes exactly what it needs
        uint32_t got = 0;
        if (lane == 0) got = atomicAdd(a.counters + kCntStage, direct ? total : kStageChunk);
        base = __shfl_sync(FULL, got, 0);
        if (!direct) chunk = make_uint2(base + total, base + kStageChunk);
    }
    base += inc - ng;
    if (valid && base + ng > w.stage_cap) {  // cannot happen with the capacity the engine allocates; never write outside
        if (g == 0u) atomicAdd(a.counters + kCntStageOverflow, 1u);
        ng = 0;
    }
    if (valid) {
        if (g == 0u) {
            a.gap_ptr[rec.x] = ng;  // count for now; order_kernel turns it into the offset
            w.soff[rec.x] = base;
            if (ng) atomicAdd(w.part_total + (rec.x >> kPartShift), ng);
        }
        for (uint32_t gq = g; gq < ng; gq += G) {
            const uint32_t f0 = 2u * gq, f1 = f0 + 1u;
            uint2 o;
            o.x = (f0 == 0u && h) ? 0u : C[f0 + 1u - 2u * h];
            o.y = (f1 == 2u * ng - 1u && tail) ? len : C[f1 + 1u - 2u * h];
            w.stage[base + gq] = o;
        }
    }
}

// The item's record for this lane (the record of the row its group sorts) — a plain 16-byte load that
// nothing touches until the batch is issued, so it stays in flight behind the current batch.
__device__ __forceinline__ uint4 load_rec(const Work &w, const ClassTab &tab, uint32_t item, uint32_t n_items, uint32_t &q, uint32_t &cls) {
    uint4 rec = make_uint4(0, 0, 0, 0);
    cls = 0;
    if (item >= n_items) return rec;
    while (item >= tab.item_base[q + 1]) ++q;
    cls = tab.order[q];
    const LaneGeo geo = lane_geo(tab, cls);
    const uint32_t e = (item - tab.item_base[q]) * geo.rpb + geo.j;
    if (geo.in_group && e < tab.count[cls]) rec = __ldg(w.recs + tab.entry_base[cls] + e);
    return rec;
}

// TMA copies of the batch's row slabs into `buf` (one per row, issued by the group's first lane).
__device__ __forceinline__ void issue_batch(const DetectArgs &a, const ClassTab &tab, uint2 *buf, unsigned long long *bar,
                                            const uint4 rec, uint32_t cls) {
    const LaneGeo geo = lane_geo(tab, cls);
    uint32_t bytes = 0, cs = 0;
    if (geo.in_group && geo.g == 0u && (rec.z & kRecValid)) {
        cs = rec.y & ~1u;
        bytes = (((rec.y + (rec.z & 0xFFFFu) + 1u) & ~1u) - cs) * 8u;
    }
    const uint32_t total = warp_sum(bytes);
    if (lane_id() == 0) mbar_expect_tx(bar, total);
    __syncwarp();
    if (bytes) tma_load_1d(buf + geo.j * (16u * geo.G + 2u), a.iv + cs, bytes, bar);
}

__global__ void __launch_bounds__(kSortThreads, YB_SORT_MIN_CTAS) sort_kernel(DetectArgs a, Work w, ClassTab tab, uint32_t c) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    WarpSmem &ws = *reinterpret_cast<WarpSmem *>(smem_raw + wid * kWarpSmemBytes);
    uint2 *buf0 = reinterpret_cast<uint2 *>(smem_raw + wid * kWarpSmemBytes + sizeof(WarpSmem));
    if (lane == 0) {
        mbar_init(&ws.mbar[0], 1);
        mbar_init(&ws.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint32_t n_items = tab.item_base[kNumClasses];
    uint32_t q = 0;
    // Dynamic schedule (batches cost between 0.3 and 2 us): a warp draws batch indices from one counter, three
    // batches ahead, so the atomic's latency hides behind a whole batch. Indices drawn by a warp only grow.
    auto draw_raw = [&]() {  // lane 0 holds the index; nobody waits for the atomic until the value is broadcast
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(a.counters + kCntTile, 1u);
        return t;
    };
    uint32_t item = __shfl_sync(FULL, draw_raw(), 0), item1 = __shfl_sync(FULL, draw_raw(), 0), item2 = __shfl_sync(FULL, draw_raw(), 0);
    // software pipeline: records of batch i+2 are loaded, the slabs of batch i+1 are in flight, batch i is sorted
    uint32_t cls0, cls1, cls2;
    uint4 rec0 = load_rec(w, tab, item, n_items, q, cls0);
    uint4 rec1 = load_rec(w, tab, item1, n_items, q, cls1);
    if (item < n_items) issue_batch(a, tab, buf0, &ws.mbar[0], rec0, cls0);
    if (item1 < n_items) issue_batch(a, tab, buf0 + kBufIntervals, &ws.mbar[1], rec1, cls1);
    uint32_t b = 0, parity = 0;
    uint2 chunk = make_uint2(0, 0);  // [next free pair, end) of the warp's staging chunk
    while (item < n_items) {
        const uint32_t raw3 = draw_raw();  // consumed at the end of this iteration
        const uint4 rec2 = load_rec(w, tab, item2, n_items, q, cls2);
        uint2 *buf = buf0 + b * kBufIntervals;
        mbar_wait(&ws.mbar[b], parity);
        const LaneGeo geo = lane_geo(tab, cls0);
        if (cls0 < (uint32_t)kNumG) process_batch<true>(a, w, ws, buf, geo, rec0, c, chunk);
        else process_batch<false>(a, w, ws, buf, geo, rec0, c, chunk);
        // generic-proxy accesses of this batch (crossings written into the slab) before the async-proxy refill
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (item2 < n_items) issue_batch(a, tab, buf, &ws.mbar[b], rec2, cls2);
        rec0 = rec1;
        rec1 = rec2;
        cls0 = cls1;
        cls1 = cls2;
        item = item1;
        item1 = item2;
        item2 = __shfl_sync(FULL, raw3, 0);
        parity ^= b;
        b ^= 1u;
    }
}

constexpr size_t kSortSmemBytes = kWarpSmemBytes * kSortWarps;

// ------------------------------------------------------------------------------------------------
// row-per-lane tier (rl_kernel): packed rows of at most 128 key slots, ONE lane per row, 32 rows of one slot class per
// warp. Everything the lane-group tier pays for sharing a row between lanes disappears: no shuffles and no
// predicated exchanges in the sort (a straight-line merge-exchange network over the N registers of the lane,
// sortnets.cuh), no per-group bookkeeping replicated G times, and the crossing tests index registers only.
//
// Sentinel ends. The closed form compares B_i with E[i-c-1] and E[i-c]: a shift by the run-time threshold c in RANK
// space. The two halves of a packed key are sorted independently, so the shift is done by the sort itself: c' + 1
// extra keys (begin = +inf, end = 0), c' = min(c, k), are added to the row. After the sort the high halves are
// E'_i = E[i - c' - 1] (0 below the first end, exactly the convention E[-1] = 0), hence
//     a_i = V1_i  = (E'_i     <= B_i)        b_i = !V0_i = (B_i < E'_{i+1})
//     up-crossing at begin i = a_i & b_i,    down-crossing (value E'_{i+1}) = b_i & a_{i+1},
// all with compile-time register indices. Pushing a_0, !b_0, a_1, !b_1, ... into one bit string Z (most significant
// first) makes the crossings, in the order U0 D0 U1 D1 ... of the bad-region list, the set bits of Z & (Z << 1).
// (c >= k behaves like c = k: the depth never exceeds k.)
//
// Staging. A warp-instruction of this tier serves 32 rows, so the kernel lives on occupancy (measured: about 0.12
// instructions per clock per resident warp, whatever the tier), and occupancy is shared memory: the batch is
// therefore staged PACKED and TRANSPOSED, Tp[slot][row] with a pitch of 33 words (4 bytes per interval instead of
// 8; 8.4 KB per warp for 64 slots). The warp walks the 32 rows, one coalesced 8-byte load per lane and row, packs
// (PRMT) and stores conflict-free; afterwards lane j reads its row down column j, conflict-free again, with
// immediate offsets. The sorted keys go back into the same column (the crossing VALUES are fetched from there by
// run-time index). cp.async / TMA would keep 8 bytes per interval in shared memory and cost more issue slots per row
// (LDGSTS: three dummy LDS per copy on sm_100a; UBLKCP: an ELECT / R2UR loop).
// Interval validity (0 <= b < e <= len) is not re-tested here: launch_validate does it once per upload.
// ------------------------------------------------------------------------------------------------
#ifndef YB_RL_WARPS_SMALL
#define YB_RL_WARPS_SMALL 4
#endif
#ifndef YB_RL_WARPS_MID
#define YB_RL_WARPS_MID 4
#endif
#ifndef YB_RL_MIN_CTAS_SMALL
#define YB_RL_MIN_CTAS_SMALL 5
#endif
#ifndef YB_RL_MIN_CTAS_MID
#define YB_RL_MIN_CTAS_MID 3
#endif
constexpr uint32_t kTpPitch = 33;  // words between consecutive slots of Tp
// per warp: Tp (nmax slots + the +inf key behind the last one) and the 32 copy descriptors
// Tp: one leading dump slot (local slot -1 of a row that starts at an odd interval lands there), nmax slots, the +inf key
__host__ __device__ constexpr size_t rl_tp_bytes(uint32_t nmax) { return ((nmax + 2u) * kTpPitch * sizeof(uint32_t) + 15u) & ~(size_t)15; }
__host__ __device__ constexpr size_t rl_warp_smem(uint32_t nmax) { return rl_tp_bytes(nmax) + 32u * sizeof(uint4); }

// Transposing copy of the batch's rows into Tp: row j's interval t -> Tp[1 + t][j] as begin | end << 16.
// desc[j] = {aligned first interval / 2, k + o, 132 o, first interval} of row j (o = first interval & 1), read back as
// broadcasts. Rows of at most 64 slots: ONE aligned 16-byte load per lane and row (two intervals), two stores; lanes
// beyond the row's end store stale registers (no second predicate) that the sentinel loop overwrites.
__device__ __forceinline__ void rl_stage16(const DetectArgs &a, uint32_t *Tp, const uint4 *desc) {
    constexpr int RB = 8;  // rows whose loads are in flight together (all loads of a group are issued before its stores:
                           // a shared-memory store would otherwise fence the next descriptor read)
    const uint32_t lane = lane_id();
    const uint4 *src_lane = reinterpret_cast<const uint4 *>(a.iv) + lane;
    char *dst_lane = reinterpret_cast<char *>(Tp + (2u * lane + 1u) * kTpPitch);
    uint4 v[RB] = {};  // zeroed once per batch: a lane beyond a row's end stores whatever an earlier row left there
#pragma unroll 1
    for (int j0 = 0; j0 < 32; j0 += RB) {
        uint32_t off[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const uint4 d = desc[j0 + r];
            off[r] = d.z;
            if (2u * lane < d.y) v[r] = __ldg(src_lane + d.x);
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            uint32_t *dst = reinterpret_cast<uint32_t *>(dst_lane - off[r]) + j0 + r;
            dst[0] = __byte_perm(v[r].x, v[r].y, 0x5410);
            dst[kTpPitch] = __byte_perm(v[r].z, v[r].w, 0x5410);
        }
    }
}

// The same for rows of up to 128 slots: CH chunks of 32 intervals, one 8-byte load per lane, row and chunk.
template <int CH>
__device__ __forceinline__ void rl_stage(const DetectArgs &a, uint32_t *Tp, const uint4 *desc) {
    constexpr int RB = 4;
    const uint32_t lane = lane_id();
    const uint2 *src_lane = a.iv + lane;
    uint32_t *dst_lane = Tp + (lane + 1u) * kTpPitch;
#pragma unroll 1
    for (int j0 = 0; j0 < 32; j0 += RB) {
        uint2 v[RB][CH];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const uint4 d = desc[j0 + r];
            const uint32_t k = d.y - (d.w & 1u);
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                v[r][ch] = make_uint2(0u, 0u);
                if (lane + 32u * ch < k) v[r][ch] = __ldg(src_lane + d.w + 32 * ch);
            }
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) dst_lane[32 * ch * kTpPitch + j0 + r] = __byte_perm(v[r][ch].x, v[r][ch].y, 0x5410);
        }
    }
}

template <int N>
__device__ __forceinline__ void rl_batch(const DetectArgs &a, const Work &w, uint32_t *Tp, uint4 *desc, const uint4 rec, uint32_t c,
                                         uint2 &chunk) {
    constexpr int W = N / 16 + ((N % 16) ? 1 : 0);  // 32-bit words of Z: 16 slots each
    const uint32_t lane = lane_id();
    const bool valid = (rec.z & kRecValid) != 0u;
    const uint32_t k = valid ? (rec.z & 0xFFFFu) : 0u, len = rec.w;
    const uint32_t cp = min(c, k);
    desc[lane] = make_uint4(rec.y >> 1, k + (rec.y & 1u), (rec.y & 1u) * kTpPitch * 4u, rec.y);
    __syncwarp();
    if (N <= 64) rl_stage16(a, Tp, desc);
    else rl_stage<(N + 31) / 32>(a, Tp, desc);
    __syncwarp();
    uint32_t *col = Tp + kTpPitch + lane;  // the lane's row: slot t at col[t * kTpPitch]
    // sentinels behind the row's intervals: c' + 1 keys (+inf, 0), then (+inf, +inf)
    for (uint32_t t = k; t < (uint32_t)N; ++t) col[t * kTpPitch] = t <= k + cp ? 0x0000FFFFu : FULL;
    __syncwarp();
    uint32_t K[N + 1];
#pragma unroll
    for (int t = 0; t < N; ++t) K[t] = col[t * kTpPitch];
    K[N] = FULL;
    {
        uint32_t(&Ks)[N] = *reinterpret_cast<uint32_t(*)[N]>(&K[0]);
        SortNet<N>::run(Ks, [](uint32_t &x, uint32_t &y) {
            const uint32_t lo = __vminu2(x, y), hi = __vmaxu2(x, y);
            x = lo;
            y = hi;
        });
    }
    // the sorted keys go back to the lane's column: crossing VALUES are fetched from there by run-time index
#pragma unroll
    for (int t = 0; t <= N; ++t) col[t * kTpPitch] = K[t];
    // Z: a_i at bit 31 - 2t, !b_i at bit 30 - 2t of word i / 16 (t = i % 16); X = crossings, in bad-region order
    uint32_t X[W];
    uint32_t m = 0;
    {
        uint32_t Z[W + 1];
#pragma unroll
        for (int wd = 0; wd < W; ++wd) {
            uint32_t z = 0;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int i = 16 * wd + t;
                if (i < N) {
                    const uint32_t q = __byte_perm(K[i], FULL, 0x1044);  // B_i << 16 | 0xFFFF: key <= q  <=>  end half <= B_i
                    z = push_le(z, K[i], q);
                    z = push_le(z, K[i + 1], q);
                } else {
                    z <<= 2;
                }
            }
            Z[wd] = z ^ 0x55555555u;
        }
        Z[W] = 0u;
#pragma unroll
        for (int wd = 0; wd < W; ++wd) {
            X[wd] = valid ? (Z[wd] & __funnelshift_l(Z[wd + 1], Z[wd], 1)) : 0u;
            m += __popc(X[wd]);
        }
    }
    if (m & 1u) m = 0;  // only malformed rows (already reported at upload): treated as "never above c"
    // staging: a row reserves an upper bound (m / 2 + 1 regions: whether the list starts at 0 / ends at len is only known
    // once the first / last crossing VALUE is read); order_kernel copies the ng regions actually written.
    // The warp owns a chunk of the staging buffer and refills it with one atomic when it runs out.
    const uint32_t ub = valid ? (m ? (m >> 1) + 1u : (len != 0u ? 1u : 0u)) : 0u;
    const uint32_t inc = warp_incl_scan(ub);
    const uint32_t total = __shfl_sync(FULL, inc, 31);
    uint32_t base;
    if (total <= chunk.y - chunk.x) {
        base = chunk.x;
        chunk.x += total;
    } else {
        const bool direct = total >= kStageChunk / 4u;  // a large batch takes exactly what it needs
        uint32_t got = 0;
        if (lane == 0) got = atomicAdd(a.counters + kCntStage, direct ? total : kStageChunk);
        base = __shfl_sync(FULL, got, 0);
        if (!direct) chunk = make_uint2(base + total, base + kStageChunk);
    }
    base += inc - ub;
    if (valid && (uint64_t)base + ub > w.stage_cap) {  // cannot happen with the capacity the engine allocates; never write outside
        atomicAdd(a.counters + kCntStageOverflow, 1u);
        m = 0;
        base = 0xFFFFFFFFu;
    }
    // flat list [0 if h] U0 D0 U1 D1 ... [len if tail], minus U0 when U0 == 0 (h false), minus D_last when D_last == len:
    // crossing j goes to flat position j + off, off = +1 (h) or -1 (!h); position -1 (U0 == 0) is simply not written
    uint32_t *F = reinterpret_cast<uint32_t *>(w.stage + (base == 0xFFFFFFFFu ? 0u : base));
    uint32_t off = 1u, j = 0, v = 0;
    if (m) F[0] = 0u;
#pragma unroll
    for (int wd = 0; wd < W; ++wd) {
        uint32_t x = m ? X[wd] : 0u;
        while (x) {
            const uint32_t p = __clz(x);
            x &= ~(0x80000000u >> p);
            const uint32_t i = 16u * wd + ((p + 1u) >> 1);  // up-crossing: B_i; down-crossing: E'_{i+1}
            v = (col[i * kTpPitch] >> (16u * (p & 1u))) & 0xFFFFu;
            if (j == 0u && v == 0u) off = 0xFFFFFFFFu;
            const uint32_t pos = j + off;
            if (pos != 0xFFFFFFFFu) F[pos] = v;
            ++j;
        }
    }
    if (valid) {
        uint32_t ng = 0;
        if (m) {  // v = D_last: the list ends with (D_last, len) unless D_last == len
            const uint32_t flat = m + off + (v != len ? 1u : 0xFFFFFFFFu);
            if (v != len) F[m + off] = len;
            ng = flat >> 1;
        } else if (base != 0xFFFFFFFFu && len != 0u) {  // depth never above c: one region (0, len)
            F[0] = 0u;
            F[1] = len;
            ng = 1u;
        }
        a.gap_ptr[rec.x] = ng;  // count for now; order_kernel turns it into the offset
        w.soff[rec.x] = base == 0xFFFFFFFFu ? 0u : base;
        if (ng) atomicAdd(w.part_total + (rec.x >> kPartShift), ng);
    }
}

template <bool MID>
__global__ void __launch_bounds__(32 * (MID ? YB_RL_WARPS_MID : YB_RL_WARPS_SMALL), MID ? YB_RL_MIN_CTAS_MID : YB_RL_MIN_CTAS_SMALL)
    rl_kernel(DetectArgs a, Work w, RLTab tab, uint32_t c) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr uint32_t NMAX = MID ? kRLMaxSlots : kRLSmallSlots;
    constexpr int NQ = kNumRL / 2;
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    uint32_t *Tp = reinterpret_cast<uint32_t *>(smem_raw + wid * rl_warp_smem(NMAX));
    uint4 *desc = reinterpret_cast<uint4 *>(smem_raw + wid * rl_warp_smem(NMAX) + rl_tp_bytes(NMAX));
    const uint32_t *item_base = MID ? tab.item_base_mid : tab.item_base_small;
    const uint32_t n_items = item_base[NQ];
    // Static schedule: warp g of the grid takes batches g, g + G, g + 2 G, ... Batches of one class cost the same and the
    // classes are walked largest first, so the warps stay within one batch of each other; warps never synchronise.
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    uint32_t q = 0;
    // the q-th class in processing order has N = NMAX - 8 q slots: slot class index N / 8 - 1
    auto load_rec = [&](uint32_t item, uint32_t &nslots) {
        uint4 rec = make_uint4(0, 0, 0, 0);
        nslots = 8u;
        if (item >= n_items) return rec;
        while (item >= item_base[q + 1]) ++q;
        nslots = NMAX - 8u * q;
        const uint32_t cl = nslots / 8u - 1u;
        const uint32_t e = (item - item_base[q]) * 32u + lane;
        if (e < tab.count[cl]) rec = __ldg(w.recs + tab.entry_base[cl] + e);
        return rec;
    };
    uint32_t item = blockIdx.x * (blockDim.x >> 5) + wid;
#ifdef YB_RL_SKEW_NS
    {   // the warps of a scheduler would otherwise walk through the same phases at the same time (same class, same cost,
        // same start): all loading, then all sorting. One start-up delay per resident CTA slot takes them out of step.
        const uint32_t per_slot = max(1u, gridDim.x / (MID ? YB_RL_MIN_CTAS_MID : YB_RL_MIN_CTAS_SMALL));
        const uint32_t slot = blockIdx.x / per_slot;
        if (slot) __nanosleep(slot * YB_RL_SKEW_NS);
    }
#endif
    uint32_t n0, n1, n2;
    uint4 rec0 = load_rec(item, n0);
    uint4 rec1 = load_rec(item + n_warps, n1);
    uint2 chunk = make_uint2(0, 0);  // [next free pair, end) of the warp's staging chunk
    while (item < n_items) {
        const uint4 rec2 = load_rec(item + 2u * n_warps, n2);  // in flight behind this batch
        // the next batch's rows are pulled into L2 while this one is sorted (lane j touches the 128-byte lines of its
        // own next row): the staging loads of the next batch then pay an L2 hit instead of a DRAM access
        if (rec1.z & kRecValid) {
            const char *p = reinterpret_cast<const char *>(a.iv + rec1.y);
            const char *e = p + 8u * (rec1.z & 0xFFFFu);
            for (p = reinterpret_cast<const char *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)127); p < e; p += 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        }
#define YB_RL_CASE(NN) case NN: rl_batch<NN>(a, w, Tp, desc, rec0, c, chunk); break;
        if (MID) {
            switch (n0) {
                YB_RL_CASE(72) YB_RL_CASE(80) YB_RL_CASE(88) YB_RL_CASE(96) YB_RL_CASE(104) YB_RL_CASE(112) YB_RL_CASE(120)
                default: rl_batch<128>(a, w, Tp, desc, rec0, c, chunk); break;
            }
        } else {
            switch (n0) {
                YB_RL_CASE(8) YB_RL_CASE(16) YB_RL_CASE(24) YB_RL_CASE(32) YB_RL_CASE(40) YB_RL_CASE(48) YB_RL_CASE(56)
                default: rl_batch<64>(a, w, Tp, desc, rec0, c, chunk); break;
            }
        }
#undef YB_RL_CASE
        __syncwarp();  // every lane is done with its column before the next batch is staged
        rec0 = rec1;
        rec1 = rec2;
        n0 = n1;
        n1 = n2;
        item += n_warps;
    }
}

// ------------------------------------------------------------------------------------------------
// ordering pass: the sorting kernels left, per row, a count and a staging offset, and per part of 256 rows the
// part's total (RED). scan_parts_kernel scans the part totals; order_kernel turns counts into offsets inside
// each part, moves the staged regions to their final place and classifies (editor/mod.rs:85-100).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) scan_parts_kernel(DetectArgs a, Work w) {
    __shared__ uint32_t sh[32];
    const uint32_t tid = threadIdx.x, n = w.n_parts;
    const uint32_t per = (n + 1023u) / 1024u, beg = min(tid * per, n), end = min(beg + per, n);
    uint32_t s = 0;
    for (uint32_t i = beg; i < end; ++i) s += w.part_total[i];
    const uint32_t incl = warp_incl_scan(s);
    if ((tid & 31u) == 31u) sh[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32u) {
        const uint32_t v = sh[tid], iv = warp_incl_scan(v);
        sh[tid] = iv - v;
        if (tid == 31u) a.gap_ptr[a.n_reads] = iv;
    }
    __syncthreads();
    uint32_t run = sh[tid >> 5] + incl - s;
    for (uint32_t i = beg; i < end; ++i) {
        w.part_prefix[i] = run;
        run += w.part_total[i];
    }
}

#ifndef YB_ORDER_MIN_CTAS
#define YB_ORDER_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(kPartRows, YB_ORDER_MIN_CTAS) order_kernel(DetectArgs a, Work w, double not_cov) {
    __shared__ uint32_t s_warp[kPartRows / 32], s_hist[kPartRows / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t part = blockIdx.x, r = part * kPartRows + tid;
    const bool live = r < a.n_reads;
    // everything the row needs is requested up front; the first two regions (most rows have <= 3) ride along
    const uint32_t cnt = live ? a.gap_ptr[r] : 0u;
    const uint32_t l = live ? __ldg(a.len + r) : 0u;
    const uint2 *src = w.stage + (live ? w.soff[r] : 0u);
    const uint32_t prefix = __ldg(w.part_prefix + part);
    uint2 v0 = make_uint2(0, 0), v1 = make_uint2(0, 0);
    if (cnt > 0u) v0 = src[0];
    if (cnt > 1u) v1 = src[1];
    const uint32_t incl = warp_incl_scan(cnt);
    if (lane == 31u) s_warp[wid] = incl;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (uint32_t q = 0; q < kPartRows / 32; ++q) before += q < wid ? s_warp[q] : 0u;
    const uint32_t gp = prefix + before + incl - cnt;
    uint32_t cl = 0;
    if (live) {
        uint32_t bad = 0, interior = 0;
        if (cnt > 0u) {
            a.gaps[gp] = v0;
            bad += v0.y - v0.x;
            interior |= (v0.x != 0u && v0.y != l) ? 1u : 0u;
        }
        if (cnt > 1u) {
            a.gaps[gp + 1u] = v1;
            bad += v1.y - v1.x;
            interior |= (v1.x != 0u && v1.y != l) ? 1u : 0u;
        }
        for (uint32_t gq = 2; gq < cnt; ++gq) {
            const uint2 v = src[gq];
            a.gaps[gp + gq] = v;
            bad += v.y - v.x;
            interior |= (v.x != 0u && v.y != l) ? 1u : 0u;
        }
        cl = classify(bad, l, interior ? 2u : 0u, not_cov);
        a.cls[r] = (uint8_t)cl;
        a.gap_ptr[r] = gp;
    }
    // 16 consecutive rows (half a warp) make one 32-bit word of the 2-bit bitmap
    uint32_t bits = cl << (2u * (lane & 15u));
    bits |= __shfl_xor_sync(FULL, bits, 1);
    bits |= __shfl_xor_sync(FULL, bits, 2);
    bits |= __shfl_xor_sync(FULL, bits, 4);
    bits |= __shfl_xor_sync(FULL, bits, 8);
    if ((lane & 15u) == 0u && live) {
        if (a.n_peers == 0u) {
            reinterpret_cast<uint32_t *>(a.bitmap)[r >> 4] = bits;
        } else {  // all-gather fused into the epilogue: the word goes to this rank's slot on every rank (NVLink stores)
            for (uint32_t p = 0; p < a.n_peers; ++p) reinterpret_cast<uint32_t *>(a.peer_slot[p])[r >> 4] = bits;
        }
    }
    // class histogram: one RED per class per CTA, spread over kHistSlots copies (same-address atomics serialise in L2)
    const uint32_t h1 = __popc(__ballot_sync(FULL, cl == 1u)), h2 = __popc(__ballot_sync(FULL, cl == 2u));
    if (lane == 0) s_hist[wid] = h1 | (h2 << 16);
    __syncthreads();
    if (tid == 0) {
        uint32_t hs = 0;
#pragma unroll
        for (uint32_t q = 0; q < kPartRows / 32; ++q) hs += s_hist[q];
        const uint32_t n_live = min(kPartRows, a.n_reads - part * kPartRows);
        uint32_t *slot = a.counters + kCntHist + 3u * (part % kHistSlots);
        const uint32_t c1 = hs & 0xFFFFu, c2 = hs >> 16;
        if (n_live - c1 - c2) atomicAdd(slot + 0, n_live - c1 - c2);
        if (c1) atomicAdd(slot + 1, c1);
        if (c2) atomicAdd(slot + 2, c2);
    }
}

// Closes a step of the peer-memory all-gather: thread p tells rank p "rank `rank` has written its slot for `epoch`"
// and then waits until rank p has said the same here. Bounded wait: a missing rank must not hang the GPU.
__global__ void __launch_bounds__(32) peer_barrier_kernel(DetectArgs a) {
    const uint32_t p = threadIdx.x;
    // the step number lives in device memory (word 31 of this rank's flag buffer), so a captured CUDA graph that is
    // replayed still counts; every rank runs the same number of steps
    uint32_t epoch = 0;
    if (p == 0) {
        epoch = a.peer_flag[a.rank][31] + 1u;
        a.peer_flag[a.rank][31] = epoch;
    }
    epoch = __shfl_sync(FULL, epoch, 0);
    if (p >= a.n_peers) return;
    __threadfence_system();  // the ordering kernel's peer stores before the flag
    uint32_t *theirs = a.peer_flag[p] + a.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    const uint32_t *mine = a.peer_flag[a.rank] + p;
    uint32_t seen = 0;
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        if ((int32_t)(seen - epoch) >= 0) break;
        __nanosleep(100);
    }
    if ((int32_t)(seen - epoch) < 0) atomicAdd(a.counters + kCntPeerTimeout, 1u);
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
    __shared__ uint32_t s_cnt[kNumClasses + 1], s_max, s_bad[3], s_hist[kRLMaxSlots];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, r = blockIdx.x * 1024u + tid;
    if (tid <= (uint32_t)kNumClasses) s_cnt[tid] = 0u;
    if (tid < kRLMaxSlots) s_hist[tid] = 0u;
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
                const unsigned long long hk = cta_words(k, l > kPackedMaxLen);
                if (hk > kCtaMaxSmemWords) atomicAdd(&out->huge_keys, hk);
            }
            if (l > kPackedMaxLen) atomicAdd(&s_bad[2], 1u);
            else if (k < kRLMaxSlots) atomicAdd(&s_hist[k], 1u);
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
    if (tid < kRLMaxSlots && s_hist[tid]) atomicAdd(&out->k_hist[tid], s_hist[tid]);
    if (tid == 0) {
        if (s_cnt[kNumClasses]) atomicAdd(&out->n_big, s_cnt[kNumClasses]);
        if (s_max) atomicMax(&out->max_k, s_max);
        if (s_bad[0]) atomicAdd(&out->bad_rowptr, s_bad[0]);
        if (s_bad[1]) atomicAdd(&out->bad_len, s_bad[1]);
        if (s_bad[2]) atomicAdd(&out->n_wide, s_bad[2]);
    }
}

// validate_kernel (upload time): 0 <= begin < end <= length for every interval. A warp takes 32 consecutive rows and
// walks each of them with coalesced loads.
__global__ void __launch_bounds__(256) validate_kernel(const uint2 *__restrict__ iv, const uint32_t *__restrict__ rowptr,
                                                       const uint32_t *__restrict__ len, uint32_t n_reads, DevRowStats *out) {
    const uint32_t lane = lane_id(), warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t bad = 0;
    for (uint32_t r0 = warp * 32u; r0 < n_reads; r0 += n_warps * 32u) {
        const uint32_t r = r0 + lane;
        uint32_t p0 = 0, p1 = 0, l = 0;
        if (r < n_reads) {
            p0 = __ldg(rowptr + r);
            p1 = __ldg(rowptr + r + 1);
            l = __ldg(len + r);
        }
        const uint32_t rows = min(32u, n_reads - r0);
        for (uint32_t j = 0; j < rows; ++j) {
            const uint32_t b = __shfl_sync(FULL, p0, j), e = __shfl_sync(FULL, p1, j), lj = __shfl_sync(FULL, l, j);
            for (uint32_t i = b + lane; i < e; i += 32u) {
                const uint2 v = __ldg(iv + i);
                bad += !(v.x < v.y && v.y <= lj);
            }
        }
    }
    bad = warp_sum(bad);
    if (lane == 0 && bad) atomicAdd(&out->malformed, bad);
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
    w.n_parts = (a.n_reads + kPartRows - 1u) / kPartRows;
    w.recs = reinterpret_cast<uint4 *>(take(sizeof(uint4) * ((size_t)a.n_reads + 1)));
    (void)big_pairs;
    w.soff = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * ((size_t)a.n_reads + 1)));
    // worst case of the bad regions (k + 1 per row) + half of it for chunk remainders (a remainder is only dropped
    // for a batch smaller than a quarter chunk, and a fresh chunk always holds at least two such batches)
    const uint64_t cap = (uint64_t)a.n_iv + a.n_reads + ((uint64_t)a.n_iv + a.n_reads) / 2 + 4096ull * kStageChunk;  // + one open chunk per resident warp
    w.stage_cap = cap > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)cap;
    w.stage = reinterpret_cast<uint2 *>(take(sizeof(uint2) * ((size_t)w.stage_cap + 1)));
    w.part_total = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * ((size_t)w.n_parts + 8)));
    w.part_prefix = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * ((size_t)w.n_parts + 8)));
    w.big_list = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * (n_big + 1)));
    w.huge_keys = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * (huge_keys + 1)));
    *total = off;
    return w;
}

// The step's class tables: where every class's records sit in the worklist and how its batches are numbered.
struct Plan {
    ClassTab tab;
    RLTab rl;
    uint32_t rl_max, items, items_small, items_mid;
};
Plan make_plan(const DetectArgs &a, uint32_t coverage, uint32_t rl_max) {
    Plan pl{};
    ClassTab &tab = pl.tab;
    RLTab &rl = pl.rl;
    pl.rl_max = rl_max;
    // row-per-lane tier: packed rows whose k + min(c, k) + 1 key slots fit 128 leave their lane-group class
    uint32_t lg_count[kNumClasses];
    for (int cl = 0; cl < kNumClasses; ++cl) lg_count[cl] = a.rows.class_count[cl];
    for (uint32_t k = 0; k < kRLMaxSlots; ++k) {
        const int q = rl_class_of_row(k, 0u, coverage, rl_max);
        if (q < 0 || !a.rows.k_hist[k]) continue;
        rl.count[q] += a.rows.k_hist[k];
        lg_count[class_of_row(k, 0u)] -= a.rows.k_hist[k];
    }
    // size classes: records grouped by class; batches ordered largest groups first (wide before packed)
    uint32_t at = 0;
    for (int cl = 0; cl < kNumClasses; ++cl) {
        tab.entry_base[cl] = at;
        tab.count[cl] = lg_count[cl];
        at += tab.count[cl];
    }
    for (int q = 0; q < kNumRL; ++q) {
        rl.entry_base[q] = at;
        at += rl.count[q];
    }
    uint32_t items_small = 0, items_mid = 0;
    for (int q = 0; q < kNumRL / 2; ++q) {  // q-th class in processing order: N = NMAX - 8 q
        rl.item_base_small[q] = items_small;
        items_small += (rl.count[kNumRL / 2 - 1 - q] + 31u) / 32u;
        rl.item_base_mid[q] = items_mid;
        items_mid += (rl.count[kNumRL - 1 - q] + 31u) / 32u;
    }
    rl.item_base_small[kNumRL / 2] = items_small;
    rl.item_base_mid[kNumRL / 2] = items_mid;
    uint32_t items = 0;
    int q = 0;
    for (int gi = kNumG - 1; gi >= 0; --gi) {
        for (int wide = 1; wide >= 0; --wide) {
            const int cl = gi + (wide ? kNumG : 0);
            const uint32_t G = class_lanes(gi), rpb = std::min(32u / G, kBufIntervals / (16u * G + 2u));
            tab.lanes[cl] = G;
            tab.rpb[cl] = rpb;
            tab.inv[cl] = (65536u + G - 1u) / G;
            tab.order[q] = (uint32_t)cl;
            tab.item_base[q] = items;
            items += (tab.count[cl] + rpb - 1u) / rpb;
            ++q;
        }
    }
    tab.item_base[kNumClasses] = items;
    pl.items = items;
    pl.items_small = items_small;
    pl.items_mid = items_mid;
    return pl;
}

}  // namespace

// The row-per-lane tier is opt-in (YB_RL_MAX_SLOTS=64 or 128; read at every launch): measured on B200 it does not
// beat the lane-group tier yet, so by default every register-tier row takes the lane-group path.
uint32_t rl_max_slots() {
    if (const char *e = getenv("YB_RL_MAX_SLOTS")) {
        const long v = strtol(e, nullptr, 10);
        return v < 0 ? 0u : v > (long)kRLMaxSlots ? kRLMaxSlots : (uint32_t)v;
    }
    return 0u;
}

int launch_row_stats(const uint32_t *rowptr, const uint32_t *len, uint32_t n_reads, DevRowStats *out, cudaStream_t stream) {
    if (cudaMemsetAsync(out, 0, sizeof(DevRowStats), stream) != cudaSuccess) return -1;
    if (n_reads == 0) return 0;
    row_stats_kernel<<<(n_reads + 1023u) / 1024u, 1024, 0, stream>>>(rowptr, len, n_reads, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_validate(const uint2 *iv, const uint32_t *rowptr, const uint32_t *len, uint32_t n_reads, uint32_t n_iv, DevRowStats *out,
                    cudaStream_t stream) {
    if (n_reads == 0 || n_iv == 0) return 0;
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    }
    uint32_t grid = (uint32_t)n_sm * 8u;
    const uint32_t want = (n_reads + 255u) / 256u;
    if (grid > want) grid = want;
    validate_kernel<<<grid, 256, 0, stream>>>(iv, rowptr, len, n_reads, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

uint64_t huge_keys_for_row(uint64_t k) {
    const uint64_t p = cta_words(k, true);
    return (k > kSmallMaxK && p > kCtaMaxSmemWords) ? p : 0;
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
    static int n_sm = 0, occ_sort = 0, occ_small = 0, occ_mid = 0;
    constexpr uint32_t kSmallThreads = 32u * YB_RL_WARPS_SMALL, kMidThreads = 32u * YB_RL_WARPS_MID;
    constexpr size_t kSmallSmem = YB_RL_WARPS_SMALL * rl_warp_smem(kRLSmallSlots), kMidSmem = YB_RL_WARPS_MID * rl_warp_smem(kRLMaxSlots);
    if (!n_sm) {
        int dev = 0, sm = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return -1;
        if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kCtaMaxSmemWords * sizeof(uint32_t))) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmemBytes) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sort, sort_kernel, kSortThreads, kSortSmemBytes) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(rl_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmem) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(rl_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMidSmem) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_small, rl_kernel<false>, kSmallThreads, kSmallSmem) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_mid, rl_kernel<true>, kMidThreads, kMidSmem) != cudaSuccess) return -1;
        if (occ_sort < 1 || occ_small < 1 || occ_mid < 1) return -1;
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
    const Plan pl = make_plan(a, coverage, rl_max_slots());
    const ClassTab &tab = pl.tab;
    const RLTab &rl = pl.rl;
    const uint32_t rl_max = pl.rl_max, items = pl.items, items_small = pl.items_small, items_mid = pl.items_mid;

    if (a.worklist_ready && rl_max == 0) {
        // the lane-group worklist does not depend on the threshold: launch_worklist built it once for this CSR
        if (cudaMemsetAsync(w.part_total, 0, sizeof(uint32_t) * w.n_parts, stream) != cudaSuccess) return -1;
    } else {
        scatter_kernel<<<(a.n_reads + kScatterRows - 1u) / kScatterRows, kScatterRows, 0, stream>>>(a, w, tab, rl, coverage, rl_max);
        ++launches;
    }
    bool forked = false;
    if (a.rows.n_big) {
        // shared memory for the largest row (wide if any read is); rows beyond kCtaMaxSmemWords sort in a global slab
        uint64_t words = cta_words(a.max_k, a.rows.n_wide != 0);
        if (words > kCtaMaxSmemWords) words = kCtaMaxSmemWords;
        uint32_t per_sm = (uint32_t)((220u * 1024u) / (words * 4u + 1024u));
        if (per_sm < 1u) per_sm = 1u;
        if (per_sm > 8u) per_sm = 8u;
        uint32_t grid = (uint32_t)n_sm * per_sm;
        if (grid > a.rows.n_big) grid = (uint32_t)a.rows.n_big;
        // the few long rows run on a side stream, beside the register tier (both only append to the staging buffer)
        forked = a.side_stream && a.ev_fork && a.ev_join && cudaEventRecord(a.ev_fork, stream) == cudaSuccess &&
                 cudaStreamWaitEvent(a.side_stream, a.ev_fork, 0) == cudaSuccess;
        big_kernel<<<grid, kCtaThreads, words * sizeof(uint32_t), forked ? a.side_stream : stream>>>(a, w, coverage, (uint32_t)words);
        ++launches;
        if (forked && cudaEventRecord(a.ev_join, a.side_stream) != cudaSuccess) return -1;
    }
    if (items) {
        uint32_t grid = (uint32_t)(n_sm * occ_sort);
        const uint32_t want = (items + kSortWarps - 1u) / kSortWarps;
        if (grid > want) grid = want;
        sort_kernel<<<grid, kSortThreads, kSortSmemBytes, stream>>>(a, w, tab, coverage);
        ++launches;
    }
    if (items_mid) {
        uint32_t grid = (uint32_t)(n_sm * occ_mid);
        const uint32_t want = (items_mid + YB_RL_WARPS_MID - 1u) / YB_RL_WARPS_MID;
        if (grid > want) grid = want;
        rl_kernel<true><<<grid, kMidThreads, kMidSmem, stream>>>(a, w, rl, coverage);
        ++launches;
    }
    if (items_small) {
        uint32_t grid = (uint32_t)(n_sm * occ_small);
        const uint32_t want = (items_small + YB_RL_WARPS_SMALL - 1u) / YB_RL_WARPS_SMALL;
        if (grid > want) grid = want;
        rl_kernel<false><<<grid, kSmallThreads, kSmallSmem, stream>>>(a, w, rl, coverage);
        ++launches;
    }
    if (forked && cudaStreamWaitEvent(stream, a.ev_join, 0) != cudaSuccess) return -1;
    scan_parts_kernel<<<1, 1024, 0, stream>>>(a, w);
    ++launches;
    order_kernel<<<w.n_parts, kPartRows, 0, stream>>>(a, w, not_coverage);
    ++launches;
    if (a.n_peers) {
        peer_barrier_kernel<<<1, 32, 0, stream>>>(a);
        ++launches;
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

int launch_worklist(const DetectArgs &a, cudaStream_t stream) {
    if (a.n_reads == 0) return 0;
    size_t total = 0;
    Work w = carve(a, a.rows.huge_keys, a.rows.n_big, a.rows.big_pairs, &total);
    if (total > a.scratch_bytes) return -1;
    if (cudaMemsetAsync(a.counters, 0, kNumCounters * sizeof(uint32_t), stream) != cudaSuccess) return -1;
    const Plan pl = make_plan(a, 0u, 0u);
    scatter_kernel<<<(a.n_reads + kScatterRows - 1u) / kScatterRows, kScatterRows, 0, stream>>>(a, w, pl.tab, pl.rl, 0u, 0u);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
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
