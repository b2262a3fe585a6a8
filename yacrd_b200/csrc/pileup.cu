// pileup.cu — sm_100a kernels of the detect hot path.
//
// Replaces, per read, FromOverlap::compute_bad_part (reference src/stack.rs:61-139) fused with
// editor::type_of_read (src/editor/mod.rs:85-100). The reference sorts the intervals and sweeps them
// with a min-heap of interval ends; the device computes the same bad-region list in closed form:
//
//   every interval (b,e) becomes two 32-bit event keys, 2b+1 (begin) and 2e (end). After an ascending
//   sort ends precede begins at equal positions — exactly the order in which the reference pops
//   `head <= begin` before pushing (stack.rs:72-81,90). depth = running (+1 begin / -1 end) sum.
//   With threshold c = `-c`:
//       up-crossing   U = begin event lifting depth c -> c+1   (the `stack.len() <= coverage` begin
//                         after which the heap is deeper than c, stack.rs:83)
//       down-crossing D = end event dropping depth c+1 -> c    (the last `last_covered = head`
//                         assignment before depth is back at c, stack.rs:77-79,93-105)
//   Crossings alternate U0 D0 U1 D1 ... and the cleaned gap list (stack.rs:107-138) is
//       [(0,U0) if U0 != 0] ++ [(D_t, U_t+1)] ++ [(D_last, len) if D_last != len]
//   or [(0,len) if len != 0] if depth never exceeds c. Written flat as u32, crossing number x simply
//   lands at flat[x + 2h - 1] with h = (U0 != 0) — no per-gap bookkeeping.
//   (tests/device_model.py is the executable form of this; tests/test_device_model.py fuzzes it
//   against the literal heap sweep.)
//
// Classification (editor/mod.rs:85-100): bad_len = sum(end-begin) in wrapping u32
//   = len + sum(U) - sum(D); NotCovered iff (double)bad_len / (double)len > n  (same IEEE divide);
//   else Chimeric iff there is an interior gap (begin != 0 && end != len) iff #U >= 2; else NotBad.
//
// Tiers (all integer work; no tensor cores — there is no contraction on this path):
//   warp tier    k <= 256 intervals: one warp per read, <= 16 event keys per lane in registers,
//                all-ascending bitonic network (register compare-exchange + shfl_xor), shuffle scan.
//   CTA tier     2k <= 16384 events: one CTA per read, keys in 64 KB of shared memory, ballot scans.
//   huge tier    anything bigger: same CTA body over u32 keys in a global scratch slab.
// then a 3-kernel exclusive scan compacts the per-read gap slots into a CSR and packs the 2-bit bitmap.
#include "pileup.cuh"

namespace yb {
namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr uint32_t kWarpMaxIntervals = 256;  // warp tier: <= 512 events in registers (16 per lane)
constexpr uint32_t kCtaMaxEvents = 16384;    // CTA tier: 64 KB of u32 event keys in shared memory
constexpr uint32_t kCtaThreads = 512;
constexpr uint32_t kScanItemsPerBlock = 4096;  // compaction: 256 threads x 16 reads

__host__ __device__ inline uint64_t next_pow2_u64(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// scratch carve-up (v1: padded bad-region slots + 3-kernel compaction)
struct Work {
    uint32_t *gap_cnt;      // n_reads
    uint2 *gaps_padded;     // n_iv + n_reads slots; read r owns [rowptr[r] + r, rowptr[r+1] + r + 1)
    uint32_t *big_list;     // reads deferred by the warp tier (capacity n_reads)
    uint32_t *block_sums;   // compaction scan partials
    uint32_t *huge_keys;    // global-scratch tier event keys
};

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

__device__ __forceinline__ void check_interval(const uint2 v, uint32_t len, uint32_t *counters) {
    if (!(v.x < v.y && v.y <= len)) atomicAdd(counters + kCntMalformed, 1u);
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t classify(uint32_t bad_len, uint32_t len, uint32_t n_up,
                                             double not_cov) {
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

// ------------------------------------------------------------------------------------------------
// warp tier
// ------------------------------------------------------------------------------------------------
// Sorts 32*EPL keys held EPL per lane, element index = lane*EPL + t, ascending. All comparators point
// the same way (the first step of every merge pairs idx with idx ^ (size-1)), so a register-register
// compare-exchange is a plain min/max pair and a cross-lane one is shfl_xor + one predicated min/max.
template <int EPL>
__device__ __forceinline__ void warp_bitonic_sort(uint32_t (&key)[EPL]) {
    const uint32_t lane = lane_id();
#pragma unroll
    for (int size = 2; size <= 32 * EPL; size <<= 1) {
        if (size <= EPL) {
#pragma unroll
            for (int t = 0; t < EPL; ++t) {
                const int u = t ^ (size - 1);
                if (t < u) ce(key[t], key[u]);
            }
        } else {
            const int lmask = size / EPL - 1;
            const bool keep_min = (lane & (uint32_t)(size / (2 * EPL))) == 0;
            uint32_t other[EPL];
#pragma unroll
            for (int t = 0; t < EPL; ++t) other[t] = __shfl_xor_sync(FULL, key[EPL - 1 - t], lmask);
#pragma unroll
            for (int t = 0; t < EPL; ++t)
                key[t] = keep_min ? min(key[t], other[t]) : max(key[t], other[t]);
        }
#pragma unroll
        for (int stride = size >> 2; stride > 0; stride >>= 1) {
            if (stride >= EPL) {
                const int ls = stride / EPL;
                const bool keep_min = (lane & (uint32_t)ls) == 0;
#pragma unroll
                for (int t = 0; t < EPL; ++t) {
                    const uint32_t o = __shfl_xor_sync(FULL, key[t], ls);
                    key[t] = keep_min ? min(key[t], o) : max(key[t], o);
                }
            } else {
#pragma unroll
                for (int t = 0; t < EPL; ++t)
                    if ((t & stride) == 0) ce(key[t], key[t | stride]);
            }
        }
    }
}

template <int EPL>
__device__ __forceinline__ void warp_pileup(const uint2 *__restrict__ row, uint32_t k, uint32_t len,
                                            uint32_t c, double not_cov, uint32_t *__restrict__ flat,
                                            uint8_t *__restrict__ cls_out,
                                            uint32_t *__restrict__ cnt_out, uint32_t *counters) {
    const uint32_t lane = lane_id();
    uint32_t key[EPL];
    if (EPL == 1) {
        const uint32_t j = lane & 15u;
        uint32_t kk = FULL;
        if (j < k) {
            const uint2 v = __ldg(row + j);
            if (lane < 16u) check_interval(v, len, counters);
            kk = lane < 16u ? 2u * v.x + 1u : 2u * v.y;
        }
        key[0] = kk;
    } else {
#pragma unroll
        for (int m = 0; m < EPL / 2; ++m) {
            const uint32_t j = lane + 32u * (uint32_t)m;
            uint32_t kb = FULL, ke = FULL;
            if (j < k) {
                const uint2 v = __ldg(row + j);
                check_interval(v, len, counters);
                kb = 2u * v.x + 1u;
                ke = 2u * v.y;
            }
            key[2 * m > EPL - 1 ? 0 : 2 * m] = kb;
            key[2 * m + 1 > EPL - 1 ? 0 : 2 * m + 1] = ke;
        }
    }
    warp_bitonic_sort<EPL>(key);

    const uint32_t n_ev = 2u * k, base = lane * (uint32_t)EPL;
    // pass 1: per-lane depth delta, then exclusive scan across lanes
    uint32_t delta = 0;
#pragma unroll
    for (int t = 0; t < EPL; ++t)
        if (base + t < n_ev) delta += (key[t] & 1u) ? 1u : FULL;
    const uint32_t depth0 = warp_incl_scan(delta) - delta;
    // pass 2: count crossings, accumulate sum(U) - sum(D)
    const uint32_t cu = c + 1u;
    uint32_t depth = depth0, ncross = 0, badsum = 0, firstpos = 0, lastpos = 0;
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
        if (base + t < n_ev) {
            const uint32_t isb = key[t] & 1u, pos = key[t] >> 1;
            depth += isb ? 1u : FULL;
            const bool up = isb && depth == cu, down = !isb && depth == c;
            if (up || down) {
                if (ncross == 0) firstpos = pos;
                lastpos = pos;
                ++ncross;
                badsum += up ? pos : 0u - pos;
            }
        }
    }
    const uint32_t xincl = warp_incl_scan(ncross);
    const uint32_t X = __shfl_sync(FULL, xincl, 31);
    badsum = warp_sum(badsum);
    uint32_t n_gaps, h = 0;
    if (X) {
        const uint32_t bal = __ballot_sync(FULL, ncross > 0);
        const uint32_t U0 = __shfl_sync(FULL, firstpos, __ffs(bal) - 1);
        const uint32_t Dl = __shfl_sync(FULL, lastpos, 31 - __clz(bal));
        h = U0 != 0u;
        n_gaps = (X >> 1) - 1u + h + (Dl != len ? 1u : 0u);
        // pass 3: crossing number x lands at flat[x + 2h - 1]
        uint32_t x = xincl - ncross;
        depth = depth0;
#pragma unroll
        for (int t = 0; t < EPL; ++t) {
            if (base + t < n_ev) {
                const uint32_t isb = key[t] & 1u, pos = key[t] >> 1;
                depth += isb ? 1u : FULL;
                if ((isb && depth == cu) || (!isb && depth == c)) {
                    const int idx = (int)(x + 2u * h) - 1;
                    if (idx >= 0) flat[idx] = pos;
                    ++x;
                }
            }
        }
        if (lane == 0) {
            if (h) flat[0] = 0u;
            if (Dl != len) flat[X + 2u * h - 1u] = len;
        }
    } else {
        n_gaps = len != 0u;
        if (lane == 0 && n_gaps) {
            flat[0] = 0u;
            flat[1] = len;
        }
    }
    if (lane == 0) {
        *cls_out = (uint8_t)classify(len + badsum, len, X >> 1, not_cov);
        *cnt_out = n_gaps;
    }
}

__global__ void __launch_bounds__(256) pileup_warp_kernel(DetectArgs a, Work w, uint32_t c,
                                                           double not_cov) {
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= a.n_reads) return;
    const uint32_t s = __ldg(a.rowptr + r), k = __ldg(a.rowptr + r + 1) - s;
    const uint32_t len = __ldg(a.len + r);
    if (k > kWarpMaxIntervals) {
        if (lane_id() == 0) w.big_list[atomicAdd(a.counters + kCntBigList, 1u)] = r;
        return;
    }
    const uint2 *row = a.iv + s;
    uint32_t *flat = reinterpret_cast<uint32_t *>(w.gaps_padded + (size_t)s + r);
    uint8_t *co = a.cls + r;
    uint32_t *no = w.gap_cnt + r;
    if (k <= 16u)
        warp_pileup<1>(row, k, len, c, not_cov, flat, co, no, a.counters);
    else if (k <= 32u)
        warp_pileup<2>(row, k, len, c, not_cov, flat, co, no, a.counters);
    else if (k <= 64u)
        warp_pileup<4>(row, k, len, c, not_cov, flat, co, no, a.counters);
    else if (k <= 128u)
        warp_pileup<8>(row, k, len, c, not_cov, flat, co, no, a.counters);
    else
        warp_pileup<16>(row, k, len, c, not_cov, flat, co, no, a.counters);
}

// ------------------------------------------------------------------------------------------------
// CTA tier (shared-memory keys) and generic tier (u64 keys in global scratch) share this body
// ------------------------------------------------------------------------------------------------
template <typename Key>
__device__ void cta_pileup(Key *ev, uint32_t n_pow2, const uint2 *__restrict__ row, uint32_t k,
                           uint32_t len, uint32_t c, double not_cov, uint32_t *__restrict__ flat,
                           uint8_t *__restrict__ cls_out, uint32_t *__restrict__ cnt_out,
                           uint32_t *sh /* 5 * 32 u32 */, uint32_t *counters) {
    const uint32_t tid = threadIdx.x, nthr = blockDim.x, n_ev = 2u * k;
    const uint32_t lane = lane_id(), wid = tid >> 5, nwarps = nthr >> 5;
    for (uint32_t i = tid; i < n_pow2; i += nthr) {
        Key kk = ~Key(0);
        if (i < n_ev) {
            const uint2 v = __ldg(row + (i >> 1));
            if (i & 1u) check_interval(v, len, counters);
            kk = (i & 1u) ? Key(v.y) * 2 : Key(v.x) * 2 + 1;
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
            const Key a = ev[lo], b = ev[hi];
            if (a > b) {
                ev[lo] = b;
                ev[hi] = a;
            }
        }
        __syncthreads();
        for (uint32_t stride = size >> 2; stride > 0; stride >>= 1) {
            for (uint32_t p = tid; p < half_n; p += nthr) {
                const uint32_t lo = 2u * stride * (p / stride) + (p % stride), hi = lo + stride;
                const Key a = ev[lo], b = ev[hi];
                if (a > b) {
                    ev[lo] = b;
                    ev[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    // Each warp owns a contiguous chunk of the sorted events and walks it 32 events per round;
    // depth inside a round comes from two ballots (begins, ends) and popc — no shuffles.
    uint32_t *sh_delta = sh, *sh_cross = sh + 32, *sh_first = sh + 64, *sh_last = sh + 96,
             *sh_bad = sh + 128;
    uint32_t chunk = n_pow2 / nwarps;
    if (chunk < 32u) chunk = 32u;
    const uint32_t beg = min(wid * chunk, n_ev), end = min(beg + chunk, n_ev);
    const uint32_t le = (2u << lane) - 1u, lt = (1u << lane) - 1u;
    uint32_t dsum = 0;
    for (uint32_t i0 = beg; i0 < end; i0 += 32u) {
        const uint32_t i = i0 + lane;
        const bool real = i < end;
        const uint32_t kb = real ? (uint32_t)(ev[i] & 1) : 0u;
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
        const Key kk = real ? ev[i] : Key(0);
        const uint32_t kb = real ? (uint32_t)(kk & 1) : 0u, pos = (uint32_t)(kk >> 1);
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
        for (uint32_t i0 = beg; i0 < end; i0 += 32u) {
            const uint32_t i = i0 + lane;
            const bool real = i < end;
            const Key kk = real ? ev[i] : Key(0);
            const uint32_t kb = real ? (uint32_t)(kk & 1) : 0u, pos = (uint32_t)(kk >> 1);
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

__global__ void __launch_bounds__(kCtaThreads) pileup_cta_kernel(DetectArgs a, Work w, uint32_t c,
                                                                  double not_cov) {
    extern __shared__ uint32_t ev_smem[];
    __shared__ uint32_t sh[160];
    __shared__ uint32_t sh_off;
    const uint32_t n_big = a.counters[kCntBigList];
    for (uint32_t i = blockIdx.x; i < n_big; i += gridDim.x) {
        const uint32_t r = w.big_list[i];
        const uint32_t s = a.rowptr[r], k = a.rowptr[r + 1] - s;
        const uint32_t n_pow2 = (uint32_t)next_pow2_u64(2ull * k);
        uint32_t *ev = ev_smem;
        if (n_pow2 > kCtaMaxEvents) {  // huge tier: keys live in a bump-allocated global slab
            if (threadIdx.x == 0) sh_off = atomicAdd(a.counters + kCntHugeBump, n_pow2);
            __syncthreads();
            ev = w.huge_keys + sh_off;
        }
        cta_pileup<uint32_t>(ev, n_pow2, a.iv + s, k, a.len[r], c, not_cov,
                             reinterpret_cast<uint32_t *>(w.gaps_padded + (size_t)s + r), a.cls + r,
                             w.gap_cnt + r, sh, a.counters);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// compaction: exclusive scan of gap counts -> gap_ptr, gather gaps, pack the 2-bit class bitmap
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kScanThreads = 256, kScanPerThread = kScanItemsPerBlock / kScanThreads;  // 16

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *sh /*33*/, uint32_t &total) {
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t incl = warp_incl_scan(v);
    if (lane == 31) sh[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint32_t ws = lane < nw ? sh[lane] : 0u;
        const uint32_t wi = warp_incl_scan(ws);
        sh[lane] = wi - ws;
        if (lane == 31) sh[32] = wi;
    }
    __syncthreads();
    const uint32_t r = sh[wid] + incl - v;
    total = sh[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_block_sums_kernel(const uint32_t *__restrict__ cnt,
                                                                        uint32_t n, uint32_t *block_sums) {
    __shared__ uint32_t sh[33];
    const uint32_t base = blockIdx.x * kScanItemsPerBlock + threadIdx.x * kScanPerThread;
    uint32_t s = 0;
#pragma unroll
    for (uint32_t i = 0; i < kScanPerThread; ++i)
        if (base + i < n) s += cnt[base + i];
    uint32_t total;
    block_excl_scan(s, sh, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_offsets_kernel(uint32_t *block_sums, uint32_t n_blocks,
                                                             uint32_t *gap_ptr_last) {
    __shared__ uint32_t sh[33];
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += blockDim.x) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_excl_scan(v, sh, total);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *gap_ptr_last = carry;
}

__global__ void __launch_bounds__(kScanThreads) compact_kernel(DetectArgs a, Work w) {
    __shared__ uint32_t sh[33];
    const uint32_t n = a.n_reads;
    const uint32_t base = blockIdx.x * kScanItemsPerBlock + threadIdx.x * kScanPerThread;
    uint32_t cnt[kScanPerThread];
    uint32_t s = 0;
#pragma unroll
    for (uint32_t i = 0; i < kScanPerThread; ++i) {
        cnt[i] = base + i < n ? w.gap_cnt[base + i] : 0u;
        s += cnt[i];
    }
    uint32_t total;
    uint32_t off = w.block_sums[blockIdx.x] + block_excl_scan(s, sh, total);
    uint32_t bits = 0, n_cls[3] = {0, 0, 0};
#pragma unroll
    for (uint32_t i = 0; i < kScanPerThread; ++i) {
        const uint32_t r = base + i;
        if (r < n) {
            a.gap_ptr[r] = off;
            const uint2 *src = w.gaps_padded + (size_t)a.rowptr[r] + r;
            for (uint32_t g = 0; g < cnt[i]; ++g) a.gaps[off + g] = src[g];
            off += cnt[i];
            const uint32_t cl = a.cls[r];
            bits |= cl << (2u * i);
            n_cls[cl < 3u ? cl : 0u]++;
        }
    }
    if (base < n) reinterpret_cast<uint32_t *>(a.bitmap)[base >> 4] = bits;  // 16 reads = 4 bytes
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const uint32_t v = warp_sum(n_cls[q]);
        if (lane_id() == 0 && v) atomicAdd(a.counters + kCntNotBad + q, v);
    }
}

// FromReport path: bad regions are given, only type_of_read (editor/mod.rs:85-100) runs. One thread
// takes 16 consecutive reads so it owns one 32-bit word of the 2-bit bitmap.
__global__ void __launch_bounds__(256) classify_kernel(const uint32_t *__restrict__ len,
                                                        const uint32_t *__restrict__ gap_ptr,
                                                        const uint2 *__restrict__ gaps, uint32_t n,
                                                        double not_cov, uint8_t *__restrict__ cls,
                                                        uint8_t *__restrict__ bitmap, uint32_t *counters) {
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

}  // namespace

int launch_classify(const uint32_t *len, const uint32_t *gap_ptr, const uint2 *gaps, uint32_t n_reads,
                    double not_coverage, uint8_t *cls, uint8_t *bitmap, uint32_t *counters,
                    cudaStream_t stream) {
    if (cudaMemsetAsync(counters, 0, kNumCounters * sizeof(uint32_t), stream) != cudaSuccess) return -1;
    if (n_reads == 0) return 0;
    const uint32_t threads = (n_reads + 15) / 16;
    classify_kernel<<<(threads + 255) / 256, 256, 0, stream>>>(len, gap_ptr, gaps, n_reads, not_coverage,
                                                               cls, bitmap, counters);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

uint64_t huge_keys_for_row(uint64_t k) {
    const uint64_t p = next_pow2_u64(2 * k);
    return p > kCtaMaxEvents ? p : 0;
}

static Work carve(const DetectArgs &a, uint64_t huge_keys, size_t *total) {
    Work w;
    size_t off = 0;
    char *base = static_cast<char *>(a.scratch);
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off += align256(bytes);
        return p;
    };
    const uint32_t n_blocks = (a.n_reads + kScanItemsPerBlock - 1) / kScanItemsPerBlock;
    w.gap_cnt = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * ((size_t)a.n_reads + 1)));
    w.gaps_padded = reinterpret_cast<uint2 *>(take(sizeof(uint2) * ((size_t)a.n_iv + a.n_reads + 1)));
    w.big_list = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * ((size_t)a.n_reads + 1)));
    w.block_sums = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * ((size_t)n_blocks + 1)));
    w.huge_keys = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * (huge_keys + 1)));
    *total = off;
    return w;
}

size_t detect_scratch_bytes(uint32_t n_reads, uint32_t n_iv, uint32_t max_k, uint64_t huge_keys) {
    DetectArgs a{};
    a.n_reads = n_reads;
    a.n_iv = n_iv;
    a.max_k = max_k;
    size_t total = 0;
    carve(a, huge_keys, &total);
    return total;
}

int launch_detect(const DetectArgs &a, uint32_t coverage, double not_coverage, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(pileup_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(kCtaMaxEvents * sizeof(uint32_t))) != cudaSuccess)
            return -1;
        attr_set = true;
    }
    int launches = 0;
    if (cudaMemsetAsync(a.counters, 0, kNumCounters * sizeof(uint32_t), stream) != cudaSuccess) return -1;
    if (a.n_reads == 0) {
        if (cudaMemsetAsync(a.gap_ptr, 0, sizeof(uint32_t), stream) != cudaSuccess) return -1;
        return 0;
    }
    size_t total = 0;
    // the huge-tier slab is whatever the engine sized after the fixed parts
    Work w = carve(a, 0, &total);
    const uint32_t c = coverage > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)coverage;
    pileup_warp_kernel<<<(a.n_reads + 7) / 8, 256, 0, stream>>>(a, w, c, not_coverage);
    ++launches;
    if (a.max_k > kWarpMaxIntervals) {
        pileup_cta_kernel<<<148 * 2, kCtaThreads, kCtaMaxEvents * sizeof(uint32_t), stream>>>(a, w, c,
                                                                                               not_coverage);
        ++launches;
    }
    const uint32_t n_blocks = (a.n_reads + kScanItemsPerBlock - 1) / kScanItemsPerBlock;
    scan_block_sums_kernel<<<n_blocks, kScanThreads, 0, stream>>>(w.gap_cnt, a.n_reads, w.block_sums);
    scan_offsets_kernel<<<1, 1024, 0, stream>>>(w.block_sums, n_blocks, a.gap_ptr + a.n_reads);
    compact_kernel<<<n_blocks, kScanThreads, 0, stream>>>(a, w);
    launches += 3;
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

}  // namespace yb
