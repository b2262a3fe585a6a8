// regtier.cuh — the register tier of the detect step (included by detect.cu, inside namespace yb::<anonymous>).
//
// Rows with k <= 512 intervals. G lanes share a row, kE = 32 keys per lane, blocked layout after the sort (element =
// g * 32 + t); a batch is floor(32 / G) rows of ONE size class (G = 1, 2, 3, 4, 5, 8, 16 -> at most 32, 64, 96, 128, 160,
// 256, 512 key slots), so every lane of the warp works and every group does the same thing. Everything the lane
// geometry decides is a template parameter: lane / G and lane % G are constants, the merge levels are unrolled, a slot
// carries its "beyond the row's end" test only where a row of the class can end.
//
//   PK  (len <= 65534): a key register holds begin | end << 16 and one VIMNMX.U16x2 moves a begin and an end through the
//       same network, so the two sorts of the closed form (detect.cu header) cost one;
//   !PK (longer reads): two u32 arrays, the network runs twice.
//
// Per batch: (1) striped conflict-free LDS.64 from the slab the TMA copies filled, pack; (2) the lane's 32 keys through
// 2 x the 60-exchange network for 16 keys + Batcher's odd-even merge 16 + 16 (65 exchanges): 185, the smallest known
// count for 32 keys; then ceil(log2 G) bitonic merge levels, the lane-bit stages by shuffle (power-of-two groups by XOR,
// the others take +inf from spare lane 31 for a partner that does not exist) — with 32 keys per lane a row of up to 64
// intervals needs ONE such stage; (3) the sorted keys go to a transposed shared copy T[t][lane], the crossing tests read
// the end that is c (+1) ranks below at a warp-uniform offset and push their carries into bit masks (IADD3 + IMAD.X);
// (4) the handful of crossings U0 D0 U1 D1 ... goes straight to a bump-allocated segment of the staging buffer as the
// pair list P[q] = (D_{q-1}, U_q) with D_{-1} = 0 and U_{n} = len: the row's bad regions are a sub-range of it, so per row
// only one 8-byte record {first pair, count} follows (order_kernel does the rest). The slab is dead as soon as the keys
// are in registers: the copies of the NEXT batch are issued into the same buffer right after the load phase and land
// while this batch is sorted (one slab + T + records + loop state = 14.3 KB per warp).
#pragma once

constexpr int kLogE = kE == 32 ? 5 : 4;
constexpr uint32_t kScrPitch2 = 33;          // T[t][lane] at scr[1 + 33 t + lane]; scr[0] stands for lane -1
constexpr uint32_t kScrWords2 = (uint32_t)kE * kScrPitch2 + 4u;

#ifndef YB_FMA_CE_MOD
#define YB_FMA_CE_MOD 0  // n > 0: every n-th in-lane compare-exchange computes its max as a + b - min with two IMADs (FMA pipe)
#endif
#ifndef YB_PACK_IMAD
#define YB_PACK_IMAD 1   // 1: begin | end << 16 and (begin << 16 | 0xFFFF) as IMADs (FMA pipe) instead of PRMTs (ALU pipe, the busy one)
#endif

template <int G> struct Geo {
    static constexpr int NP = G <= 1 ? 1 : G <= 2 ? 2 : G <= 4 ? 4 : G <= 8 ? 8 : G <= 16 ? 16 : 32;
    static constexpr bool kPow2 = NP == G;
    static constexpr int PITCH = kE * G + 2;  // row slots of the slab buffer (host: make_plan)
    static constexpr int RPB = (32 / G) < ((int)kBufIntervals / PITCH) ? (32 / G) : ((int)kBufIntervals / PITCH);
    // rows of the class have more than KMIN intervals (class_of_row picks the smallest class that fits); -1: any k >= 0
    static __host__ __device__ constexpr int prev_lanes(int gi = 0, int p = 0) {
        return gi >= kNumG ? p : class_lanes_c(gi) == G ? p : prev_lanes(gi + 1, class_lanes_c(gi));
    }
    static constexpr int KMIN = G == 1 ? -1 : kE * prev_lanes();
    static_assert(kPow2 || RPB * G <= 31, "a group that is not a power of two needs spare lane 31");
};

struct PipeMul {  // run-time 1, -1 and 65536 (kernel parameters): ptxas cannot strength-reduce the IMADs built on them
    uint32_t one, mone, shl16;
};

// n-th exchange of a network; `n` is a constant once the caller is unrolled
template <bool PK> __device__ __forceinline__ void ce_t(uint32_t &x, uint32_t &y, const PipeMul pm, const int n) {
    const uint32_t lo = PK ? __vminu2(x, y) : min(x, y);
    if (YB_FMA_CE_MOD != 0 && (n % (YB_FMA_CE_MOD ? YB_FMA_CE_MOD : 1)) == 0) {
        // max = x + y - min; packed: on both halves at once (what one half carries into the other cancels in the difference)
        uint32_t sum, hi;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(sum) : "r"(x), "r"(pm.one), "r"(y));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(lo), "r"(pm.mone), "r"(sum));
        y = hi;
    } else {
        y = PK ? __vmaxu2(x, y) : max(x, y);
    }
    x = lo;
}

// k[O .. O+16): 60 compare-exchanges in 10 layers (the smallest known network for 16 keys)
template <bool PK, int O> __device__ __forceinline__ void sort16_t(uint32_t (&k)[kE], const PipeMul pm) {
#define CE(n, i, j) ce_t<PK>(k[O + i], k[O + j], pm, n);
    CE(0, 0, 13) CE(1, 1, 12) CE(2, 2, 15) CE(3, 3, 14) CE(4, 4, 8) CE(5, 5, 6) CE(6, 7, 11) CE(7, 9, 10)
    CE(8, 0, 5) CE(9, 1, 7) CE(10, 2, 9) CE(11, 3, 4) CE(12, 6, 13) CE(13, 8, 14) CE(14, 10, 15) CE(15, 11, 12)
    CE(16, 0, 1) CE(17, 2, 3) CE(18, 4, 5) CE(19, 6, 8) CE(20, 7, 9) CE(21, 10, 11) CE(22, 12, 13) CE(23, 14, 15)
    CE(24, 0, 2) CE(25, 1, 3) CE(26, 4, 10) CE(27, 5, 11) CE(28, 6, 7) CE(29, 8, 9) CE(30, 12, 14) CE(31, 13, 15)
    CE(32, 1, 2) CE(33, 3, 12) CE(34, 4, 6) CE(35, 5, 7) CE(36, 8, 10) CE(37, 9, 11) CE(38, 13, 14)
    CE(39, 1, 4) CE(40, 2, 6) CE(41, 5, 8) CE(42, 7, 10) CE(43, 9, 13) CE(44, 11, 14)
    CE(45, 2, 4) CE(46, 3, 6) CE(47, 9, 12) CE(48, 11, 13)
    CE(49, 3, 5) CE(50, 6, 8) CE(51, 7, 9) CE(52, 10, 12)
    CE(53, 3, 4) CE(54, 5, 6) CE(55, 7, 8) CE(56, 9, 10) CE(57, 11, 12)
    CE(58, 6, 7) CE(59, 8, 9)
#undef CE
}

// the lane's 32 keys: two sorted halves, then Batcher's odd-even merge of 16 + 16 (65 exchanges)
template <bool PK> __device__ __forceinline__ void sort32_t(uint32_t (&k)[kE], const PipeMul pm) {
    sort16_t<PK, 0>(k, pm);
    if (kE < 32) return;
    sort16_t<PK, kE - 16>(k, pm);
#define CE(n, i, j) ce_t<PK>(k[(i) % kE], k[(j) % kE], pm, n + 1);
    CE(0, 0, 16) CE(1, 8, 24) CE(2, 8, 16) CE(3, 4, 20) CE(4, 12, 28) CE(5, 12, 20) CE(6, 4, 8) CE(7, 12, 16)
    CE(8, 20, 24) CE(9, 2, 18) CE(10, 10, 26) CE(11, 10, 18) CE(12, 6, 22) CE(13, 14, 30) CE(14, 14, 22) CE(15, 6, 10)
    CE(16, 14, 18) CE(17, 22, 26) CE(18, 2, 4) CE(19, 6, 8) CE(20, 10, 12) CE(21, 14, 16) CE(22, 18, 20) CE(23, 22, 24)
    CE(24, 26, 28) CE(25, 1, 17) CE(26, 9, 25) CE(27, 9, 17) CE(28, 5, 21) CE(29, 13, 29) CE(30, 13, 21) CE(31, 5, 9)
    CE(32, 13, 17) CE(33, 21, 25) CE(34, 3, 19) CE(35, 11, 27) CE(36, 11, 19) CE(37, 7, 23) CE(38, 15, 31) CE(39, 15, 23)
    CE(40, 7, 11) CE(41, 15, 19) CE(42, 23, 27) CE(43, 3, 5) CE(44, 7, 9) CE(45, 11, 13) CE(46, 15, 17) CE(47, 19, 21)
    CE(48, 23, 25) CE(49, 27, 29) CE(50, 1, 2) CE(51, 3, 4) CE(52, 5, 6) CE(53, 7, 8) CE(54, 9, 10) CE(55, 11, 12)
    CE(56, 13, 14) CE(57, 15, 16) CE(58, 17, 18) CE(59, 19, 20) CE(60, 21, 22) CE(61, 23, 24) CE(62, 25, 26) CE(63, 27, 28)
    CE(64, 29, 30)
#undef CE
}

// bitonic half-cleaners on the slot bits (the lane's 32 keys form a bitonic sequence): 80 exchanges
template <bool PK> __device__ __forceinline__ void clean32_t(uint32_t (&k)[kE], const PipeMul pm) {
#pragma unroll
    for (int s = kE >> 1; s > 0; s >>= 1) {
#pragma unroll
        for (int t = 0; t < kE; ++t)
            if ((t & s) == 0) ce_t<PK>(k[t], k[t | s], pm, t + (t >> 3) + s);
    }
}

// One exchange stage between lanes: lane g with lane g ^ M of its group (FLIP: my slot t against its slot kE - 1 - t);
// M is a run-time value so that every level of the merge runs through the same code (the bodies of the larger classes
// would not fit the instruction cache otherwise). The min / max choice is a per-lane predicate; ptxas turns it into two
// VIMNMX into fresh registers plus two predicated moves per key whatever the source looks like (select, if / else,
// predicated inline PTX) and never emits the single instruction with a predicate operand, which is why 32 keys per lane
// pay: a row of up to 64 intervals crosses lanes once.
template <int G, bool FLIP, bool PK>
__device__ __forceinline__ void xlane_t(uint32_t (&key)[kE], const uint32_t M, const uint32_t lane, const uint32_t g, const bool in_group) {
    const uint32_t HB = FLIP ? (M + 1u) >> 1 : M;  // the bit that tells the lower lane of a pair from the upper
    uint32_t src;
    bool keep_min;
    if (Geo<G>::kPow2) {
        src = lane ^ M;
        keep_min = (lane & HB) == 0u;
    } else {  // the network of NP lanes whose lanes G .. NP-1 hold +inf: spare lane 31 stands for them
        const uint32_t partner = g ^ M;
        const bool ex = in_group && partner < (uint32_t)G;
        src = ex ? lane + partner - g : 31u;
        keep_min = !ex || (g & HB) == 0u;
    }
    if (FLIP) {
#pragma unroll
        for (int t = 0; t < kE / 2; ++t) {
            const uint32_t o_hi = __shfl_sync(FULL, key[kE - 1 - t], src), o_lo = __shfl_sync(FULL, key[t], src);
            if (keep_min) {
                key[t] = PK ? __vminu2(key[t], o_hi) : min(key[t], o_hi);
                key[kE - 1 - t] = PK ? __vminu2(key[kE - 1 - t], o_lo) : min(key[kE - 1 - t], o_lo);
            } else {
                key[t] = PK ? __vmaxu2(key[t], o_hi) : max(key[t], o_hi);
                key[kE - 1 - t] = PK ? __vmaxu2(key[kE - 1 - t], o_lo) : max(key[kE - 1 - t], o_lo);
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < kE; ++t) {
            const uint32_t o = __shfl_sync(FULL, key[t], src);
            if (keep_min) key[t] = PK ? __vminu2(key[t], o) : min(key[t], o);
            else key[t] = PK ? __vmaxu2(key[t], o) : max(key[t], o);
        }
    }
}

// Sorts, for every group of G consecutive lanes, its kE G keys (ascending in element order g * kE + t): the lane's own keys,
// then ceil(log2 G) bitonic merge levels (rolled: one copy of the flip stage, of the lane-bit stage and of the slot-bit
// half-cleaners serves every level).
template <int G, bool PK>
__device__ __forceinline__ void sort_group_t(uint32_t (&key)[kE], const uint32_t lane, const uint32_t g, const bool in_group, const PipeMul pm) {
    sort32_t<PK>(key, pm);
    constexpr uint32_t NP = (uint32_t)Geo<G>::NP;
    if (NP < 2u) return;
#pragma unroll 1
    for (uint32_t ls = 2u; ls <= NP; ls <<= 1) {
        xlane_t<G, true, PK>(key, ls - 1u, lane, g, in_group);
#pragma unroll 1
        for (uint32_t m = ls >> 2; m > 0u; m >>= 1) xlane_t<G, false, PK>(key, m, lane, g, in_group);
        clean32_t<PK>(key, pm);
    }
}

struct alignas(16) WarpSmem {  // one per warp: a warp runs on its own, no CTA-wide barrier anywhere
    unsigned long long mbar;
    unsigned long long pad_;
    uint4 rec[2][32];  // worklist records of the next two batches, one per lane (cp.async)
    uint32_t geo[2][32];  // and the lanes' places in those batches (fetch_rec)
    uint32_t st[16];      // the batch loop's own state while a batch is sorted (sort_kernel: every register is needed there,
                          // and what ptxas spills to local memory comes back from L2, ~270 cycles a piece)
    uint32_t scr[kScrWords2];
};
static_assert(sizeof(WarpSmem) % 16 == 0, "the slab must stay 16-byte aligned");
constexpr size_t kWarpSmemBytes = sizeof(WarpSmem) + sizeof(uint2) * kBufIntervals;

// (inline PTX: around a plain atomicAdd in `if (lane == 0)` nvcc builds a warp-aggregated atomic whose result is broadcast
// by a shuffle right behind it, which makes the warp wait for the L2 round trip on the spot)
__device__ __forceinline__ uint32_t atom_add_u32(uint32_t *p, uint32_t v) {
    uint32_t r;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
    return r;
}

__device__ __forceinline__ uint32_t top_bits(uint32_t n) { return n >= 32u ? FULL : ~(FULL >> n); }  // the n most significant bits

// One batch: every in-group lane holds its row's record (all G lanes of a group hold the same one). `refill` issues the
// next batch's copies into the slab; it is called as soon as this batch's keys are in registers.
template <int G, bool PK, bool VAL, class Refill>
__device__ __forceinline__ void process_batch_t(const DetectArgs &a, const Work &w, uint32_t *cnt, WarpSmem &ws, uint2 *buf, const uint4 rec,
                                                const uint32_t c, uint2 &chunk, const PipeMul pm, const uint32_t lane, Refill refill) {
    using GG = Geo<G>;
    const uint32_t j = lane / (uint32_t)G, g0 = lane % (uint32_t)G;
    const bool in_group = j < (uint32_t)GG::RPB;
    const uint32_t g = in_group ? g0 : 0u;
    bool valid = in_group && (rec.z & kRecValid);
    const uint32_t k = valid ? (rec.z & 0xFFFFu) : 0u, len = rec.w;
    uint32_t nbad = 0;  // VAL: this lane's intervals that violate 0 <= begin < end <= length
    uint2 *slot = buf + (in_group ? j : 0u) * (uint32_t)GG::PITCH + (rec.y & 1u);  // the row's data starts here
    // striped load (conflict-free); the initial arrangement is irrelevant to the sort. Element t * G + g exists iff
    // t * G < k - g; the test is only compiled for slots a row of this class can end in.
    // PK: begin | end << 16. Long reads (!PK) sort their ends first (-> T) and then their begins in the SAME registers:
    // the slab is read twice and refilled after the second load.
    uint32_t K0[kE];
    const uint2 *lane_iv = slot + g;
    const uint32_t left = k > g ? k - g : 0u;
    auto load_keys = [&](const bool ends) {
#pragma unroll
        for (int t = 0; t < kE; ++t) {
            const uint2 v = lane_iv[t * G];
            // (spare lanes - 31 above all - hold +inf for xlane_t when the group is not a power of two)
            const bool absent = (t * G + G - 1 > GG::KMIN && !((uint32_t)(t * G) < left)) || (!GG::kPow2 && !in_group);
            if (VAL && (PK || ends) && !absent) nbad += !(v.x < v.y && v.y <= len);
            uint32_t key;
            if (!PK) key = ends ? v.y : v.x;
            else if (YB_PACK_IMAD) asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(key) : "r"(v.y), "r"(pm.shl16), "r"(v.x));
            else key = __byte_perm(v.x, v.y, 0x5410);
            K0[t] = absent ? INF : key;
        }
    };
    // transposed copy of the sorted ends (PK: of the packed keys; the compares only look at the end half): T[t][lane];
    // element kE l + t - c - 1 is then T[(t - c - 1) % kE][l + floor((t - c - 1) / kE)], a warp-uniform offset from the
    // lane's own column: conflict-free writes and reads, no per-element index arithmetic
    uint32_t *T = ws.scr + 1u + lane;
    auto validated = [&]() {
        // A validating step (the first one after an upload): a row with a malformed interval is not this kernel's business,
        // the closed form is only the reference's heap sweep for well-formed rows. It goes on the list of literal_kernel.
        if (!valid) nbad = 0;
        const uint32_t bb = __ballot_sync(FULL, nbad != 0u);
        if (bb) {  // rare
            const uint32_t gmask = G == 32 ? FULL : (((1u << (G & 31)) - 1u) << (lane - g));
            if (nbad) atomicAdd(a.counters + kCntMalformedIv, nbad);
            if (valid && (bb & gmask)) {
                if (g == 0u) w.lit_list[atomicAdd(a.counters + kCntLiteralList, 1u)] = rec.x;
                valid = false;
            }
        }
    };
    if (PK) {
        load_keys(false);
        refill();
        if (VAL) validated();
        sort_group_t<G, PK>(K0, lane, g, in_group, pm);
    } else {
        load_keys(true);
        if (VAL) validated();
        sort_group_t<G, PK>(K0, lane, g, in_group, pm);
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < kE; ++t) T[kScrPitch2 * t] = K0[t];
    if (!PK) {
        load_keys(false);
        refill();
        sort_group_t<G, PK>(K0, lane, g, in_group, pm);
    }
    __syncwarp();
    // V1_t = (E[32g + t - c - 1] <= B_t), t = 0..32;  V0_t = (E[32g + t - c] <= B_t), t = 0..31.
    // PK: (end_j <= begin_i)  <=>  key_j <= (begin_i << 16 | 0xFFFF) as plain u32.
    // Built most-significant-first: m1 bit (31 - t) = V1_t (t < 32), v1n = V1_32, m0 bit (31 - t) = V0_t.
    const uint32_t cc = min(c, (uint32_t)(kE * 32 + kE));  // beyond k every threshold behaves the same
    uint32_t m1 = 0, m0 = 0, v1n = 0;
    {
        uint32_t Knext = __shfl_down_sync(FULL, K0[0], 1);
        if (g == (uint32_t)G - 1u) Knext = INF;
        uint32_t qp = 0;
#pragma unroll
        for (int t = 0; t <= kE; ++t) {
            const int jr = t - (int)cc - 1;  // uniform
            int col = jr >> kLogE;
            if (cc >= (uint32_t)kE) col = max(col, -(int)lane - 1);  // stay inside scr; those elements are forced below
            const uint32_t ev = T[(int)kScrPitch2 * (jr & (kE - 1)) + col];
            const uint32_t kt = t < kE ? K0[t % kE] : Knext;
            uint32_t q;
            if (!PK) q = kt;
            else if (YB_PACK_IMAD) asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(q) : "r"(kt), "r"(pm.shl16), "r"(0xFFFFu));
            else q = __byte_perm(kt, FULL, 0x1044);
            if (t < kE) m1 = push_le(m1, ev, q);
            else v1n = push_le(0u, ev, q);
            if (t > 0) m0 = push_le(m0, ev, qp);
            qp = q;
        }
        // elements below the row's first end are 0 (E[-1] = 0): V1_t true for 32g + t <= c, V0_t for 32g + t < c
        m1 <<= 32 - kE;
        m0 <<= 32 - kE;
        const int z = (int)cc - kE * (int)g;
        if (z >= 0) {
            m1 |= top_bits((uint32_t)z + 1u);
            m0 |= top_bits((uint32_t)z);
            if (z >= kE) v1n = 1u;
        }
    }
    // kE bits were pushed: move them to the top of the word, then bit (31 - t): U at begin t = V1_t & !V0_t; D at end t =
    // !V0_t & V1_{t+1}
    uint32_t um = (m1 & ~m0) & top_bits(kE), dm = (((m1 << 1) | (v1n << (32 - kE))) & ~m0) & top_bits(kE);
    if (!valid) um = dm = 0;
    // ranks of this lane's crossings among the row's ups / downs (packed segmented scan over the group)
    const uint32_t mine = __popc(um) | (__popc(dm) << 16);
    uint32_t incl = mine;
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, incl, off);
        if (g >= (uint32_t)off) incl += o;
    }
    const uint32_t tot = G == 1 ? incl : __shfl_sync(FULL, incl, min(lane - g + (uint32_t)G - 1u, 31u));
    uint32_t ru = (incl - mine) & 0xFFFFu, rd = (incl - mine) >> 16;
    const uint32_t n_up = tot & 0xFFFFu;
    // first up-crossing and last down-crossing of the row decide where its bad-region list starts and ends
    // (stack.rs:107-113): every lane knows its own first U / last D, the group's come from its first / last lane that has one
    uint32_t U0, Dl;
    {
        uint32_t fu = 0, ld = 0;
        if (PK) {
            if (um) fu = T[(int)kScrPitch2 * __clz(um)] & 0xFFFFu;
        } else {
#pragma unroll
            for (int t = kE - 1; t >= 0; --t) fu = (um >> (31 - t)) & 1u ? K0[t] : fu;
        }
        if (dm) {
            const int jr = 32 - __ffs(dm) - (int)cc;  // slot t = 31 - (ffs - 1)
            ld = T[(int)kScrPitch2 * (jr & (kE - 1)) + (jr >> kLogE)];
            if (PK) ld >>= 16;
        }
        if (G == 1) {
            U0 = fu;
            Dl = ld;
        } else {
            const uint32_t gmask = G == 32 ? FULL : (((1u << (G & 31)) - 1u) << (lane - g));
            const uint32_t bu = __ballot_sync(FULL, um != 0u) & gmask, bd = __ballot_sync(FULL, dm != 0u) & gmask;
            U0 = __shfl_sync(FULL, fu, bu ? (uint32_t)__ffs(bu) - 1u : lane);
            Dl = __shfl_sync(FULL, ld, bd ? 31u - (uint32_t)__clz(bd) : lane);
        }
    }
    // The row's list P[q] = (D_{q-1}, U_q), q = 0 .. n_up, with D_{-1} = 0 and U_{n_up} = len, takes n_up + 1 pairs of the
    // staging buffer; the bad regions are P[q0 .. q0 + ng) with q0 = (U0 == 0) and ng = n_up + (D_last != len) - q0; a row that
    // never rises above c has P[0] = (0, len). The warp owns a chunk of the staging buffer and refills it with one atomic
    // when it runs out. Every lane of a group carries the group's numbers, so a scan over the groups needs the strides
    // G, 2G, 4G, ... only.
    const uint32_t need = valid ? n_up + 1u : 0u;
    uint32_t inc = need;
#pragma unroll
    for (int off = G; off < 32; off <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, inc, off);
        if (lane >= (uint32_t)off) inc += o;
    }
    const uint32_t total = __shfl_sync(FULL, inc, (uint32_t)(GG::RPB * G - 1));
    uint32_t base;
    if (total <= chunk.y - chunk.x) {
        base = chunk.x;
        chunk.x += total;
    } else {
        const bool direct = total >= kStageChunk / 4u;  // a large batch takes exactly what it needs
        uint32_t got = 0;
        if (lane == 0) got = atom_add_u32(cnt + kCntStage, direct ? total : kStageChunk);
        base = __shfl_sync(FULL, got, 0);
        if (!direct) chunk = make_uint2(base + total, base + kStageChunk);
    }
    base += inc - need;
    if (!valid) return;
    if ((uint64_t)base + need > w.stage_cap) {  // cannot happen with the capacity the engine allocates; never write outside
        if (g == 0u) {
            atom_add_u32(cnt + kCntStageOverflow, 1u);
            w.meta[rec.x] = make_uint2(0u, 0u);
        }
        return;
    }
    uint32_t *P = reinterpret_cast<uint32_t *>(w.stage + base);
    if (PK) {
        while (um) {  // sparse: a row has a handful of crossings
            const int t = __clz(um);
            um &= ~(0x80000000u >> t);
            P[2u * ru++ + 1u] = T[(int)kScrPitch2 * t] & 0xFFFFu;
        }
        while (dm) {
            const int t = __clz(dm);
            dm &= ~(0x80000000u >> t);
            const int jr = t - (int)cc;
            P[2u * rd++ + 2u] = T[(int)kScrPitch2 * (jr & (kE - 1)) + (jr >> kLogE)] >> 16;
        }
    } else if (um | dm) {
#pragma unroll
        for (int t = 0; t < kE; ++t) {
            if (um & (0x80000000u >> t)) P[2u * ru++ + 1u] = K0[t];
            if (dm & (0x80000000u >> t)) {
                const int jr = t - (int)cc;
                P[2u * rd++ + 2u] = T[(int)kScrPitch2 * (jr & (kE - 1)) + (jr >> kLogE)];
            }
        }
    }
    if (g == 0u) {
        P[0] = 0u;
        P[2u * n_up + 1u] = len;
        const uint32_t q0 = n_up ? (U0 == 0u) : 0u;
        const uint32_t ng = n_up ? n_up + (Dl != len) - q0 : (len != 0u);
        w.meta[rec.x] = make_uint2(base + q0, ng);
    }
}
