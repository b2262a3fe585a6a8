// pileup.cuh — internal interface between the C-ABI engine (engine.cu) and the sm_100a kernels
// (detect.cu). Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace yb {

// Read lengths (and so positions) stay below 2^31: the wide keys keep their top bit for +inf padding and the carry-chain
// compares subtract positions in 32 bits. The engine enforces it at upload (YB_ERR_TOO_LARGE).
constexpr uint32_t kMaxLength = 0x7FFFFFFFu;

// Device counters. The buffer holds TWO sets of kNumCounters words plus a few persistent words behind them: the detect
// step with (device-resident) step number e uses set e & 1; the last CTA of its ordering kernel zeroes the OTHER set for
// step e + 1 and then increments the step number, so a step needs no memset node and its counters stay readable until
// the end of the next step (yb_download reads set (e - 1) & 1 of the step number e it finds).
enum Counter : uint32_t {
    kCntNotBad = 0,      // class histogram (classify_kernel: FromReport path)
    kCntChimeric = 1,
    kCntNotCovered = 2,
    kCntMalformed = 3,   // intervals violating 0 <= begin < end <= length seen by the CTA tier
    kCntHugeBump = 6,    // bump allocator (in u32 keys) of the global-scratch tier
    kCntStage = 11,      // bump allocator (in pairs) of the bad-region staging buffer
    kCntTicket = 13,     // order_kernel: dynamic part index (a part only waits for parts that already run)
    kCntStageOverflow = 14,  // rows whose bad regions did not fit the staging buffer (must stay 0)
    kCntPeerTimeout = 15,  // a peer never signalled the previous step (a rank died or never launched)
    kCntDone = 16,       // order_kernel: parts finished (the last one closes the step)
    kCntLiteral = 17,    // cursor over the list of malformed rows (literal heap sweep)
    kCntDynTicket = 18,  // (two words, a double) sort_kernel: cursor over the dynamically dealt batches
    kCntScanTicket = 20, // bigscan_kernel: cursor over the rows of the scan list
    kCntHist = 32,       // kHistSlots x {NotBad, Chimeric, NotCovered}: the detect step's class histogram, striped
    kNumCounters = 32 + 3 * 32
};
constexpr uint32_t kHistSlots = 32;
// persistent words behind the two sets
constexpr uint32_t kCntEpoch = 2 * kNumCounters;          // number of detect steps finished on this buffer
constexpr uint32_t kCntBigList = 2 * kNumCounters + 1;    // upload: rows with more than 512 intervals (big tier)
constexpr uint32_t kCntLiteralList = 2 * kNumCounters + 2;  // validating step: rows with a malformed interval found so far
constexpr uint32_t kCntMalformedIv = 2 * kNumCounters + 6;  // validating step: malformed intervals found so far
constexpr uint32_t kCntLiteralLast = 2 * kNumCounters + 7;  // the two counts of the last finished validating step ...
constexpr uint32_t kCntMalformedLast = 2 * kNumCounters + 9;  // ... (the closing CTA of order_kernel moves them here)
constexpr uint32_t kCntScanList = 2 * kNumCounters + 4;     // upload: big rows of the position-scan path, the heavy ones (filled from the front of the list)
constexpr uint32_t kCntScanListBack = 2 * kNumCounters + 8; // ... and the light ones (filled from the back): bigscan_kernel takes the list in order, heavy rows first
constexpr uint32_t kCntOrderTimeout = 2 * kNumCounters + 5; // order_kernel: the closing CTA gave up waiting for the others
constexpr uint32_t kCntPeerTimeoutWait = 2 * kNumCounters + 3;  // peer_wait_kernel gave up (a rank died or never launched)
constexpr uint32_t kCntClassCursor = 2 * kNumCounters + 10;  // upload: kNumClasses cursors of the worklist scatter
constexpr uint32_t kCounterWords = 2 * kNumCounters + 10 + 24;


// Size classes of the register tier: a row with k intervals is sorted by G lanes x kE keys, G the smallest entry of
// kClassLanes with kE G >= k. Classes 0..kNumG-1 hold rows whose positions fit 16 bits (packed u16x2 keys),
// kNumG..2 kNumG-1 the same sizes for longer reads (two u32 key arrays). Rows with k > 512 are "big".
#ifndef YB_KE
#define YB_KE 32   // keys per lane: 32 (a row of up to 64 intervals crosses lanes once) or 16 (half the shared memory per warp)
#endif
constexpr int kE = YB_KE;
#if YB_KE == 32
constexpr int kNumG = 7;
__host__ __device__ inline uint32_t class_lanes(int gi) {
    return gi == 0 ? 1u : gi == 1 ? 2u : gi == 2 ? 3u : gi == 3 ? 4u : gi == 4 ? 5u : gi == 5 ? 8u : 16u;
}
#elif YB_KE == 16
constexpr int kNumG = 10;
__host__ __device__ inline uint32_t class_lanes(int gi) {
    return gi == 0 ? 1u : gi == 1 ? 2u : gi == 2 ? 3u : gi == 3 ? 4u : gi == 4 ? 5u : gi == 5 ? 6u : gi == 6 ? 8u : gi == 7 ? 10u : gi == 8 ? 16u : 32u;
}
#else
#error "YB_KE must be 16 or 32"
#endif
constexpr int kNumClasses = 2 * kNumG;
__host__ __device__ constexpr int class_lanes_c(int gi) {
#if YB_KE == 32
    return gi == 0 ? 1 : gi == 1 ? 2 : gi == 2 ? 3 : gi == 3 ? 4 : gi == 4 ? 5 : gi == 5 ? 8 : 16;
#else
    return gi == 0 ? 1 : gi == 1 ? 2 : gi == 2 ? 3 : gi == 3 ? 4 : gi == 4 ? 5 : gi == 5 ? 6 : gi == 6 ? 8 : gi == 7 ? 10 : gi == 8 ? 16 : 32;
#endif
}

// Rows whose length is <= kPackedMaxLen are sorted as packed u16x2 keys (begin | end << 16).
constexpr uint32_t kPackedMaxLen = 65534u;
constexpr uint32_t kRegisterTierMaxK = 512u;

// A big row (k > 512) is not sorted at all when its read is short enough: its begins and ends are counted per position in
// shared memory (16-bit counts) and the depth is scanned over the read's length, a window of kScanWindow positions at a
// time (4 bytes per position; at most kScanMaxWindows windows, beyond that sorting is cheaper).
constexpr uint32_t kScanWindow = 16384u, kScanMaxWindows = 24u, kScanMaxK = 65534u;
__host__ __device__ inline bool big_row_scans(uint32_t k, uint32_t len) {
    return k <= kScanMaxK && len < kScanWindow * kScanMaxWindows;
}

// Size class of a row, or -1 for a big row (k > 512).
__host__ __device__ inline int class_of_row(uint32_t k, uint32_t len) {
    if (k > kRegisterTierMaxK) return -1;
    int gi = 0;
    while ((uint32_t)kE * class_lanes(gi) < k) ++gi;
    return gi + (len > kPackedMaxLen ? kNumG : 0);
}

// What the host knows about the rows at freeze time (sizes the scratch exactly).
struct RowStats {
    uint64_t n_big = 0;      // rows with k > 512 (big tier), those of the scan path included
    uint64_t n_scan = 0;     // big rows that take the position-scan path (big_row_scans)
    uint32_t max_len_scan = 0;  // longest of them
    uint32_t max_k_sort = 0;    // largest row of the sort path of the big tier
    uint64_t big_pairs = 0;  // sum over them of k + 1
    uint64_t huge_keys = 0;  // sum over rows beyond the shared-memory tier of next_pow2(2k)
    uint64_t n_wide = 0;     // rows longer than kPackedMaxLen (positions do not fit 16 bits)
    uint32_t class_count[kNumClasses] = {};  // rows per size class (k <= 512)
};


struct DetectArgs {
    // CSR input, resident in HBM
    const uint2 *iv;         // flat (begin,end) buffer, n_iv entries
    const uint32_t *rowptr;  // n_reads + 1
    const uint32_t *len;     // n_reads
    uint32_t n_reads;
    uint32_t n_iv;
    uint32_t max_k;          // largest row (host knows it from the row pointers)
    uint32_t validate;       // 1: the step tests 0 <= begin < end <= length on every interval it loads and hands the rows that
                             // fail to literal_kernel (the first step after an upload, and every step if that one found any)
    RowStats rows;
    // outputs, resident in HBM
    uint8_t *cls;            // n_reads, yb_read_type
    uint32_t *gap_ptr;       // n_reads + 1: exclusive scan of per-read bad-region counts
    uint2 *gaps;             // CSR of bad regions, capacity n_iv + n_reads
    uint8_t *bitmap;         // ceil(n_reads / 4) bytes rounded up to 4, 2 bits per read
    // peer-memory all-gather (n_peers == 0: off): this rank's slot in every rank's gather buffer, every rank's flags
    uint32_t n_peers, rank;
    uint8_t *peer_slot[16];  // peer p's gather buffer + rank * slot_bytes (the slot of even steps)
    size_t peer_parity_bytes; // n_ranks * slot_bytes: odd steps use the second half of every gather buffer
    uint32_t *peer_flag[16]; // peer p's flag array (word q: steps rank q has finished; word 31: steps this rank has finished)
    uint32_t *counters;      // kCounterWords
    // host side only: a second stream and two events so that the CTA tier (rows with k > 512) runs beside the register
    // tier instead of in front of it (null: same stream, one after the other)
    cudaStream_t side_stream;
    cudaEvent_t ev_fork, ev_join;
    // scratch
    void *scratch;
    size_t scratch_bytes;
};

// Row statistics gathered on the device at upload time (the host loops over the rows only for the chunks of a streamed run).
struct DevRowStats {
    uint32_t class_count[kNumClasses];
    uint32_t n_big, n_wide, max_k;
    uint32_t n_scan, max_len_scan, max_k_sort, pad0_;
    uint32_t bad_rowptr;           // rows with rowptr[r + 1] < rowptr[r]
    uint32_t bad_len;              // rows longer than kMaxLength
    uint32_t pad_;
    unsigned long long big_pairs;  // sum over big rows of k + 1
    unsigned long long huge_keys;  // sum over rows beyond the shared-memory tier of next_pow2(2k)
    uint32_t pad2_[4];
};
// Zeroes *out and fills it from the device-resident rowptr / len (one kernel on `stream`). Returns launches or -1.
int launch_row_stats(const uint32_t *rowptr, const uint32_t *len, uint32_t n_reads, DevRowStats *out, cudaStream_t stream);
// The same from host arrays (only differences of the row pointers are used), for the streamed path.
void host_row_stats(const uint32_t *rowptr, const uint32_t *len, uint32_t n_reads, DevRowStats *out);

// Once per uploaded CSR (it depends on rowptr / len only, not on the threshold - the reference builds its read index, a
// hash map, while it ingests): scatter_kernel, the size-class worklist of the register tier (16 bytes per read: row,
// first interval, k, class, length) and the lists of the CTA tier. Also zeroes the counter buffer (kCounterWords).
// a.rows must hold the row statistics. Returns launches or -1.
// (The intervals themselves are tested by the first detect step: DetectArgs::validate.)
int launch_upload_kernels(const DetectArgs &a, cudaStream_t stream);

// Bytes of scratch launch_detect needs for a CSR of this shape.
size_t detect_scratch_bytes(uint32_t n_reads, uint32_t n_iv, const RowStats &rs);
// Per-row contribution to RowStats::huge_keys (the engine sums it over the rows at freeze time).
uint64_t huge_keys_for_row(uint64_t k);

// With peers bound: enqueues a wait for every rank's slot of the last finished step (the consumer side of the fused
// all-gather). Returns launches or -1.
int launch_peer_wait(const DetectArgs &a, cudaStream_t stream);

// Enqueues one whole detect step on `stream`. Returns the number of kernel launches enqueued, or -1
// on a launch error.
int launch_detect(const DetectArgs &a, uint32_t coverage, double not_coverage, cudaStream_t stream);

// Classification only, over an existing bad-region CSR (FromReport path, stack.rs:176-257 +
// editor/mod.rs:85-100). Returns launches or -1.
int launch_classify(const uint32_t *len, const uint32_t *gap_ptr, const uint2 *gaps, uint32_t n_reads,
                    double not_coverage, uint8_t *cls, uint8_t *bitmap, uint32_t *counters,
                    cudaStream_t stream);

}  // namespace yb
