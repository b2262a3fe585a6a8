// pileup.cuh — internal interface between the C-ABI engine (engine.cu) and the sm_100a kernels
// (detect.cu). Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace yb {

// A read with k intervals has 2k events (one begin key 2b+1, one end key 2e per interval); a u32
// event key needs positions < 2^31, which the engine enforces at upload (YB_ERR_TOO_LARGE).
constexpr uint32_t kMaxLength = 0x7FFFFFFFu;

// Device counters written by one detect step (u32 each).
enum Counter : uint32_t {
    kCntNotBad = 0,      // class histogram
    kCntChimeric = 1,
    kCntNotCovered = 2,
    kCntMalformed = 3,   // intervals violating 0 <= begin < end <= length
    kCntTile = 4,        // dynamic tile scheduler of the packed (fast) pass
    kCntBigList = 5,     // rows with more than 512 intervals (big tier)
    kCntHugeBump = 6,    // bump allocator (in u32 keys) of the global-scratch tier
    kCntBigBump = 7,     // bump allocator (in pairs) of the big tier's side buffer
    kCntTierWarp = 8,    // reads taken by each tier
    kCntTierCta = 9,
    kCntStage = 11,      // bump allocator (in pairs) of the bad-region staging buffer
    kCntTileSmall = 12,  // dynamic batch scheduler of the row-per-lane tier, rows of <= 64 slots
    kCntTileMid = 10,    // the same for rows of 72 .. 128 slots
    kCntStageOverflow = 14,  // rows whose bad regions did not fit the staging buffer (must stay 0)
    kCntPeerTimeout = 15,  // the peer barrier gave up waiting (a rank died or never launched)
    kCntTicket = 13,     // order_kernel: dynamic part index (decoupled look-back needs in-order starts)
    kCntClassCursor = 16,  // kNumClasses + kNumRL cursors of the worklist scatter
    kCntHist = 64,         // kHistSlots x {NotBad, Chimeric, NotCovered}: the detect step's class histogram, striped
    kNumCounters = 64 + 3 * 32
};
constexpr uint32_t kHistSlots = 32;


// Size classes of the register tier: a row with k intervals is sorted by G lanes x 16 keys, G the smallest
// entry with 16 G >= k. Classes 0..kNumG-1 hold rows whose positions fit 16 bits (packed u16x2 keys),
// kNumG..2 kNumG-1 the same sizes for longer reads (two u32 key arrays). Rows with k > 512 are "big".
constexpr int kNumG = 10;
constexpr int kNumClasses = 2 * kNumG;
__host__ __device__ inline uint32_t class_lanes(int gi) {
    return gi == 0 ? 1u : gi == 1 ? 2u : gi == 2 ? 3u : gi == 3 ? 4u : gi == 4 ? 5u : gi == 5 ? 6u : gi == 6 ? 8u : gi == 7 ? 10u : gi == 8 ? 16u : 32u;
}

// Rows whose length is <= kPackedMaxLen are sorted as packed u16x2 keys (begin | end << 16).
constexpr uint32_t kPackedMaxLen = 65534u;
constexpr uint32_t kRegisterTierMaxK = 512u;

// Row-per-lane tier: a packed row with k intervals at threshold c takes s = k + min(c, k) + 1 key slots (its
// intervals plus min(c, k) + 1 sentinel ends, see detect.cu); rows with s <= kRLMaxSlots are sorted by ONE lane in
// registers, 32 rows of one slot class per warp. Slot classes N = 8, 16, ..., 128.
constexpr int kNumRL = 16;
constexpr uint32_t kRLMaxSlots = 128u;
constexpr uint32_t kRLSmallSlots = 64u;  // classes up to here run in the low-register kernel
constexpr int kNumAllClasses = kNumClasses + kNumRL;
// Slot class (0 .. kNumRL-1) of a row in the row-per-lane tier, or -1 if the row does not belong there.
// max_slots (<= kRLMaxSlots) is where the tier ends: a tuning knob (YB_RL_MAX_SLOTS), 0 switches the tier off.
__host__ __device__ inline int rl_class_of_row(uint32_t k, uint32_t len, uint32_t c, uint32_t max_slots = kRLMaxSlots) {
    if (len > kPackedMaxLen || k >= max_slots) return -1;
    const uint32_t s = k + (c < k ? c : k) + 1u;
    return s <= max_slots ? (int)((s + 7u) / 8u) - 1 : -1;
}

// Size class of a row, or -1 for a big row (k > 512).
__host__ __device__ inline int class_of_row(uint32_t k, uint32_t len) {
    if (k > kRegisterTierMaxK) return -1;
    const int gi = k <= 16u ? 0 : k <= 32u ? 1 : k <= 48u ? 2 : k <= 64u ? 3 : k <= 80u ? 4 : k <= 96u ? 5 : k <= 128u ? 6 : k <= 160u ? 7 : k <= 256u ? 8 : 9;
    return gi + (len > kPackedMaxLen ? kNumG : 0);
}

// What the host knows about the rows at freeze time (sizes the scratch exactly).
struct RowStats {
    uint64_t n_big = 0;      // rows with k > 512 (big tier)
    uint64_t big_pairs = 0;  // sum over them of k + 1
    uint64_t huge_keys = 0;  // sum over rows beyond the shared-memory tier of next_pow2(2k)
    uint64_t n_wide = 0;     // rows longer than kPackedMaxLen (positions do not fit 16 bits)
    uint32_t class_count[kNumClasses] = {};  // rows per size class (k <= 512)
    uint32_t k_hist[kRLMaxSlots] = {};       // packed rows (len <= kPackedMaxLen) with exactly k < 128 intervals
};


struct DetectArgs {
    // CSR input, resident in HBM
    const uint2 *iv;         // flat (begin,end) buffer, n_iv entries
    const uint32_t *rowptr;  // n_reads + 1
    const uint32_t *len;     // n_reads
    uint32_t n_reads;
    uint32_t n_iv;
    uint32_t max_k;          // largest row (host knows it from the row pointers)
    uint32_t worklist_ready; // launch_worklist has filled the lane-group worklist of this CSR (scratch is untouched since)
    RowStats rows;
    // outputs, resident in HBM
    uint8_t *cls;            // n_reads, yb_read_type
    uint32_t *gap_ptr;       // n_reads + 1: exclusive scan of per-read bad-region counts
    uint2 *gaps;             // CSR of bad regions, capacity n_iv + n_reads
    uint8_t *bitmap;         // ceil(n_reads / 4) bytes rounded up to 4, 2 bits per read
    // peer-memory all-gather (n_peers == 0: off): this rank's slot in every rank's gather buffer, every rank's flags
    uint32_t n_peers, rank;
    uint8_t *peer_slot[16];  // peer p's gather buffer + rank * slot_bytes
    uint32_t *peer_flag[16]; // peer p's flag array (word q: last step rank q finished; word 31: own step counter)
    uint32_t *counters;      // kNumCounters
    // host side only: a second stream and two events so that the CTA tier (rows with k > 512) runs beside the register
    // tier instead of in front of it (null: same stream, one after the other)
    cudaStream_t side_stream;
    cudaEvent_t ev_fork, ev_join;
    // scratch
    void *scratch;
    size_t scratch_bytes;
};

// Row statistics gathered on the device at upload time (the host never loops over the rows of a bulk CSR).
struct DevRowStats {
    uint32_t class_count[kNumClasses];
    uint32_t n_big, n_wide, max_k;
    uint32_t bad_rowptr;           // rows with rowptr[r + 1] < rowptr[r]
    uint32_t bad_len;              // rows longer than kMaxLength
    uint32_t pad_;
    unsigned long long big_pairs;  // sum over big rows of k + 1
    unsigned long long huge_keys;  // sum over rows beyond the shared-memory tier of next_pow2(2k)
    uint32_t k_hist[kRLMaxSlots];  // packed rows with exactly k < 128 intervals
    uint32_t malformed;            // intervals violating 0 <= begin < end <= length (launch_validate)
    uint32_t pad2_[3];
};
// Zeroes *out and fills it from the device-resident rowptr / len (one kernel on `stream`). Returns launches or -1.
int launch_row_stats(const uint32_t *rowptr, const uint32_t *len, uint32_t n_reads, DevRowStats *out, cudaStream_t stream);

// Counts the intervals violating 0 <= begin < end <= length into out->malformed (one streaming kernel on `stream`,
// after launch_row_stats which zeroes *out). The CSR is immutable once uploaded, so this runs once per upload and the
// row-per-lane kernels do not repeat the test at every detect step. Returns launches or -1.
int launch_validate(const uint2 *iv, const uint32_t *rowptr, const uint32_t *len, uint32_t n_reads, uint32_t n_iv,
                    DevRowStats *out, cudaStream_t stream);

// Bytes of scratch launch_detect needs for a CSR of this shape.
size_t detect_scratch_bytes(uint32_t n_reads, uint32_t n_iv, const RowStats &rs);
// Per-row contributions to RowStats (the engine sums them over the rows at freeze time).
uint64_t huge_keys_for_row(uint64_t k);
uint64_t big_pairs_for_row(uint64_t k);

// Builds, once per uploaded CSR, the size-class worklist of the lane-group tier (16 bytes per read: row, first interval,
// k, class, length) in the scratch buffer. It depends on rowptr / len only, not on the threshold, so like the interval
// validation it belongs to the upload (the reference builds its read index, a hash map, while it ingests). With
// DetectArgs::worklist_ready set, launch_detect skips its scatter kernel. Returns launches or -1.
int launch_worklist(const DetectArgs &a, cudaStream_t stream);
// Where the opt-in row-per-lane tier ends (YB_RL_MAX_SLOTS, read at every call; 0 = off). A step that runs with the tier
// on rewrites the worklist with its own classes, so the caller must drop worklist_ready until the next upload.
uint32_t rl_max_slots();

// Enqueues one whole detect step on `stream`. Returns the number of kernel launches enqueued, or -1
// on a launch error.
int launch_detect(const DetectArgs &a, uint32_t coverage, double not_coverage, cudaStream_t stream);

// Classification only, over an existing bad-region CSR (FromReport path, stack.rs:176-257 +
// editor/mod.rs:85-100). Returns launches or -1.
int launch_classify(const uint32_t *len, const uint32_t *gap_ptr, const uint2 *gaps, uint32_t n_reads,
                    double not_coverage, uint8_t *cls, uint8_t *bitmap, uint32_t *counters,
                    cudaStream_t stream);

}  // namespace yb
