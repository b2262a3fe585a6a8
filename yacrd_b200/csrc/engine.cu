// engine.cu — the C ABI (include/yacrd_b200.h): host-side mirror of the reference's Reads2Ovl producer
// (src/reads2ovl/mod.rs:43-163, FullMemory semantics src/reads2ovl/fullmemory.rs:46-90) and BadPart
// consumer (src/stack.rs:35-41,143-173), with the device boundary placed where the reference calls
// get_overlaps (stack.rs:149): host store -> CSR -> HBM -> sm_100a kernels (pileup.cu) -> classes and
// bad-region CSR back to the host. No CPU implementation of the pile-up exists in this library: without a
// CUDA device yb_create fails.
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/yacrd_b200.h"
#include "pileup.cuh"
#include "store.hpp"

namespace {

thread_local std::string g_create_error;

const char *const kTypeNames[3] = {"NotBad", "Chimeric", "NotCovered"};  // editor/mod.rs:51-58

template <typename T>
struct PinnedBuf {  // page-locked host buffer, grown geometrically (plain memory in a host-only context)
    T *p = nullptr;
    size_t cap = 0;
    bool plain = false;
    bool reserve(size_t n) {
        if (n <= cap) return true;
        release();
        size_t want = n + n / 8 + 16;
        if (plain) {
            p = static_cast<T *>(malloc(want * sizeof(T)));
            if (!p) return false;
        } else if (cudaMallocHost(reinterpret_cast<void **>(&p), want * sizeof(T)) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            cap = 0;
            return false;
        }
        cap = want;
        return true;
    }
    bool grow_keep(size_t n, size_t used) {  // like reserve, but the first `used` elements survive
        if (n <= cap) return true;
        PinnedBuf<T> bigger;
        bigger.plain = plain;
        if (!bigger.reserve(n + n / 4)) return false;
        if (used) memcpy(bigger.p, p, used * sizeof(T));
        release();
        p = bigger.p;
        cap = bigger.cap;
        bigger.p = nullptr;
        return true;
    }
    void release() {
        if (p) {
            if (plain) free(p);
            else cudaFreeHost(p);
        }
        p = nullptr;
        cap = 0;
    }
};

template <typename T>
struct DeviceBuf {
    T *p = nullptr;
    size_t cap = 0;
    bool reserve(size_t n) {
        if (n <= cap) return true;
        release();
        size_t want = n + n / 16 + 64;
        if (cudaMalloc(reinterpret_cast<void **>(&p), want * sizeof(T)) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            cap = 0;
            return false;
        }
        cap = want;
        return true;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

inline char *put_u64(char *p, uint64_t v) {
    char tmp[24];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

}  // namespace

struct yb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side_stream = nullptr;          // the CTA tier (rows with k > 512) runs here, beside the register tier
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    uint32_t read_buffer_size = 8192;
    uint32_t flags = 0;
    uint32_t ingest_threads = 0;
    bool host_only = false;
    cudaEvent_t ev_upload = nullptr;   // last upload work on `stream` (a compute on another stream waits for it)
    cudaEvent_t ev_compute = nullptr;  // last detect step enqueued on a caller's stream (yb_download waits for it)
    bool compute_foreign = false;      // ev_compute is pending
    // The first detect step after an upload tests every interval it loads (0 <= begin < end <= length); rows that fail go
    // to the literal heap sweep. Its two counts come back behind the step (4 + 4 bytes); if they are zero the next steps
    // run the plain kernels.
    bool validated = false;            // the counts of a validating step over this CSR are known
    bool valid_pending = false;        // a validating step is in flight; ev_valid follows its counts' copy
    cudaEvent_t ev_valid = nullptr;
    PinnedBuf<uint32_t> h_valid;
    uint32_t n_literal = 0, n_malformed = 0;
    bool bulk_frozen = false;  // the CSR came straight from the parallel ingester: `pending` does not hold it
    // Streamed batches (yb_set_chunk_intervals): the host CSR goes through the device in row chunks of about this many
    // intervals, on two lanes (child contexts with their own streams and device buffers), so that chunk k + 1 crosses
    // PCIe while chunk k is computed and chunk k - 1's results come back.
    uint32_t chunk_intervals = 0;
    std::vector<yb_ctx *> lanes;
    bool is_lane = false;
    uint32_t rowptr_base = 0;  // lane: the chunk's first row pointer (subtracted on the device)
    uint32_t n_chunks_last = 0;
    std::string error;

    // ---- host store (Reads2Ovl producer side) ----
    yb::IdTable ids;                       // named mode: id -> dense first-seen index
    std::vector<uint64_t> length;          // per read (usize in the reference)
    std::vector<yb::PendingRecord> pending;  // arrival-order intervals not yet frozen
    bool indexed = false;                  // reads are named by their decimal index (bulk CSR input)
    uint32_t n_indexed = 0;
    // borrowed CSR (yb_bind_csr): caller-owned host buffers used in place
    const uint32_t *b_rowptr = nullptr, *b_iv = nullptr, *b_len = nullptr;

    // ---- frozen CSR (pinned host) ----
    PinnedBuf<uint32_t> h_rowptr, h_len;
    PinnedBuf<uint2> h_iv;
    std::vector<uint32_t> arrival_rowptr;  // KEEP_HOST_INTERVALS
    uint32_t n_reads = 0, n_iv = 0, max_k = 0;
    yb::RowStats rows;
    bool frozen = false, uploaded = false, computed = false, downloaded = false, from_report = false;

    // ---- device ----
    DeviceBuf<uint32_t> d_rowptr, d_len, d_gap_ptr, d_counters;
    DeviceBuf<uint2> d_iv, d_gaps;
    DeviceBuf<uint8_t> d_cls, d_bitmap, d_scratch;
    DeviceBuf<yb::DevRowStats> d_rowstats;
    PinnedBuf<yb::DevRowStats> h_rowstats;
    PinnedBuf<uint32_t> h_peer_step;
    uint8_t *ext_bitmap = nullptr;  // yb_bind_device_bitmap
    size_t ext_bitmap_bytes = 0;
    // peer-memory all-gather (yb_bind_peers)
    uint32_t n_peers = 0, peer_rank = 0;
    size_t peer_slot_bytes = 0;
    uint8_t *peer_gather[YB_MAX_PEERS] = {};
    uint32_t *peer_flags[YB_MAX_PEERS] = {};

    // ---- results (pinned host) ----
    PinnedBuf<uint8_t> h_cls, h_bitmap;
    PinnedBuf<uint32_t> h_gap_ptr, h_counters;
    PinnedBuf<uint2> h_gaps;
    uint32_t n_gaps = 0;
    double not_coverage = 0.8;
    uint64_t coverage = 0;

    yb_stats stats{};
    std::string scratch_id;  // yb_read_at in indexed mode

    int fail(int code, const char *fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        error = buf;
        return code;
    }
    int cuda_fail(cudaError_t e, const char *what) {
        cudaGetLastError();
        return fail(YB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    size_t bitmap_bytes() const { return (((size_t)n_reads + 15) / 16) * 4; }
    uint32_t total_reads() const { return indexed ? n_indexed : ids.size(); }
    void invalidate() { frozen = uploaded = computed = downloaded = false; }
    const uint32_t *rowptr_host() const { return b_rowptr ? b_rowptr : h_rowptr.p; }
    const uint32_t *len_host() const { return b_len ? b_len : h_len.p; }
    const uint2 *iv_host() const { return b_iv ? reinterpret_cast<const uint2 *>(b_iv) : h_iv.p; }
};

namespace {

#define YB_CUDA(ctx, call)                                   \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) return (ctx)->cuda_fail(e_, #call); \
    } while (0)

bool parse_index_name(const char *id, size_t n, uint64_t *out) {
    if (n == 0 || n > 10) return false;
    uint64_t v = 0;
    for (size_t i = 0; i < n; ++i) {
        const unsigned d = (unsigned)(id[i] - '0');
        if (d > 9) return false;
        v = v * 10 + d;
    }
    if (n > 1 && id[0] == '0') return false;
    *out = v;
    return true;
}

int64_t find_read(const yb_ctx *c, const char *id, size_t n) {
    if (c->indexed) {
        uint64_t v;
        if (!parse_index_name(id, n, &v) || v >= c->n_indexed) return -1;
        return (int64_t)v;
    }
    const uint32_t i = c->ids.find(id, n);
    return i == yb::IdTable::kNone ? -1 : (int64_t)i;
}

// A CSR built by the parallel ingester goes back to arrival-order records when the caller keeps adding to it.
void thaw(yb_ctx *c) {
    if (!c->bulk_frozen) return;
    const uint32_t *rp = c->h_rowptr.p;
    c->pending.reserve((size_t)c->n_iv + 16);
    for (uint32_t r = 0; r < c->n_reads; ++r)
        for (uint32_t j = rp[r]; j < rp[r + 1]; ++j) c->pending.push_back({r, c->h_iv.p[j].x, c->h_iv.p[j].y});
    c->bulk_frozen = false;
}

int intern_read(yb_ctx *c, const char *id, size_t n, bool *is_new, uint32_t *idx) {
    thaw(c);
    if (c->indexed || c->from_report)
        return c->fail(YB_ERR_STATE, "this context holds bulk/report input; per-record adds are not allowed");
    if (c->ids.size() == 0xFFFFFFFEu) return c->fail(YB_ERR_TOO_LARGE, "too many reads");
    *idx = c->ids.intern(id, n, is_new);
    if (*is_new) c->length.push_back(0);
    return YB_OK;
}

// Freeze: counting sort of the arrival-order records by read -> CSR in pinned memory. Within a read the
// arrival order is kept (the kernels sort anyway; yb_overlap shows arrival order like Reads2Ovl::overlap).
int freeze(yb_ctx *c) {
    if (c->frozen) return YB_OK;
    if (c->b_rowptr) {  // borrowed CSR: nothing to build; the rows are inspected on the device at upload
        const uint32_t n = c->n_indexed;
        if (n && c->b_rowptr[0] != c->rowptr_base) return c->fail(YB_ERR_INVALID_ARGUMENT, "rowptr[0] must be 0");
        c->n_reads = n;
        c->n_iv = n ? c->b_rowptr[n] - c->rowptr_base : 0;
        c->frozen = true;
        return YB_OK;
    }
    const uint32_t n = c->ids.size();
    const size_t m = c->pending.size();
    if (m > 0xFFFFFFF0ull) return c->fail(YB_ERR_TOO_LARGE, "more than 2^32-16 intervals in one context");
    if (!c->h_rowptr.reserve((size_t)n + 1) || !c->h_len.reserve((size_t)n + 1) || !c->h_iv.reserve(m + 1))
        return c->fail(YB_ERR_NOMEM, "pinned host allocation failed");
    uint32_t *rp = c->h_rowptr.p;
    memset(rp, 0, sizeof(uint32_t) * ((size_t)n + 1));
    for (size_t i = 0; i < m; ++i) rp[c->pending[i].read + 1]++;
    for (uint32_t r = 0; r < n; ++r) rp[r + 1] += rp[r];
    std::vector<uint32_t> cur(rp, rp + n);
    for (size_t i = 0; i < m; ++i) {
        const yb::PendingRecord &p = c->pending[i];
        c->h_iv.p[cur[p.read]++] = make_uint2(p.begin, p.end);
    }
    for (uint32_t r = 0; r < n; ++r) {
        if (c->length[r] > yb::kMaxLength)
            return c->fail(YB_ERR_TOO_LARGE, "read %u is longer than 2^31-1 bases", r);
        c->h_len.p[r] = (uint32_t)c->length[r];
    }
    c->n_reads = n;
    c->n_iv = (uint32_t)m;
    c->frozen = true;
    return YB_OK;
}

int ensure_result_buffers(yb_ctx *c) {
    const size_t n = c->n_reads;
    if (!c->d_cls.reserve(n + 16) || !c->d_gap_ptr.reserve(n + 1) || !c->d_bitmap.reserve(c->bitmap_bytes() + 4) ||
        !c->d_counters.reserve(yb::kCounterWords) || !c->d_gaps.reserve((size_t)c->n_iv + n + 1))
        return c->fail(YB_ERR_NOMEM, "device allocation failed (%zu reads, %u intervals)", n, c->n_iv);
    return YB_OK;
}

}  // namespace

extern "C" {

const char *yb_version(void) { return YB_VERSION; }
const char *yb_type_name(int t) { return (t >= 0 && t < 3) ? kTypeNames[t] : "?"; }
const char *yb_create_error(void) { return g_create_error.c_str(); }
const char *yb_last_error(const yb_ctx *ctx) { return ctx ? ctx->error.c_str() : "null context"; }

// Opens the context's device (first CUDA call of the process: context creation, about half a second), its stream and the
// side stream. yb_create does it unless YB_FLAG_LAZY_DEVICE asks to leave it to the first call that needs the device.
static int open_device(yb_ctx *c) {
    if (c->stream) return YB_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return c->fail(YB_ERR_CUDA, "no usable CUDA device (the detect path has no CPU fallback): %s",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    int dev = c->device;
    if (dev < 0) {
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    }
    if (dev >= count || cudaSetDevice(dev) != cudaSuccess) {
        cudaGetLastError();
        return c->fail(YB_ERR_CUDA, "invalid CUDA device ordinal %d", dev);
    }
    c->device = dev;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        cudaGetLastError();
        c->stream = nullptr;
        return c->fail(YB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    // optional: without them the tiers simply run one after the other
    if (cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        if (c->side_stream) cudaStreamDestroy(c->side_stream);
        c->side_stream = nullptr;
    }
    if (cudaEventCreateWithFlags(&c->ev_upload, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_valid, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_compute, cudaEventDisableTiming) != cudaSuccess) {
        e = cudaGetLastError();
        return c->fail(YB_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
    }
    return YB_OK;
}
#define YB_DEVICE(ctx)                                   \
    do {                                                 \
        if (int rc_ = open_device(ctx)) return rc_;      \
    } while (0)

int yb_device_warmup(int device) {
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        return YB_ERR_CUDA;
    }
    return cudaFree(nullptr) == cudaSuccess ? YB_OK : YB_ERR_CUDA;
}

yb_ctx *yb_create(const yb_opts *opts) {
    if (opts && (opts->flags & YB_FLAG_HOST_ONLY)) {  // producer side only: no device, no compute
        yb_ctx *c = new (std::nothrow) yb_ctx();
        if (!c) {
            g_create_error = "out of memory";
            return nullptr;
        }
        c->host_only = true;
        c->device = -1;
        if (opts->read_buffer_size) c->read_buffer_size = opts->read_buffer_size;
        c->flags = opts->flags;
        c->ingest_threads = opts->ingest_threads;
        c->h_rowptr.plain = c->h_len.plain = c->h_iv.plain = true;
        return c;
    }
    yb_ctx *c = new (std::nothrow) yb_ctx();
    if (!c) {
        g_create_error = "out of memory";
        return nullptr;
    }
    c->device = opts ? opts->device : -1;
    if (opts) {
        if (opts->read_buffer_size) c->read_buffer_size = opts->read_buffer_size;
        c->flags = opts->flags;
        c->ingest_threads = opts->ingest_threads;
    }
    if (opts && (opts->flags & YB_FLAG_LAZY_DEVICE)) return c;  // the device is opened by the first call that needs it
    if (open_device(c) != YB_OK) {
        g_create_error = c->error;
        delete c;
        return nullptr;
    }
    return c;
}

void yb_destroy(yb_ctx *c) {
    if (!c) return;
    for (yb_ctx *l : c->lanes) yb_destroy(l);
    c->lanes.clear();
    if (c->host_only) {
        c->h_rowptr.release();
        c->h_len.release();
        c->h_iv.release();
        delete c;
        return;
    }
    if (c->stream) {  // (a lazy context that never needed its device has nothing to release there)
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    if (c->side_stream) {
        cudaStreamSynchronize(c->side_stream);
        cudaStreamDestroy(c->side_stream);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_upload) cudaEventDestroy(c->ev_upload);
    if (c->ev_valid) cudaEventDestroy(c->ev_valid);
    if (c->ev_compute) cudaEventDestroy(c->ev_compute);
    c->h_peer_step.release();
    c->h_rowptr.release();
    c->h_len.release();
    c->h_iv.release();
    c->h_cls.release();
    c->h_bitmap.release();
    c->h_gap_ptr.release();
    c->h_counters.release();
    c->h_gaps.release();
    c->d_rowptr.release();
    c->d_len.release();
    c->d_gap_ptr.release();
    c->d_counters.release();
    c->d_iv.release();
    c->d_gaps.release();
    c->d_cls.release();
    c->d_bitmap.release();
    c->d_scratch.release();
    c->d_rowstats.release();
    c->h_valid.release();
    c->h_rowstats.release();
    delete c;
}

// Page-locked when a CUDA driver is present; otherwise (CPU-only box: tests of the generator and of the
// host logic) plain aligned memory, remembered so that yb_host_free releases it the right way.
static std::mutex g_host_mu;
static std::unordered_set<void *> g_host_plain;

// Drop the store and the results but keep every host/device allocation: the next get_overlaps batch
// (stack.rs:148-161 loops over batches) reuses them.
int yb_reset(yb_ctx *c) {
    if (!c) return YB_ERR_INVALID_ARGUMENT;
    if (!c->host_only && c->stream) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
    }
    c->ids.clear();
    c->length.clear();
    c->pending.clear();
    c->indexed = false;
    c->n_indexed = 0;
    c->b_rowptr = c->b_iv = c->b_len = nullptr;
    c->n_reads = c->n_iv = c->max_k = c->n_gaps = 0;
    c->rows = yb::RowStats();
    c->from_report = false;
    c->bulk_frozen = false;
    c->invalidate();
    return YB_OK;
}

void *yb_host_alloc(size_t n_bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, n_bytes ? n_bytes : 1) == cudaSuccess) return p;
    cudaGetLastError();
    p = aligned_alloc(256, ((n_bytes ? n_bytes : 1) + 255) & ~(size_t)255);
    if (p) {
        std::lock_guard<std::mutex> lk(g_host_mu);
        g_host_plain.insert(p);
    }
    return p;
}
void yb_host_free(void *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        auto it = g_host_plain.find(p);
        if (it != g_host_plain.end()) {
            g_host_plain.erase(it);
            free(p);
            return;
        }
    }
    cudaFreeHost(p);
}

// ---- producer side ---------------------------------------------------------------------------------
int yb_add_overlap_and_length(yb_ctx *c, const char *id, size_t id_len, uint32_t begin, uint32_t end,
                              uint64_t length) {
    if (!c || (!id && id_len)) return YB_ERR_INVALID_ARGUMENT;
    bool is_new = false;
    uint32_t idx = 0;
    if (int rc = intern_read(c, id, id_len, &is_new, &idx)) return rc;
    if (is_new) c->length[idx] = length;  // fullmemory.rs:82-90: first-seen length wins
    c->pending.push_back({idx, begin, end});
    c->invalidate();
    return YB_OK;
}

int yb_add_overlap(yb_ctx *c, const char *id, size_t id_len, uint32_t begin, uint32_t end) {
    if (!c || (!id && id_len)) return YB_ERR_INVALID_ARGUMENT;
    bool is_new = false;
    uint32_t idx = 0;
    if (int rc = intern_read(c, id, id_len, &is_new, &idx)) return rc;  // fullmemory.rs:68-76: length 0
    c->pending.push_back({idx, begin, end});
    c->invalidate();
    return YB_OK;
}

int yb_add_length(yb_ctx *c, const char *id, size_t id_len, uint64_t length) {
    if (!c || (!id && id_len)) return YB_ERR_INVALID_ARGUMENT;
    bool is_new = false;
    uint32_t idx = 0;
    if (int rc = intern_read(c, id, id_len, &is_new, &idx)) return rc;
    c->length[idx] = length;  // fullmemory.rs:78-80: sets unconditionally
    c->invalidate();
    return YB_OK;
}

int yb_add_csr(yb_ctx *c, const uint32_t *rowptr, const uint32_t *iv, const uint32_t *length, uint32_t n_reads,
               const char *const *ids, const size_t *id_lens) {
    if (!c || !rowptr || !length || (!iv && n_reads && rowptr[n_reads] != rowptr[0]))
        return YB_ERR_INVALID_ARGUMENT;
    if (c->from_report || c->b_rowptr) return c->fail(YB_ERR_STATE, "context already holds report/borrowed input");
    if (!ids) {
        // index-named bulk input: copy straight into the pinned CSR (appending)
        if (c->ids.size()) return c->fail(YB_ERR_STATE, "index-named bulk input cannot follow named reads");
        const uint32_t old_n = c->n_indexed;
        const uint64_t old_m = old_n ? c->h_rowptr.p[old_n] : 0;
        const uint64_t add_m = (uint64_t)rowptr[n_reads] - rowptr[0];
        if (old_m + add_m > 0xFFFFFFF0ull || (uint64_t)old_n + n_reads > 0xFFFFFFF0ull)
            return c->fail(YB_ERR_TOO_LARGE, "more than 2^32-16 intervals or reads in one context");
        PinnedBuf<uint32_t> nr, nl;
        PinnedBuf<uint2> ni;
        const size_t tn = (size_t)old_n + n_reads;
        if (tn + 1 > c->h_rowptr.cap || old_m + add_m + 1 > c->h_iv.cap) {
            if (!nr.reserve(tn + 1) || !nl.reserve(tn + 1) || !ni.reserve(old_m + add_m + 1))
                return c->fail(YB_ERR_NOMEM, "pinned host allocation failed");
            if (old_n) {
                memcpy(nr.p, c->h_rowptr.p, sizeof(uint32_t) * ((size_t)old_n + 1));
                memcpy(nl.p, c->h_len.p, sizeof(uint32_t) * old_n);
                memcpy(ni.p, c->h_iv.p, sizeof(uint2) * old_m);
            } else {
                nr.p[0] = 0;
            }
            c->h_rowptr.release();
            c->h_len.release();
            c->h_iv.release();
            c->h_rowptr = nr;
            c->h_len = nl;
            c->h_iv = ni;
        } else if (!old_n) {
            c->h_rowptr.p[0] = 0;
        }
        for (uint32_t r = 0; r < n_reads; ++r) {
            if (rowptr[r + 1] < rowptr[r]) return c->fail(YB_ERR_INVALID_ARGUMENT, "rowptr is not monotone at read %u", r);
            if (length[r] > yb::kMaxLength) return c->fail(YB_ERR_TOO_LARGE, "read %u is longer than 2^31-1 bases", r);
            c->h_rowptr.p[old_n + r + 1] = (uint32_t)(old_m + (rowptr[r + 1] - rowptr[0]));
            c->h_len.p[old_n + r] = length[r];
        }
        if (add_m) memcpy(c->h_iv.p + old_m, iv + 2 * (size_t)rowptr[0], sizeof(uint2) * add_m);
        c->indexed = true;
        c->n_indexed = (uint32_t)tn;
        c->n_reads = (uint32_t)tn;
        c->n_iv = (uint32_t)(old_m + add_m);
        c->invalidate();
        c->frozen = true;
        return YB_OK;
    }
    if (c->indexed) return c->fail(YB_ERR_STATE, "named reads cannot follow index-named bulk input");
    for (uint32_t r = 0; r < n_reads; ++r) {
        const size_t idl = id_lens ? id_lens[r] : strlen(ids[r]);
        bool is_new;
        uint32_t idx;
        if (int rc = intern_read(c, ids[r], idl, &is_new, &idx)) return rc;
        if (is_new) c->length[idx] = length[r];
        for (uint32_t j = rowptr[r]; j < rowptr[r + 1]; ++j) c->pending.push_back({idx, iv[2 * (size_t)j], iv[2 * (size_t)j + 1]});
    }
    c->invalidate();
    return YB_OK;
}

int yb_bind_csr(yb_ctx *c, const uint32_t *rowptr, const uint32_t *iv, const uint32_t *length, uint32_t n_reads) {
    if (!c || !rowptr || !length || (!iv && n_reads && rowptr[n_reads])) return YB_ERR_INVALID_ARGUMENT;
    if (c->ids.size() || c->from_report || (c->indexed && !c->b_rowptr))
        return c->fail(YB_ERR_STATE, "yb_bind_csr needs an empty context");
    c->b_rowptr = rowptr;
    c->b_iv = iv;
    c->b_len = length;
    c->indexed = true;
    c->n_indexed = n_reads;
    c->invalidate();
    return YB_OK;
}

static bool add_sink(void *sink, const char *id, size_t id_len, uint32_t b, uint32_t e, uint64_t len) {
    return yb_add_overlap_and_length(static_cast<yb_ctx *>(sink), id, id_len, b, e, len) == YB_OK;
}

static bool alloc_csr_sink(void *sink, size_t n_reads, size_t n_iv, uint32_t **rowptr, uint32_t **len, uint32_t **iv) {
    yb_ctx *c = static_cast<yb_ctx *>(sink);
    if (!c->host_only && c->device >= 0) cudaSetDevice(c->device);  // (lazy context: page-lock on its device, not on device 0)
    if (!c->h_rowptr.reserve(n_reads + 1) || !c->h_len.reserve(n_reads + 1) || !c->h_iv.reserve(n_iv + 1)) return false;
    *rowptr = c->h_rowptr.p;
    *len = c->h_len.p;
    *iv = reinterpret_cast<uint32_t *>(c->h_iv.p);
    return true;
}

int yb_init_buffer(yb_ctx *c, const char *text, size_t n_bytes, int format) {
    if (!c || (!text && n_bytes) || (format != 'p' && format != 'm')) return YB_ERR_INVALID_ARGUMENT;
    yb::IngestError err;
    // An empty context and more than a few records: tokenize, intern and build the CSR on all host cores.
    int threads = (int)c->ingest_threads;
    if (threads == 0) {
        threads = (int)std::thread::hardware_concurrency();
        if (threads > 32) threads = 32;
        if (n_bytes < (1u << 20)) threads = 1;
    }
    if (threads >= 1 && n_bytes >= (1u << 16) && c->total_reads() == 0 && !c->indexed && !c->from_report && c->pending.empty()) {
        yb::BulkIds ids;
        if (!yb::ingest_buffer_parallel(text, n_bytes, format, threads, alloc_csr_sink, c, &ids, &err))
            return c->fail(err.code, "%s", err.message.c_str());
        c->length = std::move(ids.length);
        c->ids.adopt(std::move(ids.bytes), std::move(ids.off));
        c->n_reads = ids.n_reads;
        c->n_iv = (uint32_t)ids.n_iv;
        c->invalidate();
        c->frozen = true;
        c->bulk_frozen = true;
        return YB_OK;
    }
    if (!yb::ingest_buffer(text, n_bytes, format, add_sink, c, &err)) {
        if (err.code == YB_ERR_NOMEM && !c->error.empty()) return YB_ERR_STATE;  // add refused (state)
        return c->fail(err.code, "%s", err.message.c_str());
    }
    return YB_OK;
}

int yb_file_type(const char *path) {  // util.rs:39-55, same precedence
    if (!path) return 0;
    const std::string f(path);
    auto has = [&](const char *s) { return f.find(s) != std::string::npos; };
    if (has(".m4") || has(".mhap")) return 'm';
    if (has(".paf")) return 'p';
    if (has(".yacrd")) return 'y';
    if (has(".fastq") || has(".fq")) return 'q';
    if (has(".fasta") || has(".fa")) return 'a';
    if (has(".yovl")) return 'o';
    return 0;
}

static int slurp(yb_ctx *c, const char *path, std::vector<char> *out) {
    FILE *fh = fopen(path, "rb");
    if (!fh) return c->fail(YB_ERR_CANT_READ_FILE, "Can't open file %s: %s", path, strerror(errno));
    std::vector<char> buf((size_t)c->read_buffer_size > 65536 ? c->read_buffer_size : 65536);
    size_t got;
    while ((got = fread(buf.data(), 1, buf.size(), fh)) > 0) out->insert(out->end(), buf.data(), buf.data() + got);
    const bool bad = ferror(fh);
    fclose(fh);
    if (bad) return c->fail(YB_ERR_CANT_READ_FILE, "Error while reading %s", path);
    return YB_OK;
}

int yb_init_file(yb_ctx *c, const char *path) {
    if (!c || !path) return YB_ERR_INVALID_ARGUMENT;
    const int t = yb_file_type(path);  // reads2ovl/mod.rs:51-79
    if (t == 0 || t == 'o') return c->fail(YB_ERR_UNKNOWN_FORMAT, "Format detection for '%s' file not possible", path);
    if (t != 'p' && t != 'm')
        return c->fail(YB_ERR_WRONG_FORMAT, "Can't run overlap parsing on %s file %s",
                       t == 'y' ? "yacrd" : (t == 'q' ? "fastq" : "fasta"), path);
    // regular files are parsed in place from a read-only mapping; pipes and the like are read into memory
    const char *data = nullptr;
    size_t size = 0;
    void *map = nullptr;
    std::vector<char> text;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return c->fail(YB_ERR_CANT_READ_FILE, "Can't open file %s: %s", path, strerror(errno));
    struct stat sb;
    if (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
        map = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (map != MAP_FAILED) {
            madvise(map, (size_t)sb.st_size, MADV_SEQUENTIAL | MADV_WILLNEED);
            data = static_cast<const char *>(map);
            size = (size_t)sb.st_size;
        } else {
            map = nullptr;
        }
    }
    close(fd);
    if (!map) {
        if (int rc = slurp(c, path, &text)) return rc;
        data = text.data();
        size = text.size();
    }
    int rc = YB_OK;
    const size_t map_size = size;
    std::vector<char> inflated;
    const bool packed = (size >= 2 && (unsigned char)data[0] == 0x1f && (unsigned char)data[1] == 0x8b) ||
                        (size >= 3 && !memcmp(data, "BZh", 3)) || (size >= 6 && !memcmp(data, "\xfd" "7zXZ\0", 6));
    if (packed) {
        // util.rs:57-72 (niffler sniffs the magic number): gzip / bzip2 / xz input is inflated into memory, then parsed
        // like a plain file
        yb::Codec codec;
        std::string why;
        yb::ByteSource *src = yb::open_source(path, &codec, &why);
        if (!src) rc = c->fail(YB_ERR_CANT_READ_FILE, "%s", why.c_str());
        else {
            std::vector<char> buf(1u << 22);
            long got;
            while ((got = src->read(buf.data(), buf.size())) > 0) inflated.insert(inflated.end(), buf.data(), buf.data() + got);
            if (got < 0) rc = c->fail(YB_ERR_CANT_READ_FILE, "Error in compression detection of file %s: corrupt compressed stream", path);
            delete src;
            data = inflated.data();
            size = inflated.size();
        }
    }
    if (rc == YB_OK) {
        rc = yb_init_buffer(c, data, size, t);
        if (rc != YB_OK) c->error += std::string(" (Filename: ") + path + ")";
    }
    if (map) munmap(map, map_size);
    return rc;
}

uint64_t yb_length(const yb_ctx *c, const char *id, size_t id_len) {
    if (!c) return 0;
    const int64_t i = find_read(c, id, id_len);
    if (i < 0) return 0;
    if (c->indexed) return c->frozen || c->b_len ? c->len_host()[i] : 0;
    return c->length[(size_t)i];
}

uint32_t yb_n_reads(const yb_ctx *c) { return c ? c->total_reads() : 0; }

int yb_read_at(const yb_ctx *cc, uint32_t idx, const char **id, size_t *id_len) {
    yb_ctx *c = const_cast<yb_ctx *>(cc);
    if (!c || !id || !id_len) return YB_ERR_INVALID_ARGUMENT;
    if (idx >= c->total_reads()) return c->fail(YB_ERR_INVALID_ARGUMENT, "read index %u out of range", idx);
    if (c->indexed) {
        c->scratch_id = std::to_string(idx);
        *id = c->scratch_id.data();
        *id_len = c->scratch_id.size();
    } else {
        *id = c->ids.id(idx, id_len);
    }
    return YB_OK;
}

int64_t yb_read_index(const yb_ctx *c, const char *id, size_t id_len) { return c ? find_read(c, id, id_len) : -1; }

int yb_overlap(yb_ctx *c, const char *id, size_t id_len, const uint32_t **iv_pairs, uint32_t *n_intervals) {
    if (!c || !iv_pairs || !n_intervals) return YB_ERR_INVALID_ARGUMENT;
    *iv_pairs = nullptr;
    *n_intervals = 0;
    const int64_t i = find_read(c, id, id_len);
    if (i < 0) return YB_OK;  // fullmemory.rs:52-58: unknown read => empty
    if (c->from_report) return YB_OK;
    if (int rc = freeze(c)) return rc;
    const uint32_t *rp = c->rowptr_host();
    *iv_pairs = reinterpret_cast<const uint32_t *>(c->iv_host() + rp[i]);
    *n_intervals = rp[i + 1] - rp[i];
    return YB_OK;
}

// The detect kernels' view of the uploaded CSR and of the result buffers.
static int detect_args(yb_ctx *c, yb::DetectArgs *out) {
    yb::DetectArgs a{};
    a.iv = c->d_iv.p;
    a.rowptr = c->d_rowptr.p;
    a.len = c->d_len.p;
    a.n_reads = c->n_reads;
    a.n_iv = c->n_iv;
    a.max_k = c->max_k;
    a.validate = (!c->validated || c->n_literal) ? 1u : 0u;
    a.rows = c->rows;
    a.cls = c->d_cls.p;
    a.gap_ptr = c->d_gap_ptr.p;
    a.gaps = c->d_gaps.p;
    a.bitmap = c->ext_bitmap ? c->ext_bitmap : c->d_bitmap.p;
    a.counters = c->d_counters.p;
    a.side_stream = c->side_stream;
    a.ev_fork = c->ev_fork;
    a.ev_join = c->ev_join;
    a.scratch = c->d_scratch.p;
    a.scratch_bytes = c->d_scratch.cap;
    if (c->ext_bitmap && c->ext_bitmap_bytes < c->bitmap_bytes())
        return c->fail(YB_ERR_INVALID_ARGUMENT, "bound bitmap buffer too small (%zu < %zu bytes)", c->ext_bitmap_bytes, c->bitmap_bytes());
    a.n_peers = c->n_peers;
    if (c->n_peers) {
        if (c->peer_slot_bytes < c->bitmap_bytes())
            return c->fail(YB_ERR_INVALID_ARGUMENT, "peer slot too small (%zu < %zu bytes)", c->peer_slot_bytes, c->bitmap_bytes());
        a.rank = c->peer_rank;
        a.peer_parity_bytes = (size_t)c->n_peers * c->peer_slot_bytes;
        for (uint32_t p = 0; p < c->n_peers; ++p) {
            a.peer_slot[p] = c->peer_gather[p] + (size_t)c->peer_rank * c->peer_slot_bytes;
            a.peer_flag[p] = c->peer_flags[p];
        }
        a.bitmap = a.peer_slot[c->peer_rank];
    }
    *out = a;
    return YB_OK;
}

// ---- staged device API -------------------------------------------------------------------------------
// A lane's row pointers arrive as they are in the caller's CSR; the chunk's first one is subtracted in place.
__global__ void __launch_bounds__(256) rebase_kernel(uint32_t *rowptr, uint32_t n, uint32_t base) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i < n) rowptr[i] -= base;
}

int yb_upload(yb_ctx *c) {
    if (!c) return YB_ERR_INVALID_ARGUMENT;
    if (c->host_only) return c->fail(YB_ERR_CUDA, "host-only context: the detect path runs on a CUDA device only");
    if (c->from_report) return c->fail(YB_ERR_STATE, "context was loaded from a report; nothing to upload");
    YB_DEVICE(c);
    YB_CUDA(c, cudaSetDevice(c->device));
    if (int rc = freeze(c)) return rc;
    if (c->uploaded) return YB_OK;
    const size_t n = c->n_reads, m = c->n_iv;
    if (!c->d_rowptr.reserve(n + 1) || !c->d_len.reserve(n + 1) || !c->d_iv.reserve(m + 2))
        return c->fail(YB_ERR_NOMEM, "device allocation failed (%zu reads, %zu intervals)", n, m);
    if (int rc = ensure_result_buffers(c)) return rc;
    if (!c->d_rowstats.reserve(1) || !c->h_rowstats.reserve(1) || !c->h_valid.reserve(4))
        return c->fail(YB_ERR_NOMEM, "allocation failed");
    if (n) {
        YB_CUDA(c, cudaMemcpyAsync(c->d_rowptr.p, c->rowptr_host(), sizeof(uint32_t) * (n + 1), cudaMemcpyHostToDevice, c->stream));
        YB_CUDA(c, cudaMemcpyAsync(c->d_len.p, c->len_host(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c->stream));
    }
    if (c->is_lane) {
        // a chunk of a streamed run: the statistics come from the host copy (the host thread is ahead of the transfers
        // anyway), so nothing below waits for the device and the H2D engine goes from one chunk straight to the next
        if (n && m) YB_CUDA(c, cudaMemcpyAsync(c->d_iv.p, c->iv_host(), sizeof(uint2) * m, cudaMemcpyHostToDevice, c->stream));
        if (n && c->rowptr_base) {  // behind the copies: a kernel between them would hold the H2D engine until an SM is free
            rebase_kernel<<<(unsigned)((n + 256) / 256), 256, 0, c->stream>>>(c->d_rowptr.p, (uint32_t)n + 1u, c->rowptr_base);
            c->stats.kernel_launches += 1;
        }
        yb::host_row_stats(c->rowptr_host(), c->len_host(), c->n_reads, c->h_rowstats.p);  // (while the chunk crosses PCIe)
    } else {
        // size classes, big-row scratch needs and input sanity come from one small kernel over rowptr / len; its
        // 128-byte result crosses PCIe while the interval buffer is still on its way
        const int sl = yb::launch_row_stats(c->d_rowptr.p, c->d_len.p, c->n_reads, c->d_rowstats.p, c->stream);
        if (sl < 0) return c->cuda_fail(cudaGetLastError(), "row statistics kernel");
        c->stats.kernel_launches += (uint64_t)sl;
        YB_CUDA(c, cudaMemcpyAsync(c->h_rowstats.p, c->d_rowstats.p, sizeof(yb::DevRowStats), cudaMemcpyDeviceToHost, c->stream));
        cudaEvent_t ev;
        YB_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        YB_CUDA(c, cudaEventRecord(ev, c->stream));
        if (n && m) YB_CUDA(c, cudaMemcpyAsync(c->d_iv.p, c->iv_host(), sizeof(uint2) * m, cudaMemcpyHostToDevice, c->stream));
        const cudaError_t ee = cudaEventSynchronize(ev);
        cudaEventDestroy(ev);
        if (ee != cudaSuccess) return c->cuda_fail(ee, "cudaEventSynchronize");
    }
    {
        const yb::DevRowStats &ds = *c->h_rowstats.p;
        if (ds.bad_rowptr) return c->fail(YB_ERR_INVALID_ARGUMENT, "rowptr is not monotone (%u reads)", ds.bad_rowptr);
        if (ds.bad_len) return c->fail(YB_ERR_TOO_LARGE, "%u read(s) are longer than 2^31-1 bases", ds.bad_len);
        yb::RowStats rs;
        rs.n_big = ds.n_big;
        rs.n_scan = ds.n_scan;
        rs.max_len_scan = ds.max_len_scan;
        rs.max_k_sort = ds.max_k_sort;
        rs.big_pairs = ds.big_pairs;
        rs.huge_keys = ds.huge_keys;
        rs.n_wide = ds.n_wide;
        for (int q = 0; q < yb::kNumClasses; ++q) rs.class_count[q] = ds.class_count[q];
        c->rows = rs;
        c->max_k = ds.max_k;
    }
    // once per CSR (it cannot change afterwards): the size-class worklist of the register tier. The intervals themselves
    // are tested by the first detect step.
    // the device counts staged regions, offsets and worklist entries in 32 bits: refuse what could wrap them
    // (1.5 (n_iv + n_reads) + 2 n_reads staged pairs + one open chunk per resident warp, detect.cu:carve)
    {
        const uint64_t s = (uint64_t)c->n_iv + c->n_reads;
        if (s + s / 2 + 2ull * c->n_reads + 4096ull * 2048ull > 0xFFFFFFF0ull)
            return c->fail(YB_ERR_TOO_LARGE, "batch too large for one context (%u reads, %u intervals): split it (yb_set_chunk_intervals)",
                           c->n_reads, c->n_iv);
    }
    const size_t sb = yb::detect_scratch_bytes(c->n_reads, c->n_iv, c->rows);
    if (!c->d_scratch.reserve(sb)) return c->fail(YB_ERR_NOMEM, "device scratch allocation failed (%zu bytes)", sb);
    {
        yb::DetectArgs a{};
        a.iv = c->d_iv.p;
        a.rowptr = c->d_rowptr.p;
        a.len = c->d_len.p;
        a.n_reads = c->n_reads;
        a.n_iv = c->n_iv;
        a.max_k = c->max_k;
        a.rows = c->rows;
        a.counters = c->d_counters.p;
        a.scratch = c->d_scratch.p;
        a.scratch_bytes = c->d_scratch.cap;
        const int ul = yb::launch_upload_kernels(a, c->stream);
        if (ul < 0) return c->cuda_fail(cudaGetLastError(), "upload kernels (worklist)");
        c->stats.kernel_launches += (uint64_t)ul;
        YB_CUDA(c, cudaEventRecord(c->ev_upload, c->stream));
        c->validated = c->valid_pending = false;
        c->n_literal = c->n_malformed = 0;
    }
    c->stats.h2d_bytes += sizeof(uint32_t) * (2 * n + 1) + sizeof(uint2) * m;
    c->stats.d2h_bytes += sizeof(yb::DevRowStats);
    c->uploaded = true;
    c->computed = c->downloaded = false;
    return YB_OK;
}

int yb_compute_device(yb_ctx *c, uint64_t coverage, double not_coverage, void *stream) {
    if (!c) return YB_ERR_INVALID_ARGUMENT;
    if (c->host_only) return c->fail(YB_ERR_CUDA, "host-only context: the detect path runs on a CUDA device only");
    YB_DEVICE(c);
    YB_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    c->coverage = coverage;
    c->not_coverage = not_coverage;
    // a caller's stream: the step must run behind the upload (and a previous download) enqueued on the context's stream,
    // and yb_download behind the step. While the caller captures a CUDA graph the events stay out of it (the caller
    // synchronises around capture and replay).
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (st != c->stream) cudaStreamIsCapturing(st, &cap);
    if (st == c->stream) cap = cudaStreamCaptureStatusNone;
    const bool foreign = st != c->stream && cap == cudaStreamCaptureStatusNone;
    int launches;
    if (c->from_report) {
        if (foreign) {
            YB_CUDA(c, cudaEventRecord(c->ev_upload, c->stream));
            YB_CUDA(c, cudaStreamWaitEvent(st, c->ev_upload, 0));
        }
        launches = yb::launch_classify(c->d_len.p, c->d_gap_ptr.p, c->d_gaps.p, c->n_reads, not_coverage, c->d_cls.p,
                                       c->ext_bitmap ? c->ext_bitmap : c->d_bitmap.p, c->d_counters.p, st);
    } else {
        if (!c->uploaded) return c->fail(YB_ERR_STATE, "yb_compute_device before yb_upload");
        if (c->valid_pending && cap == cudaStreamCaptureStatusNone) {  // the verdict of the validating step before this one
            YB_CUDA(c, cudaEventSynchronize(c->ev_valid));
            c->n_literal = c->h_valid.p[0];
            c->n_malformed = c->h_valid.p[1];
            c->validated = true;
            c->valid_pending = false;
        }
        if (foreign) {
            YB_CUDA(c, cudaEventRecord(c->ev_upload, c->stream));  // (also covers a download of the previous step)
            YB_CUDA(c, cudaStreamWaitEvent(st, c->ev_upload, 0));
        }
        yb::DetectArgs a{};
        if (int rc = detect_args(c, &a)) return rc;
        launches = yb::launch_detect(a, coverage > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)coverage, not_coverage, st);
        if (launches >= 0 && a.validate && cap == cudaStreamCaptureStatusNone && c->n_reads) {  // its two counts follow the step
            YB_CUDA(c, cudaMemcpyAsync(c->h_valid.p, c->d_counters.p + yb::kCntLiteralLast, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            YB_CUDA(c, cudaMemcpyAsync(c->h_valid.p + 1, c->d_counters.p + yb::kCntMalformedLast, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            YB_CUDA(c, cudaEventRecord(c->ev_valid, st));
            c->valid_pending = true;
        }
    }
    if (launches < 0) return c->cuda_fail(cudaGetLastError(), "kernel launch");
    if (foreign) {
        YB_CUDA(c, cudaEventRecord(c->ev_compute, st));
        c->compute_foreign = true;
    }
    c->stats.kernel_launches += (uint64_t)launches;
    c->computed = true;
    c->downloaded = false;
    return YB_OK;
}

int yb_time_upload_kernels(yb_ctx *c, float *ms_out) {
    if (!c || !ms_out) return YB_ERR_INVALID_ARGUMENT;
    if (!c->uploaded || c->from_report) return c->fail(YB_ERR_STATE, "yb_time_upload_kernels needs an uploaded CSR");
    YB_DEVICE(c);
    YB_CUDA(c, cudaSetDevice(c->device));
    cudaEvent_t e0, e1;
    YB_CUDA(c, cudaEventCreate(&e0));
    YB_CUDA(c, cudaEventCreate(&e1));
    yb::DetectArgs a{};
    int rc = detect_args(c, &a);
    if (rc == YB_OK) {
        cudaEventRecord(e0, c->stream);
        const int l0 = yb::launch_row_stats(c->d_rowptr.p, c->d_len.p, c->n_reads, c->d_rowstats.p, c->stream);
        const int l1 = yb::launch_upload_kernels(a, c->stream);
        cudaEventRecord(e1, c->stream);
        if (l0 < 0 || l1 < 0 || cudaEventSynchronize(e1) != cudaSuccess || cudaEventElapsedTime(ms_out, e0, e1) != cudaSuccess)
            rc = c->cuda_fail(cudaGetLastError(), "upload kernels");
        else
            c->stats.kernel_launches += (uint64_t)(l0 + l1);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    c->computed = c->downloaded = false;  // the counters were reset with the worklist
    return rc;
}

int yb_time_one_shot(yb_ctx *c, uint64_t coverage, double not_coverage, float *ms_upload_kernels, float *ms_first_step) {
    if (!c || !ms_upload_kernels || !ms_first_step) return YB_ERR_INVALID_ARGUMENT;
    if (!c->uploaded || c->from_report) return c->fail(YB_ERR_STATE, "yb_time_one_shot needs an uploaded CSR");
    YB_DEVICE(c);
    YB_CUDA(c, cudaSetDevice(c->device));
    cudaEvent_t e0, e1, e2;
    YB_CUDA(c, cudaEventCreate(&e0));
    YB_CUDA(c, cudaEventCreate(&e1));
    YB_CUDA(c, cudaEventCreate(&e2));
    c->validated = c->valid_pending = false;  // as after yb_upload: the step below tests every interval
    c->n_literal = c->n_malformed = 0;
    yb::DetectArgs a{};
    int rc = detect_args(c, &a);
    if (rc == YB_OK) {
        cudaEventRecord(e0, c->stream);
        const int l0 = yb::launch_row_stats(c->d_rowptr.p, c->d_len.p, c->n_reads, c->d_rowstats.p, c->stream);
        const int l1 = yb::launch_upload_kernels(a, c->stream);
        cudaEventRecord(e1, c->stream);
        const int l2 = yb::launch_detect(a, coverage > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)coverage, not_coverage, c->stream);
        cudaEventRecord(e2, c->stream);
        if (l0 < 0 || l1 < 0 || l2 < 0 || cudaEventSynchronize(e2) != cudaSuccess || cudaEventElapsedTime(ms_upload_kernels, e0, e1) != cudaSuccess ||
            cudaEventElapsedTime(ms_first_step, e1, e2) != cudaSuccess)
            rc = c->cuda_fail(cudaGetLastError(), "one-shot timing");
        else
            c->stats.kernel_launches += (uint64_t)(l0 + l1 + l2);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    c->coverage = coverage;
    c->not_coverage = not_coverage;
    c->computed = rc == YB_OK;
    c->downloaded = false;
    return rc;
}

int yb_peer_wait(yb_ctx *c, void *stream) {
    if (!c) return YB_ERR_INVALID_ARGUMENT;
    if (!c->n_peers) return YB_OK;
    YB_DEVICE(c);
    YB_CUDA(c, cudaSetDevice(c->device));
    yb::DetectArgs a{};
    if (int rc = detect_args(c, &a)) return rc;
    const int l = yb::launch_peer_wait(a, stream ? static_cast<cudaStream_t>(stream) : c->stream);
    if (l < 0) return c->cuda_fail(cudaGetLastError(), "peer wait kernel");
    c->stats.kernel_launches += (uint64_t)l;
    return YB_OK;
}

int yb_synchronize(yb_ctx *c) {
    if (!c) return YB_ERR_INVALID_ARGUMENT;
    if (c->host_only) return YB_OK;
    YB_DEVICE(c);
    YB_CUDA(c, cudaSetDevice(c->device));
    YB_CUDA(c, cudaStreamSynchronize(c->stream));
    return YB_OK;
}

// The results of c's last step go to dst's host arrays: dst == c (row0 = gap0 = 0) for a whole CSR; for a lane of a
// streamed run dst is the parent, row0 the chunk's first read and gap0 the bad regions of the chunks before it.
static int download_impl(yb_ctx *c, yb_ctx *dst, size_t row0, size_t gap0) {
    if (!c->computed) return c->fail(YB_ERR_STATE, "yb_download before yb_compute_device");
    if (c->downloaded) return YB_OK;
    YB_DEVICE(c);
    YB_CUDA(c, cudaSetDevice(c->device));
    if (c->compute_foreign) {  // the step ran on a caller's stream: the copies below go behind it
        YB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_compute, 0));
        c->compute_foreign = false;
    }
    const size_t n = c->n_reads;
    const bool peers = c->n_peers && !c->from_report;
    if ((dst == c && (!c->h_cls.reserve(n + 1) || !c->h_gap_ptr.reserve(n + 1) || !c->h_bitmap.reserve(2 * (c->bitmap_bytes() + 4)))) ||
        !c->h_counters.reserve(yb::kCounterWords) || !c->h_peer_step.reserve(1))
        return c->fail(YB_ERR_NOMEM, "pinned host allocation failed");
    uint32_t *const gap_ptr_dst = dst->h_gap_ptr.p + row0;
    uint8_t *const bitmap_dst = dst->h_bitmap.p + row0 / 4;  // (a chunk starts at a multiple of 1024 reads)
    YB_CUDA(c, cudaMemcpyAsync(c->h_counters.p, c->d_counters.p, sizeof(uint32_t) * yb::kCounterWords, cudaMemcpyDeviceToHost, c->stream));
    YB_CUDA(c, cudaMemcpyAsync(gap_ptr_dst, c->d_gap_ptr.p, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost, c->stream));
    const size_t bmb = c->bitmap_bytes();
    if (n) {
        YB_CUDA(c, cudaMemcpyAsync(dst->h_cls.p + row0, c->d_cls.p, n, cudaMemcpyDeviceToHost, c->stream));
        if (peers) {  // this rank's slot of the last step: even steps use the first half of the gather buffer, odd ones the second
            const uint8_t *own = c->peer_gather[c->peer_rank] + (size_t)c->peer_rank * c->peer_slot_bytes;
            YB_CUDA(c, cudaMemcpyAsync(c->h_peer_step.p, c->peer_flags[c->peer_rank] + 31, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            YB_CUDA(c, cudaMemcpyAsync(c->h_bitmap.p, own, bmb, cudaMemcpyDeviceToHost, c->stream));
            YB_CUDA(c, cudaMemcpyAsync(c->h_bitmap.p + bmb + 4, own + (size_t)c->n_peers * c->peer_slot_bytes, bmb, cudaMemcpyDeviceToHost, c->stream));
        } else {
            YB_CUDA(c, cudaMemcpyAsync(bitmap_dst, c->ext_bitmap ? c->ext_bitmap : c->d_bitmap.p, bmb, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    YB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (peers && n && ((*c->h_peer_step.p - 1u) & 1u)) memcpy(c->h_bitmap.p, c->h_bitmap.p + bmb + 4, bmb);
    // the counter set of the last step: the step number was incremented when the step closed
    const uint32_t *hc = c->h_counters.p;
    const uint32_t *set = c->from_report ? hc : hc + ((hc[yb::kCntEpoch] - 1u) & 1u) * yb::kNumCounters;
    if (!c->from_report && n && hc[yb::kCntEpoch] == 0)
        return c->fail(YB_ERR_STATE, "internal error: the detect step did not complete");
    if (set[yb::kCntPeerTimeout] || hc[yb::kCntPeerTimeoutWait])
        return c->fail(YB_ERR_STATE, "peer all-gather: a rank never signalled its step (%u / %u)", set[yb::kCntPeerTimeout],
                       hc[yb::kCntPeerTimeoutWait]);
    if (hc[yb::kCntOrderTimeout]) return c->fail(YB_ERR_STATE, "internal error: the ordering kernel did not complete");
    if (set[yb::kCntStageOverflow])
        return c->fail(YB_ERR_STATE, "internal error: bad-region staging buffer overflow (%u reads)", set[yb::kCntStageOverflow]);
    c->n_gaps = gap_ptr_dst[n];
    size_t d2h = sizeof(uint32_t) * (n + 1 + yb::kCounterWords) + n + bmb;
    if (!c->from_report) {
        if (!(dst == c ? c->h_gaps.reserve((size_t)c->n_gaps + 1) : dst->h_gaps.grow_keep(gap0 + c->n_gaps + 1, gap0)))
            return c->fail(YB_ERR_NOMEM, "pinned host allocation failed");
        if (c->n_gaps) {
            YB_CUDA(c, cudaMemcpyAsync(dst->h_gaps.p + gap0, c->d_gaps.p, sizeof(uint2) * c->n_gaps, cudaMemcpyDeviceToHost, c->stream));
            if (gap0)  // (while the regions cross PCIe)
                for (size_t i = 0; i <= n; ++i) gap_ptr_dst[i] += (uint32_t)gap0;
            YB_CUDA(c, cudaStreamSynchronize(c->stream));
        } else if (gap0) {
            for (size_t i = 0; i <= n; ++i) gap_ptr_dst[i] += (uint32_t)gap0;
        }
        d2h += sizeof(uint2) * c->n_gaps;
    }
    c->stats.d2h_bytes += d2h;
    c->stats.n_reads = n;
    c->stats.n_intervals = c->n_iv;
    c->stats.n_gaps = c->n_gaps;
    uint64_t hist[3] = {0, 0, 0};
    if (n) {
        for (int t = 0; t < 3; ++t) hist[t] = set[yb::kCntNotBad + t];  // classify_kernel (FromReport path)
        for (uint32_t sl = 0; sl < yb::kHistSlots; ++sl)  // the detect step stripes its histogram over kHistSlots copies
            for (int t = 0; t < 3; ++t) hist[t] += set[yb::kCntHist + 3 * sl + t];
    }
    c->stats.n_not_bad = hist[0];
    c->stats.n_chimeric = hist[1];
    c->stats.n_not_covered = hist[2];
    c->stats.max_intervals_per_read = c->max_k;
    if (!c->from_report && n) {  // what the last validating step over this CSR found (copied with the counters)
        c->n_literal = hc[yb::kCntLiteralLast];
        c->n_malformed = hc[yb::kCntMalformedLast];
        c->validated = true;
        c->valid_pending = false;
    }
    c->stats.n_malformed_intervals = c->from_report ? 0 : c->n_malformed;
    c->stats.n_literal_reads = c->from_report ? 0 : c->n_literal;
    c->downloaded = true;
    return YB_OK;
}

int yb_download(yb_ctx *c) {
    if (!c) return YB_ERR_INVALID_ARGUMENT;
    return download_impl(c, c, 0, 0);
}

int yb_set_chunk_intervals(yb_ctx *c, uint32_t n_intervals) {
    if (!c) return YB_ERR_INVALID_ARGUMENT;
    c->chunk_intervals = n_intervals;
    return YB_OK;
}

// Streamed yb_compute_all_bad_part: row chunks of about chunk_intervals intervals alternate between two lanes. The
// host thread only ever blocks for the chunk two behind the one it just enqueued, so the H2D engine always has the
// next chunk queued (PCIe is what bounds the whole call: 8 bytes per interval in, about 0.5 out), the kernels of a
// chunk run under the next chunk's transfer, and the results land directly in this context's host arrays.
static int compute_streamed(yb_ctx *c, uint64_t coverage, double not_coverage) {
    const uint32_t n = c->n_reads;
    const uint32_t *rp = c->rowptr_host();
    if (!c->h_cls.reserve((size_t)n + 1) || !c->h_gap_ptr.reserve((size_t)n + 1) || !c->h_bitmap.reserve(2 * (c->bitmap_bytes() + 4)))
        return c->fail(YB_ERR_NOMEM, "pinned host allocation failed");
    while (c->lanes.size() < 2) {
        yb_opts o{};
        o.device = c->device;
        o.read_buffer_size = c->read_buffer_size;
        o.flags = c->flags & ~(uint32_t)(YB_FLAG_HOST_ONLY | YB_FLAG_LAZY_DEVICE);
        yb_ctx *l = yb_create(&o);
        if (!l) return c->fail(YB_ERR_CUDA, "streamed batches: %s", yb_create_error());
        l->is_lane = true;
        c->lanes.push_back(l);
    }
    struct InFlight {
        bool busy = false;
        uint32_t row0 = 0;
    } fl[2];
    size_t gap0 = 0;
    yb_stats tot{};
    uint32_t n_chunks = 0;
    auto finish = [&](int li) -> int {
        yb_ctx *l = c->lanes[li];
        if (int rc = download_impl(l, c, fl[li].row0, gap0)) return c->fail(rc, "%s", l->error.c_str());
        gap0 += l->n_gaps;
        if (gap0 > 0xFFFFFFF0ull) return c->fail(YB_ERR_TOO_LARGE, "more than 2^32-16 bad regions");
        tot.n_not_bad += l->stats.n_not_bad;
        tot.n_chimeric += l->stats.n_chimeric;
        tot.n_not_covered += l->stats.n_not_covered;
        tot.n_malformed_intervals += l->stats.n_malformed_intervals;
        tot.n_literal_reads += l->stats.n_literal_reads;
        tot.max_intervals_per_read = std::max(tot.max_intervals_per_read, l->stats.max_intervals_per_read);
        tot.h2d_bytes += l->stats.h2d_bytes;
        tot.d2h_bytes += l->stats.d2h_bytes;
        tot.kernel_launches += l->stats.kernel_launches;
        fl[li].busy = false;
        return YB_OK;
    };
    uint32_t r0 = 0;
    int li = 0;
    while (r0 < n) {
        // the chunk ends at the first multiple of 1024 reads that holds chunk_intervals intervals (or at the end)
        const uint32_t target = (uint32_t)std::min<uint64_t>((uint64_t)rp[r0] + c->chunk_intervals, 0xFFFFFFFFull);
        const uint32_t *stop = std::lower_bound(rp + r0 + 1, rp + n, target);
        uint32_t r1 = (uint32_t)(stop - rp);
        r1 = (uint32_t)std::min<uint64_t>(n, ((uint64_t)r1 + 1023u) & ~(uint64_t)1023u);
        if (fl[li].busy)
            if (int rc = finish(li)) return rc;
        yb_ctx *l = c->lanes[li];
        if (int rc = yb_reset(l)) return rc;
        l->stats = yb_stats{};
        l->rowptr_base = rp[r0];
        if (int rc = yb_bind_csr(l, rp + r0, reinterpret_cast<const uint32_t *>(c->iv_host() + rp[r0]), c->len_host() + r0, r1 - r0))
            return c->fail(rc, "%s", l->error.c_str());
        if (int rc = yb_upload(l)) return c->fail(rc, "%s", l->error.c_str());
        if (int rc = yb_compute_device(l, coverage, not_coverage, nullptr)) return c->fail(rc, "%s", l->error.c_str());
        fl[li].busy = true;
        fl[li].row0 = r0;
        ++n_chunks;
        r0 = r1;
        li ^= 1;
    }
    for (int k = 0; k < 2; ++k, li ^= 1)  // the older chunk first: its regions come first
        if (fl[li].busy)
            if (int rc = finish(li)) return rc;
    c->n_gaps = (uint32_t)gap0;
    c->h_gap_ptr.p[n] = (uint32_t)gap0;
    c->max_k = tot.max_intervals_per_read;
    c->stats.n_reads = n;
    c->stats.n_intervals = c->n_iv;
    c->stats.n_gaps = gap0;
    c->stats.n_not_bad = tot.n_not_bad;
    c->stats.n_chimeric = tot.n_chimeric;
    c->stats.n_not_covered = tot.n_not_covered;
    c->stats.n_malformed_intervals = tot.n_malformed_intervals;
    c->stats.n_literal_reads = tot.n_literal_reads;
    c->stats.max_intervals_per_read = tot.max_intervals_per_read;
    c->stats.h2d_bytes += tot.h2d_bytes;
    c->stats.d2h_bytes += tot.d2h_bytes;
    c->stats.kernel_launches += tot.kernel_launches;
    c->n_chunks_last = n_chunks;
    c->coverage = coverage;
    c->not_coverage = not_coverage;
    c->computed = c->downloaded = true;
    return YB_OK;
}


int yb_compute_all_bad_part(yb_ctx *c, uint64_t coverage, double not_coverage) {
    if (!c) return YB_ERR_INVALID_ARGUMENT;
    if (!c->from_report && c->chunk_intervals && !c->host_only && !c->n_peers && !c->ext_bitmap && !c->uploaded) {
        YB_DEVICE(c);
        YB_CUDA(c, cudaSetDevice(c->device));
        if (int rc = freeze(c)) return rc;
        if (c->n_iv > c->chunk_intervals) return compute_streamed(c, coverage, not_coverage);
    }
    if (!c->from_report)
        if (int rc = yb_upload(c)) return rc;
    if (int rc = yb_compute_device(c, coverage, not_coverage, nullptr)) return rc;
    return yb_download(c);
}

int yb_bind_device_bitmap(yb_ctx *c, void *device_ptr, size_t n_bytes) {
    if (!c) return YB_ERR_INVALID_ARGUMENT;
    c->ext_bitmap = static_cast<uint8_t *>(device_ptr);
    c->ext_bitmap_bytes = device_ptr ? n_bytes : 0;
    return YB_OK;
}

void *yb_peer_alloc(yb_ctx *c, size_t n_bytes, void *handle_out) {
    if (!c || c->host_only || !handle_out || n_bytes == 0) return nullptr;
    static_assert(sizeof(cudaIpcMemHandle_t) == YB_IPC_HANDLE_BYTES, "handle size");
    if (open_device(c) != YB_OK || cudaSetDevice(c->device) != cudaSuccess) return nullptr;
    void *p = nullptr;
    cudaIpcMemHandle_t h;
    if (cudaMalloc(&p, n_bytes) != cudaSuccess || cudaMemset(p, 0, n_bytes) != cudaSuccess || cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
        c->cuda_fail(cudaGetLastError(), "yb_peer_alloc");
        if (p) cudaFree(p);
        return nullptr;
    }
    memcpy(handle_out, &h, sizeof h);
    return p;
}

void *yb_peer_open(yb_ctx *c, const void *handle) {
    if (!c || c->host_only || !handle) return nullptr;
    if (open_device(c) != YB_OK || cudaSetDevice(c->device) != cudaSuccess) return nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        c->cuda_fail(cudaGetLastError(), "cudaIpcOpenMemHandle");
        return nullptr;
    }
    return p;
}

int yb_peer_close(yb_ctx *c, void *mapped) {
    if (!c || !mapped) return YB_ERR_INVALID_ARGUMENT;
    YB_DEVICE(c);
    YB_CUDA(c, cudaSetDevice(c->device));
    YB_CUDA(c, cudaStreamSynchronize(c->stream));
    YB_CUDA(c, cudaIpcCloseMemHandle(mapped));
    return YB_OK;
}

int yb_peer_free(yb_ctx *c, void *allocated) {
    if (!c || !allocated) return YB_ERR_INVALID_ARGUMENT;
    YB_DEVICE(c);
    YB_CUDA(c, cudaSetDevice(c->device));
    YB_CUDA(c, cudaStreamSynchronize(c->stream));
    YB_CUDA(c, cudaFree(allocated));
    return YB_OK;
}

int yb_bind_peers(yb_ctx *c, void *const *gather_bufs, void *const *flag_bufs, uint32_t n_ranks, uint32_t rank, size_t slot_bytes) {
    if (!c) return YB_ERR_INVALID_ARGUMENT;
    if (n_ranks == 0) {
        c->n_peers = 0;
        return YB_OK;
    }
    if (!gather_bufs || !flag_bufs || n_ranks > YB_MAX_PEERS || rank >= n_ranks || (slot_bytes & 3u))
        return c->fail(YB_ERR_INVALID_ARGUMENT, "yb_bind_peers: at most %d ranks, rank < n_ranks, slot_bytes a multiple of 4", YB_MAX_PEERS);
    for (uint32_t p = 0; p < n_ranks; ++p) {
        if (!gather_bufs[p] || !flag_bufs[p]) return c->fail(YB_ERR_INVALID_ARGUMENT, "yb_bind_peers: null buffer for rank %u", p);
        c->peer_gather[p] = static_cast<uint8_t *>(gather_bufs[p]);
        c->peer_flags[p] = static_cast<uint32_t *>(flag_bufs[p]);
    }
    c->n_peers = n_ranks;
    c->peer_rank = rank;
    c->peer_slot_bytes = slot_bytes;
    return YB_OK;
}

void *yb_device_class_bitmap(yb_ctx *c, size_t *n_bytes) {
    if (!c) return nullptr;
    if (n_bytes) *n_bytes = c->bitmap_bytes();
    return c->ext_bitmap ? c->ext_bitmap : c->d_bitmap.p;
}
void *yb_device_classes(yb_ctx *c, size_t *n) {
    if (!c) return nullptr;
    if (n) *n = c->n_reads;
    return c->d_cls.p;
}
void *yb_device_gap_ptr(yb_ctx *c, size_t *n) {
    if (!c) return nullptr;
    if (n) *n = (size_t)c->n_reads + 1;
    return c->d_gap_ptr.p;
}
void *yb_device_gaps(yb_ctx *c, size_t *cap) {
    if (!c) return nullptr;
    if (cap) *cap = c->d_gaps.cap;
    return c->d_gaps.p;
}
void *yb_stream(yb_ctx *c) { return c && open_device(c) == YB_OK ? c->stream : nullptr; }

int yb_get_stats(yb_ctx *c, yb_stats *out) {
    if (!c || !out) return YB_ERR_INVALID_ARGUMENT;
    *out = c->stats;
    return YB_OK;
}

// ---- consumer side -----------------------------------------------------------------------------------
int yb_get_bad_part_at(yb_ctx *c, uint32_t idx, const uint32_t **gap_pairs, uint32_t *n_gaps, uint64_t *length,
                       uint8_t *cls) {
    if (!c) return YB_ERR_INVALID_ARGUMENT;
    if (!c->downloaded) return c->fail(YB_ERR_STATE, "results queried before yb_compute_all_bad_part / yb_download");
    if (idx >= c->n_reads) return c->fail(YB_ERR_INVALID_ARGUMENT, "read index %u out of range", idx);
    const uint32_t g0 = c->h_gap_ptr.p[idx], g1 = c->h_gap_ptr.p[idx + 1];
    if (gap_pairs) *gap_pairs = reinterpret_cast<const uint32_t *>(c->h_gaps.p + g0);
    if (n_gaps) *n_gaps = g1 - g0;
    if (length) *length = c->from_report ? c->length[idx] : c->len_host()[idx];
    if (cls) *cls = c->h_cls.p[idx];
    return YB_OK;
}

int yb_get_bad_part(yb_ctx *c, const char *id, size_t id_len, const uint32_t **gap_pairs, uint32_t *n_gaps,
                    uint64_t *length, uint8_t *cls) {
    if (!c) return YB_ERR_INVALID_ARGUMENT;
    if (!c->downloaded) return c->fail(YB_ERR_STATE, "results queried before yb_compute_all_bad_part / yb_download");
    const int64_t i = find_read(c, id, id_len);
    if (i < 0) {  // stack.rs:164-169: unknown id => (vec![], 0); 0/0 = NaN is not > n => NotBad
        if (gap_pairs) *gap_pairs = nullptr;
        if (n_gaps) *n_gaps = 0;
        if (length) *length = 0;
        if (cls) *cls = YB_NOT_BAD;
        return YB_OK;
    }
    return yb_get_bad_part_at(c, (uint32_t)i, gap_pairs, n_gaps, length, cls);
}

int yb_edit(yb_ctx *c, int editor, const char *input_path, const char *output_path) {
    if (!c || !input_path || !output_path) return YB_ERR_INVALID_ARGUMENT;
    if (!c->downloaded) return c->fail(YB_ERR_STATE, "editors need the results: call yb_compute_all_bad_part first");
    std::string msg;
    const int rc = yb::run_editor(c, editor, input_path, output_path, c->read_buffer_size, &msg);
    if (rc != YB_OK && !msg.empty()) return c->fail(rc, "%s", msg.c_str());
    return rc;
}

const uint8_t *yb_classes(yb_ctx *c, size_t *n) {
    if (!c || !c->downloaded) return nullptr;
    if (n) *n = c->n_reads;
    return c->h_cls.p;
}
const uint8_t *yb_class_bitmap(yb_ctx *c, size_t *n_bytes) {
    if (!c || !c->downloaded) return nullptr;
    if (n_bytes) *n_bytes = c->bitmap_bytes();
    return c->h_bitmap.p;
}
const uint32_t *yb_gap_ptr(yb_ctx *c, size_t *n) {
    if (!c || !c->downloaded) return nullptr;
    if (n) *n = (size_t)c->n_reads + 1;
    return c->h_gap_ptr.p;
}
const uint32_t *yb_gaps(yb_ctx *c, size_t *n_pairs) {
    if (!c || !c->downloaded) return nullptr;
    if (n_pairs) *n_pairs = c->n_gaps;
    return reinterpret_cast<const uint32_t *>(c->h_gaps.p);
}

// editor/mod.rs:72-79,102-107: "{type}\t{id}\t{len}\t{len,beg,end;...}"
static size_t format_line(yb_ctx *c, uint32_t idx, std::vector<char> *buf) {
    const char *id;
    size_t idl;
    std::string tmp;
    if (c->indexed) {
        tmp = std::to_string(idx);
        id = tmp.data();
        idl = tmp.size();
    } else {
        id = c->ids.id(idx, &idl);
    }
    const uint32_t g0 = c->h_gap_ptr.p[idx], g1 = c->h_gap_ptr.p[idx + 1];
    const size_t need = 16 + idl + 24 + (size_t)(g1 - g0) * 34 + 2;
    const size_t at = buf->size();
    buf->resize(at + need);
    char *p = buf->data() + at;
    const char *t = kTypeNames[c->h_cls.p[idx] < 3 ? c->h_cls.p[idx] : 0];
    const size_t tl = strlen(t);
    memcpy(p, t, tl);
    p += tl;
    *p++ = '\t';
    memcpy(p, id, idl);
    p += idl;
    *p++ = '\t';
    p = put_u64(p, c->from_report ? c->length[idx] : c->len_host()[idx]);
    *p++ = '\t';
    for (uint32_t g = g0; g < g1; ++g) {
        const uint2 v = c->h_gaps.p[g];
        if (g != g0) *p++ = ';';
        p = put_u64(p, (uint32_t)(v.y - v.x));
        *p++ = ',';
        p = put_u64(p, v.x);
        *p++ = ',';
        p = put_u64(p, v.y);
    }
    const size_t used = (size_t)(p - (buf->data() + at));
    buf->resize(at + used);
    return used;
}

int64_t yb_format_report_line(yb_ctx *c, uint32_t idx, char *out, size_t cap) {
    if (!c || !out) return YB_ERR_INVALID_ARGUMENT;
    if (!c->downloaded) return c->fail(YB_ERR_STATE, "results queried before compute");
    if (idx >= c->n_reads) return c->fail(YB_ERR_INVALID_ARGUMENT, "read index %u out of range", idx);
    std::vector<char> buf;
    const size_t n = format_line(c, idx, &buf);
    if (n + 1 > cap) return c->fail(YB_ERR_INVALID_ARGUMENT, "line needs %zu bytes", n + 1);
    memcpy(out, buf.data(), n);
    out[n] = 0;
    return (int64_t)n;
}

int yb_write_report(yb_ctx *c, const char *path) {  // main.rs:62-84
    if (!c || !path) return YB_ERR_INVALID_ARGUMENT;
    if (!c->downloaded) return c->fail(YB_ERR_STATE, "yb_write_report before yb_compute_all_bad_part");
    FILE *fh = fopen(path, "wb");
    if (!fh) return c->fail(YB_ERR_CANT_WRITE_FILE, "Can't create file %s: %s", path, strerror(errno));
    std::vector<char> buf;
    buf.reserve(1 << 22);
    bool ok = true;
    for (uint32_t r = 0; r < c->n_reads && ok; ++r) {
        format_line(c, r, &buf);
        buf.push_back('\n');
        if (buf.size() > (1u << 22) - 4096) {
            ok = fwrite(buf.data(), 1, buf.size(), fh) == buf.size();
            buf.clear();
        }
    }
    if (ok && !buf.empty()) ok = fwrite(buf.data(), 1, buf.size(), fh) == buf.size();
    if (fclose(fh) != 0) ok = false;
    if (!ok) return c->fail(YB_ERR_WRITING, "Error during writing of file in yacrd format (%s)", path);
    return YB_OK;
}

// ---- FromReport (stack.rs:176-257) -------------------------------------------------------------------
static int init_report_impl(yb_ctx *c, const char *text, size_t n_bytes);
int yb_init_report_buffer(yb_ctx *c, const char *text, size_t n_bytes) {
    if (!c || (!text && n_bytes)) return YB_ERR_INVALID_ARGUMENT;
    const int rc = init_report_impl(c, text, n_bytes);
    if (rc != YB_OK && rc != YB_ERR_STATE) {  // a failed load leaves an empty context, not a half-filled one
        const std::string why = c->error;
        yb_reset(c);
        c->error = why;
    }
    return rc;
}
static int init_report_impl(yb_ctx *c, const char *text, size_t n_bytes) {
    if (c->host_only) return c->fail(YB_ERR_CUDA, "host-only context: reports are classified on a CUDA device only");
    if (c->total_reads() || c->from_report) return c->fail(YB_ERR_STATE, "yb_init_report needs an empty context");
    YB_DEVICE(c);
    YB_CUDA(c, cudaSetDevice(c->device));
    std::vector<uint32_t> gp(1, 0);
    std::vector<uint2> gaps;
    std::vector<uint32_t> slot_gp_begin;  // per read index: where its gaps start (last occurrence wins)
    struct Row { uint64_t len; uint32_t g0, g1; };
    std::vector<Row> rows;
    const char *p = text, *const end = text + n_bytes;
    uint64_t line = 0;
    auto corrupt = [&](uint64_t ln) { return c->fail(YB_ERR_CORRUPT_REPORT, "Your yacrd file is corrupt at line %llu", (unsigned long long)ln); };
    auto parse_u = [](const char *s, const char *e, uint64_t max, uint64_t *out) {
        if (s < e && *s == '+') ++s;
        if (s == e) return false;
        uint64_t v = 0;
        for (; s < e; ++s) {
            const unsigned d = (unsigned)(*s - '0');
            if (d > 9 || v > (max - d) / 10) return false;
            v = v * 10 + d;
        }
        *out = v;
        return true;
    };
    while (p < end) {
        const char *q = p;
        while (q < end && *q != '\n' && *q != '\r') ++q;
        const char *le = q;
        if (q < end) {
            if (*q == '\r' && q + 1 < end && q[1] == '\n') ++q;
            ++q;
        }
        const char *ls = p;
        p = q;
        if (le == ls) continue;
        // split 4 tab-separated fields: type, id, length, bad string (stack.rs:203-211)
        const char *f[5];
        int nf = 0;
        f[nf++] = ls;
        for (const char *s = ls; s < le && nf < 4; ++s)
            if (*s == '\t') f[nf++] = s + 1;
        if (nf < 4) return corrupt(line);
        const char *id = f[1], *id_e = f[2] - 1, *len_s = f[2], *len_e = f[3] - 1, *bad = f[3];
        const char *bad_e = bad;
        while (bad_e < le && *bad_e != '\t') ++bad_e;
        uint64_t len;
        if (!parse_u(len_s, len_e, UINT64_MAX, &len)) return corrupt(line);
        if (len > 0xFFFFFFFFull)  // the classifier holds lengths in 32 bits; the reference divides by the full usize
            return c->fail(YB_ERR_TOO_LARGE, "line %llu: read length %llu does not fit 32 bits", (unsigned long long)line, (unsigned long long)len);
        bool is_new;
        const uint32_t idx = c->ids.intern(id, (size_t)(id_e - id), &is_new);
        const uint32_t g0 = (uint32_t)gaps.size();
        if (bad != bad_e) {  // parse_bad_string, stack.rs:217-241: "len,begin,end" joined by ';'
            const char *s = bad;
            while (s <= bad_e) {
                const char *se = s;
                while (se < bad_e && *se != ';') ++se;
                const char *c1 = s;
                while (c1 < se && *c1 != ',') ++c1;
                if (c1 == se) return corrupt(line);
                const char *c2 = c1 + 1;
                while (c2 < se && *c2 != ',') ++c2;
                if (c2 == se) return corrupt(line);
                const char *c3 = c2 + 1;
                while (c3 < se && *c3 != ',') ++c3;
                uint64_t b, e;
                if (!parse_u(c1 + 1, c2, UINT32_MAX, &b) || !parse_u(c2 + 1, c3, UINT32_MAX, &e)) return corrupt(line);
                gaps.push_back(make_uint2((uint32_t)b, (uint32_t)e));
                s = se + 1;
                if (se == bad_e) break;
            }
        }
        const Row row{len, g0, (uint32_t)gaps.size()};
        if (is_new)
            rows.push_back(row);
        else
            rows[idx] = row;  // HashMap::insert: a repeated id keeps the last line (stack.rs:213)
        ++line;
    }
    // compact rows into a gap CSR in read order
    const size_t n = rows.size();
    c->length.resize(n);
    if (!c->h_gap_ptr.reserve(n + 1) || !c->h_len.reserve(n + 1)) return c->fail(YB_ERR_NOMEM, "pinned host allocation failed");
    size_t tot = 0;
    for (size_t r = 0; r < n; ++r) tot += rows[r].g1 - rows[r].g0;
    if (!c->h_gaps.reserve(tot + 1)) return c->fail(YB_ERR_NOMEM, "pinned host allocation failed");
    size_t at = 0;
    for (size_t r = 0; r < n; ++r) {
        c->h_gap_ptr.p[r] = (uint32_t)at;
        for (uint32_t g = rows[r].g0; g < rows[r].g1; ++g) c->h_gaps.p[at++] = gaps[g];
        c->length[r] = rows[r].len;
        c->h_len.p[r] = (uint32_t)rows[r].len;  // `length as u32`, editor/mod.rs:95
    }
    c->h_gap_ptr.p[n] = (uint32_t)at;
    c->n_reads = (uint32_t)n;
    c->n_iv = 0;
    c->n_gaps = (uint32_t)at;
    c->from_report = true;
    if (!c->d_len.reserve(n + 1) || !c->d_gap_ptr.reserve(n + 1) || !c->d_gaps.reserve(at + 1) || !c->d_cls.reserve(n + 16) ||
        !c->d_bitmap.reserve(c->bitmap_bytes() + 4) || !c->d_counters.reserve(yb::kCounterWords))
        return c->fail(YB_ERR_NOMEM, "device allocation failed");
    YB_CUDA(c, cudaMemcpyAsync(c->d_len.p, c->h_len.p, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c->stream));
    YB_CUDA(c, cudaMemcpyAsync(c->d_gap_ptr.p, c->h_gap_ptr.p, sizeof(uint32_t) * (n + 1), cudaMemcpyHostToDevice, c->stream));
    if (at) YB_CUDA(c, cudaMemcpyAsync(c->d_gaps.p, c->h_gaps.p, sizeof(uint2) * at, cudaMemcpyHostToDevice, c->stream));
    YB_CUDA(c, cudaStreamSynchronize(c->stream));
    return YB_OK;
}

int yb_init_report(yb_ctx *c, const char *path) {
    if (!c || !path) return YB_ERR_INVALID_ARGUMENT;
    std::vector<char> text;
    if (int rc = slurp(c, path, &text)) return rc;
    return yb_init_report_buffer(c, text.data(), text.size());
}

}  // extern "C"
