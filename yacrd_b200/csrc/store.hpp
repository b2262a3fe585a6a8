// store.hpp — host-side mirror of the reference's `Reads2Ovl` producer surface
// (src/reads2ovl/mod.rs:43-163) with FullMemory's semantics (src/reads2ovl/fullmemory.rs:46-90),
// re-designed for a device consumer: instead of FxHashMap<String,(Vec<(u32,u32)>,usize)> it interns
// read ids to dense first-seen indices and keeps intervals as arrival-order records that are frozen
// into one CSR (flat (begin,end) buffer + row pointers + lengths) — the layout the kernels stream.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

namespace yb {

// Open-addressing id -> dense index table over a byte arena (ids are copied once, on first sight).
class IdTable {
  public:
    static constexpr uint32_t kNone = 0xFFFFFFFFu;
    IdTable() { slots_.assign(1024, kNone); }
    uint32_t size() const { return (uint32_t)(off_.size() - 1); }
    const char *id(uint32_t i, size_t *len) const {
        *len = (size_t)(off_[i + 1] - off_[i]);
        return bytes_.data() + off_[i];
    }
    uint32_t find(const char *s, size_t n) const {
        const uint64_t h = hash(s, n);
        const size_t mask = slots_.size() - 1;
        for (size_t p = (size_t)h & mask;; p = (p + 1) & mask) {
            const uint32_t i = slots_[p];
            if (i == kNone) return kNone;
            if (equal(i, s, n)) return i;
        }
    }
    // Returns the index of `s`, inserting it if new (*inserted tells which).
    uint32_t intern(const char *s, size_t n, bool *inserted) {
        const uint64_t h = hash(s, n);
        size_t mask = slots_.size() - 1;
        size_t p = (size_t)h & mask;
        for (;; p = (p + 1) & mask) {
            const uint32_t i = slots_[p];
            if (i == kNone) break;
            if (equal(i, s, n)) {
                *inserted = false;
                return i;
            }
        }
        const uint32_t idx = size();
        bytes_.insert(bytes_.end(), s, s + n);
        off_.push_back((uint64_t)bytes_.size());
        slots_[p] = idx;
        if ((uint64_t)(idx + 1) * 10 > (uint64_t)slots_.size() * 6) grow();
        *inserted = true;
        return idx;
    }
    void clear() {
        bytes_.clear();
        off_.assign(1, 0);
        slots_.assign(1024, kNone);
    }
    // Takes over an arena of distinct ids already in index order (parallel ingestion) and builds the lookup table.
    void adopt(std::vector<char> &&bytes, std::vector<uint64_t> &&off) {
        bytes_ = std::move(bytes);
        off_ = std::move(off);
        size_t cap = 1024;
        while ((uint64_t)size() * 10 > (uint64_t)cap * 5) cap <<= 1;
        slots_.assign(cap, kNone);
        const size_t mask = cap - 1;
        for (uint32_t i = 0; i < size(); ++i) {
            size_t n;
            const char *s = id(i, &n);
            size_t p = (size_t)hash(s, n) & mask;
            while (slots_[p] != kNone) p = (p + 1) & mask;
            slots_[p] = i;
        }
    }

  private:
    static uint64_t hash(const char *s, size_t n) {
        uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)n;
        while (n >= 8) {
            uint64_t w;
            memcpy(&w, s, 8);
            h = (h ^ w) * 0xD6E8FEB86659FD93ull;
            h ^= h >> 32;
            s += 8;
            n -= 8;
        }
        uint64_t w = 0;
        memcpy(&w, s, n);
        h = (h ^ w) * 0xD6E8FEB86659FD93ull;
        h ^= h >> 29;
        return h * 0x9E3779B97F4A7C15ull >> 7;
    }
    bool equal(uint32_t i, const char *s, size_t n) const {
        const uint64_t a = off_[i], b = off_[i + 1];
        return (size_t)(b - a) == n && memcmp(bytes_.data() + a, s, n) == 0;
    }
    void grow() {
        std::vector<uint32_t> ns(slots_.size() * 2, kNone);
        const size_t mask = ns.size() - 1;
        for (uint32_t i = 0; i < size(); ++i) {
            size_t n;
            const char *s = id(i, &n);
            size_t p = (size_t)hash(s, n) & mask;
            while (ns[p] != kNone) p = (p + 1) & mask;
            ns[p] = i;
        }
        slots_.swap(ns);
    }
    std::vector<char> bytes_;
    std::vector<uint64_t> off_{0};
    std::vector<uint32_t> slots_;
};

struct PendingRecord {
    uint32_t read, begin, end;
};

// Parse status of the text ingesters (ingest.cpp).
struct IngestError {
    int code = 0;         // yb_status
    uint64_t line = 0;    // 1-based record number
    std::string message;
};

class Engine;  // capi.cu

// Parses a whole PAF (format 'p': tab-delimited, first 9 columns, src/io.rs:24-34) or BLASR m4
// (format 'm': space-delimited, 12 columns, src/io.rs:37-50) buffer and feeds
// add_overlap_and_length twice per record (src/reads2ovl/mod.rs:83-145). Returns false on error.
typedef bool (*AddFn)(void *sink, const char *id, size_t id_len, uint32_t b, uint32_t e, uint64_t len);
bool ingest_buffer(const char *text, size_t n, int format, AddFn add, void *sink, IngestError *err);

// Multi-threaded form with the same observable result (first-seen read order, first-seen lengths, arrival order
// inside a read), producing the CSR directly. `alloc` hands out the (pinned) rowptr[n_reads + 1], len[n_reads] and
// iv[2 * n_iv] arrays once their sizes are known; the ids come back as one arena in read order.
struct BulkIds {
    std::vector<char> bytes;
    std::vector<uint64_t> off;     // n_reads + 1
    std::vector<uint64_t> length;  // first-seen length of every read (usize in the reference)
    uint32_t n_reads = 0;
    uint64_t n_iv = 0;
};
typedef bool (*CsrAllocFn)(void *sink, size_t n_reads, size_t n_iv, uint32_t **rowptr, uint32_t **len, uint32_t **iv);
bool ingest_buffer_parallel(const char *text, size_t n, int format, int threads, CsrAllocFn alloc, void *sink, BulkIds *ids,
                            IngestError *err);

// codec.cpp: compressed files (util.rs:57-87). open_source sniffs the magic number; open_sink writes `codec` at level 1.
enum Codec { kPlain = 0, kGzip = 1, kBzip2 = 2, kXz = 3 };
class ByteSource {
  public:
    virtual ~ByteSource() {}
    virtual long read(void *buf, size_t n) = 0;  // bytes read, 0 at the end, -1 on a corrupt stream / read error
};
class ByteSink {
  public:
    virtual ~ByteSink() {}
    virtual bool write(const void *p, size_t n) = 0;
    virtual bool close() = 0;  // finishes the stream; false if anything failed
};
ByteSource *open_source(const char *path, Codec *codec, std::string *err);
ByteSink *open_sink(const char *path, Codec codec, std::string *err);

// editors.cpp: runs one post-detection editor (yb_editor) over input_path -> output_path, asking `ctx` for the
// results through the public ABI. Returns a yb_status; *error gets the message unless the context already has it.
int run_editor(struct ::yb_ctx *ctx, int op, const char *input_path, const char *output_path, size_t buffer_size, std::string *error);

}  // namespace yb
