// ingest.cpp — PAF / BLASR-m4 text ingestion, host side of Reads2Ovl::init
// (reference src/reads2ovl/mod.rs:44-145, record layouts src/io.rs:24-50).
//
// The reference uses csv 1.3 (flexible, no headers, delimiter '\t' for PAF and ' ' for m4) + serde
// positional deserialisation of the first 9 / 12 columns; extra columns are ignored, a record with
// too few columns or a column that does not parse as its serde type is ReadingErrorNoFilename
// (mod.rs:93-97,125-129). This tokenizer restates those rules on a flat byte buffer:
//   - records end at "\n", "\r\n" or "\r"; empty records are skipped (csv Terminator::CRLF);
//   - integers follow Rust's FromStr for unsigned types: optional '+', decimal digits, no blanks,
//     overflow is an error; `char` columns must hold exactly one UTF-8 scalar; the m4 f64 column
//     must be a Rust float literal.
// Divergence (documented in DESIGN.md): csv's double-quote field quoting is not interpreted —
// minimap2 / BLASR never emit quoted fields.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/yacrd_b200.h"
#include "store.hpp"

namespace yb {
namespace {

struct Field {
    const char *p;
    size_t n;
};

inline bool parse_unsigned(const Field &f, uint64_t max, uint64_t *out) {
    const char *p = f.p;
    size_t n = f.n;
    if (n && *p == '+') {
        ++p;
        --n;
    }
    if (n == 0) return false;
    uint64_t v = 0;
    if (n <= 18) {  // cannot overflow 64 bits: one range test at the end
        for (size_t i = 0; i < n; ++i) {
            const unsigned d = (unsigned)(p[i] - '0');
            if (d > 9) return false;
            v = v * 10 + d;
        }
        if (v > max) return false;
        *out = v;
        return true;
    }
    for (size_t i = 0; i < n; ++i) {
        const unsigned d = (unsigned)(p[i] - '0');
        if (d > 9) return false;
        if (v > (max - d) / 10) return false;
        v = v * 10 + d;
    }
    *out = v;
    return true;
}

inline bool is_single_char(const Field &f) {
    if (f.n == 0) return false;
    const unsigned char c = (unsigned char)f.p[0];
    size_t need = c < 0x80 ? 1 : (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 0;
    if (need == 0 || f.n != need) return false;
    for (size_t i = 1; i < need; ++i)
        if (((unsigned char)f.p[i] >> 6) != 0x2) return false;
    return true;
}

inline bool ieq(const char *p, size_t n, const char *lit) {
    if (strlen(lit) != n) return false;
    for (size_t i = 0; i < n; ++i)
        if ((p[i] | 0x20) != lit[i]) return false;
    return true;
}

inline bool is_rust_float(const Field &f) {
    const char *p = f.p;
    size_t n = f.n;
    if (n && (*p == '+' || *p == '-')) {
        ++p;
        --n;
    }
    if (n == 0) return false;
    if (ieq(p, n, "inf") || ieq(p, n, "infinity") || ieq(p, n, "nan")) return true;
    size_t i = 0, digits = 0;
    while (i < n && p[i] >= '0' && p[i] <= '9') ++i, ++digits;
    if (i < n && p[i] == '.') {
        ++i;
        while (i < n && p[i] >= '0' && p[i] <= '9') ++i, ++digits;
    }
    if (digits == 0) return false;
    if (i < n && (p[i] == 'e' || p[i] == 'E')) {
        ++i;
        if (i < n && (p[i] == '+' || p[i] == '-')) ++i;
        size_t ed = 0;
        while (i < n && p[i] >= '0' && p[i] <= '9') ++i, ++ed;
        if (ed == 0) return false;
    }
    return i == n;
}

// One parsed record: both (begin, end, length) triples and where the two ids sit in the text.
struct Parsed {
    const char *ida, *idb;
    uint32_t na, nb;
    uint32_t ba, ea, bb, eb;
    uint64_t la, lb;
};

// Parses the record starting at p (p < end). Returns 1 = record, 0 = empty record (skipped), -1 = a record that
// does not deserialize (mod.rs:93-97,125-129). *next is the first byte after the record's terminator.
inline int parse_one(const char *p, const char *end, bool paf, Parsed *out, const char **next) {
    const char delim = paf ? '\t' : ' ';
    const int need = paf ? 9 : 12;
    Field f[12];
    int nf = 0;
    const char *q = p;
    const char *fs = p;
#if defined(__SSE2__)
    {   // 16 bytes at a time: the positions of delimiters and terminators come out of three compares and a movemask
        const __m128i vd = _mm_set1_epi8(delim), vn = _mm_set1_epi8('\n'), vr = _mm_set1_epi8('\r');
        bool done = false;
        while (!done && q + 16 <= end) {
            const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i *>(q));
            unsigned m = (unsigned)_mm_movemask_epi8(_mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(x, vd), _mm_cmpeq_epi8(x, vn)), _mm_cmpeq_epi8(x, vr)));
            while (m) {
                const char *c = q + __builtin_ctz(m);
                m &= m - 1;
                if (*c != delim) {  // terminator
                    q = c;
                    done = true;
                    break;
                }
                if (nf < need) {
                    f[nf].p = fs;
                    f[nf].n = (size_t)(c - fs);
                    ++nf;
                    fs = c + 1;
                }
            }
            if (!done) q += 16;
        }
    }
#endif
    while (q < end && *q != '\n' && *q != '\r') {
        if (*q == delim && nf < need) {
            f[nf].p = fs;
            f[nf].n = (size_t)(q - fs);
            ++nf;
            fs = q + 1;
        }
        ++q;
    }
    if (nf < need) {
        f[nf].p = fs;
        f[nf].n = (size_t)(q - fs);
        ++nf;
    }
    const bool empty = q == p;
    if (q < end) {  // terminator: "\r\n", "\n" or "\r"
        if (*q == '\r' && q + 1 < end && q[1] == '\n') ++q;
        ++q;
    }
    *next = q;
    if (empty) return 0;
    if (nf < need) return -1;
    uint64_t la, ba, ea, lb, bb, eb;
    const Field *ida, *idb;
    bool ok;
    if (paf) {  // io.rs:24-34
        ida = &f[0];
        idb = &f[5];
        ok = parse_unsigned(f[1], UINT64_MAX, &la) && parse_unsigned(f[2], UINT32_MAX, &ba) &&
             parse_unsigned(f[3], UINT32_MAX, &ea) && is_single_char(f[4]) && parse_unsigned(f[6], UINT64_MAX, &lb) &&
             parse_unsigned(f[7], UINT32_MAX, &bb) && parse_unsigned(f[8], UINT32_MAX, &eb);
    } else {  // io.rs:37-50
        uint64_t shared;
        ida = &f[0];
        idb = &f[1];
        ok = is_rust_float(f[2]) && parse_unsigned(f[3], UINT64_MAX, &shared) && is_single_char(f[4]) &&
             parse_unsigned(f[5], UINT32_MAX, &ba) && parse_unsigned(f[6], UINT32_MAX, &ea) &&
             parse_unsigned(f[7], UINT64_MAX, &la) && is_single_char(f[8]) && parse_unsigned(f[9], UINT32_MAX, &bb) &&
             parse_unsigned(f[10], UINT32_MAX, &eb) && parse_unsigned(f[11], UINT64_MAX, &lb);
    }
    if (!ok) return -1;
    out->ida = ida->p;
    out->na = (uint32_t)ida->n;
    out->idb = idb->p;
    out->nb = (uint32_t)idb->n;
    out->ba = (uint32_t)ba;
    out->ea = (uint32_t)ea;
    out->bb = (uint32_t)bb;
    out->eb = (uint32_t)eb;
    out->la = la;
    out->lb = lb;
    return 1;
}

void reading_error(IngestError *err, uint64_t line, bool paf) {
    err->code = YB_ERR_READING;
    err->line = line;
    err->message = std::string("Reading of the file at record ") + std::to_string(line) + " impossible, file in " +
                   (paf ? "paf" : "m4") + " format";
}

inline uint64_t hash_id(const char *s, size_t n) {
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
    return h * 0x9E3779B97F4A7C15ull;
}

template <typename F> void parallel_for(int threads, F body) {  // body(t) on `threads` threads
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(body, t);
    body(0);
    for (auto &th : pool) th.join();
}

}  // namespace

bool ingest_buffer(const char *text, size_t n, int format, AddFn add, void *sink, IngestError *err) {
    const bool paf = format == 'p';
    const char *p = text, *const end = text + n;
    uint64_t line = 0;
    Parsed r;
    while (p < end) {
        const char *next;
        const int rc = parse_one(p, end, paf, &r, &next);
        p = next;
        if (rc == 0) continue;
        ++line;
        if (rc < 0) {
            reading_error(err, line, paf);
            return false;
        }
        // mod.rs:108-109 / 140-141: A first, then B
        if (!add(sink, r.ida, r.na, r.ba, r.ea, r.la) || !add(sink, r.idb, r.nb, r.bb, r.eb, r.lb)) {
            err->code = YB_ERR_NOMEM;
            err->line = line;
            err->message = "out of memory while ingesting";
            return false;
        }
    }
    return true;
}

// Parallel form of the same ingestion. Same observable result as the sequential loop
// above: reads are numbered in first-seen order over the whole file, a read's length is the one on the record
// that first mentions it (fullmemory.rs:82-90), and inside a read the intervals keep their arrival order
// (Reads2Ovl::overlap, mod.rs:150; pinned by reads2ovl/mod.rs:181-237).
//   1. the text is cut at record boundaries into one chunk per thread; every thread tokenizes its chunk;
//   2. ids are interned per hash partition: the partition's owner walks the chunks in file order, so the first
//      occurrence (record number, side) of every id is exact;
//   3. distinct ids are sorted by first occurrence -> dense first-seen index;
//   4. per-thread per-read counts -> row pointers and, per thread, where its intervals of each read start;
//   5. every thread writes its intervals in place.
bool ingest_buffer_parallel(const char *text, size_t n, int format, int threads, CsrAllocFn alloc, void *sink, BulkIds *ids,
                            IngestError *err) {
    const bool paf = format == 'p';
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    const int T = threads;
    const bool trace = getenv("YB_INGEST_TRACE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[ingest] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    const char *const end = text + n;
    // ---- 1. chunks and tokenizing ----
    std::vector<const char *> cut(T + 1);
    cut[0] = text;
    cut[T] = end;
    for (int t = 1; t < T; ++t) {
        const char *p = text + (n / T) * t;
        if (p < cut[t - 1]) p = cut[t - 1];
        while (p > text && p < end && p[-1] != '\n' && p[-1] != '\r') ++p;  // start of a record
        if (p > text && p < end && p[-1] == '\r' && *p == '\n') ++p;       // never split a "\r\n"
        cut[t] = p;
    }
    struct Rec {
        uint32_t offa, offb;  // id offsets from the chunk start
        uint16_t na, nb;
        uint32_t ba, ea, bb, eb;
        uint64_t la, lb;
        uint32_t ra, rb;      // partition-local id, later the global read index
    };
    struct Chunk {
        std::vector<Rec> recs;
        std::vector<std::vector<uint32_t>> by_part;  // per partition: record index * 2 + side, in file order
        uint64_t bad_line = 0;                       // 1-based local number of the first record that does not parse
        uint64_t lines = 0;
        bool too_long = false;
    };
    int P = 1;
    while (P < T) P <<= 1;
    std::vector<Chunk> chunks(T);
    parallel_for(T, [&](int t) {
        Chunk &c = chunks[t];
        c.by_part.resize(P);
        const char *base = cut[t], *p = base, *const e = cut[t + 1];
        if ((size_t)(e - base) > 0xFFFFFFF0ull) {
            c.too_long = true;
            return;
        }
        c.recs.reserve((size_t)(e - base) / 64 + 16);
        Parsed r;
        while (p < e) {
            const char *next;
            const int rc = parse_one(p, e, paf, &r, &next);
            p = next;
            if (rc == 0) continue;
            ++c.lines;
            if (rc < 0 || r.na > 0xFFFFu || r.nb > 0xFFFFu) {
                c.bad_line = c.lines;
                return;
            }
            const uint32_t i = (uint32_t)c.recs.size();
            c.recs.push_back(Rec{(uint32_t)(r.ida - base), (uint32_t)(r.idb - base), (uint16_t)r.na, (uint16_t)r.nb, r.ba, r.ea,
                                 r.bb, r.eb, r.la, r.lb, 0u, 0u});
            c.by_part[(size_t)(hash_id(r.ida, r.na) >> 40) & (size_t)(P - 1)].push_back(i * 2u);
            c.by_part[(size_t)(hash_id(r.idb, r.nb) >> 40) & (size_t)(P - 1)].push_back(i * 2u + 1u);
        }
    });
    lap("tokenize");
    uint64_t line_base = 0;
    std::vector<uint64_t> rec_base(T + 1, 0);
    for (int t = 0; t < T; ++t) {
        if (chunks[t].too_long) {
            err->code = YB_ERR_TOO_LARGE;
            err->message = "input chunk larger than 4 GiB per thread; use more threads";
            return false;
        }
        if (chunks[t].bad_line) {
            reading_error(err, line_base + chunks[t].bad_line, paf);
            return false;
        }
        line_base += chunks[t].lines;
        rec_base[t + 1] = rec_base[t] + chunks[t].recs.size();
    }
    const uint64_t n_rec = rec_base[T];
    if (2 * n_rec > 0xFFFFFFF0ull) {
        err->code = YB_ERR_TOO_LARGE;
        err->message = "more than 2^32-16 intervals in one context";
        return false;
    }
    // ---- 2. interning per partition, in file order ----
    struct Entry {
        uint64_t first;   // (global record number) * 2 + side of the first occurrence
        const char *id;
        uint32_t n;
        uint64_t len;     // length on that record
    };
    struct Part {
        std::vector<Entry> ent;
        std::vector<uint32_t> slots;
        bool overflow = false;
    };
    std::vector<Part> parts(P);
    parallel_for(T, [&](int t) {
        for (int pi = t; pi < P; pi += T) {
            Part &pt = parts[pi];
            size_t expect = 0;
            for (int ct = 0; ct < T; ++ct) expect += chunks[ct].by_part[pi].size();
            size_t cap = 1024;
            while (cap < expect / 8 + 16) cap <<= 1;  // ~ distinct ids guess; grows below
            pt.slots.assign(cap, 0xFFFFFFFFu);
            for (int ct = 0; ct < T; ++ct) {
                Chunk &c = chunks[ct];
                const char *base = cut[ct];
                // A lookup is three dependent cache misses (slot -> entry -> the entry's id bytes in the text). They are
                // taken out of the critical path by walking the list in blocks: hashes and slot prefetches for the
                // whole block, then entry prefetches, then id prefetches, then the real (sequential, order-preserving)
                // probes, which now hit the cache. Prefetches are hints only: entries inserted by the block itself are
                // still found by the probes.
                const std::vector<uint32_t> &list = c.by_part[pi];
                constexpr size_t kBlock = 16;
                for (size_t b0 = 0; b0 < list.size() && !pt.overflow; b0 += kBlock) {
                    const size_t bn = std::min(kBlock, list.size() - b0);
                    const char *bs[kBlock];
                    uint32_t bl[kBlock];
                    uint64_t bh[kBlock];
                    {
                        const size_t mask = pt.slots.size() - 1;
                        for (size_t i = 0; i < bn; ++i) {
                            const uint32_t code = list[b0 + i];
                            const Rec &r = c.recs[code >> 1];
                            bs[i] = base + ((code & 1u) ? r.offb : r.offa);
                            bl[i] = (code & 1u) ? r.nb : r.na;
                            bh[i] = hash_id(bs[i], bl[i]);
                            __builtin_prefetch(&pt.slots[(size_t)bh[i] & mask]);
                        }
                        for (size_t i = 0; i < bn; ++i) {
                            const uint32_t e = pt.slots[(size_t)bh[i] & mask];
                            if (e != 0xFFFFFFFFu) __builtin_prefetch(&pt.ent[e]);
                        }
                        for (size_t i = 0; i < bn; ++i) {
                            const uint32_t e = pt.slots[(size_t)bh[i] & mask];
                            if (e != 0xFFFFFFFFu) __builtin_prefetch(pt.ent[e].id);
                        }
                    }
                    for (size_t i = 0; i < bn; ++i) {
                        const uint32_t code = list[b0 + i];
                        Rec &r = c.recs[code >> 1];
                        const bool side = code & 1u;
                        const char *s = bs[i];
                        const uint32_t sn = bl[i];
                        const uint64_t h = bh[i];
                        size_t mask = pt.slots.size() - 1, pos = (size_t)h & mask;
                        uint32_t found = 0xFFFFFFFFu;
                        for (;; pos = (pos + 1) & mask) {
                            const uint32_t e = pt.slots[pos];
                            if (e == 0xFFFFFFFFu) break;
                            if (pt.ent[e].n == sn && memcmp(pt.ent[e].id, s, sn) == 0) {
                                found = e;
                                break;
                            }
                        }
                        if (found == 0xFFFFFFFFu) {
                            found = (uint32_t)pt.ent.size();
                            if (found >= (1u << 26)) {
                                pt.overflow = true;
                                break;
                            }
                            pt.ent.push_back(Entry{(rec_base[ct] + (code >> 1)) * 2 + (side ? 1u : 0u), s, sn, side ? r.lb : r.la});
                            pt.slots[pos] = found;
                            if ((pt.ent.size() + 1) * 10 > pt.slots.size() * 6) {  // grow
                                std::vector<uint32_t> ns(pt.slots.size() * 2, 0xFFFFFFFFu);
                                const size_t m2 = ns.size() - 1;
                                for (uint32_t q2 = 0; q2 < pt.ent.size(); ++q2) {
                                    size_t q = (size_t)hash_id(pt.ent[q2].id, pt.ent[q2].n) & m2;
                                    while (ns[q] != 0xFFFFFFFFu) q = (q + 1) & m2;
                                    ns[q] = q2;
                                }
                                pt.slots.swap(ns);
                            }
                        }
                        (side ? r.rb : r.ra) = found | ((uint32_t)pi << 26);  // partition in the top 6 bits
                    }
                }
            }
        }
    });
    lap("intern per partition");
    // ---- 3. first-seen order ----
    std::vector<uint64_t> part_base(P + 1, 0);
    for (int pi = 0; pi < P; ++pi) {
        if (parts[pi].overflow) {
            err->code = YB_ERR_TOO_LARGE;
            err->message = "too many reads";
            return false;
        }
        part_base[pi + 1] = part_base[pi] + parts[pi].ent.size();
    }
    const uint64_t n_reads = part_base[P];
    if (n_reads > 0xFFFFFFF0ull) {
        err->code = YB_ERR_TOO_LARGE;
        err->message = "too many reads";
        return false;
    }
    struct Key {
        uint64_t first;
        uint32_t part, local;
    };
    std::vector<Key> order(n_reads);
    parallel_for(T, [&](int t) {
        for (int pi = t; pi < P; pi += T)
            for (uint32_t i = 0; i < parts[pi].ent.size(); ++i) order[part_base[pi] + i] = Key{parts[pi].ent[i].first, (uint32_t)pi, i};
    });
    lap("collect keys");
    std::sort(order.begin(), order.end(), [](const Key &a, const Key &b) { return a.first < b.first; });
    lap("sort by first occurrence");
    std::vector<std::vector<uint32_t>> to_global(P);
    for (int pi = 0; pi < P; ++pi) to_global[pi].resize(parts[pi].ent.size());
    ids->off.assign(n_reads + 1, 0);
    ids->length.resize(n_reads);
    for (uint64_t g = 0; g < n_reads; ++g) {
        const Entry &e = parts[order[g].part].ent[order[g].local];
        to_global[order[g].part][order[g].local] = (uint32_t)g;
        ids->off[g + 1] = ids->off[g] + e.n;
        ids->length[g] = e.len;
    }
    ids->bytes.resize(ids->off[n_reads]);
    parallel_for(T, [&](int t) {
        const uint64_t g0 = n_reads * t / T, g1 = n_reads * (t + 1) / T;
        for (uint64_t g = g0; g < g1; ++g) {
            const Entry &e = parts[order[g].part].ent[order[g].local];
            memcpy(ids->bytes.data() + ids->off[g], e.id, e.n);
        }
    });
    lap("id arena");
    // ---- 4. counts -> row pointers ----
    uint32_t *rowptr = nullptr, *len32 = nullptr, *iv = nullptr;
    if (!alloc(sink, (size_t)n_reads, (size_t)(2 * n_rec), &rowptr, &len32, &iv)) {
        err->code = YB_ERR_NOMEM;
        err->message = "pinned host allocation failed";
        return false;
    }
    lap("alloc csr");
    std::vector<std::vector<uint32_t>> cnt(T);
    parallel_for(T, [&](int t) {
        cnt[t].assign(n_reads, 0u);
        for (Rec &r : chunks[t].recs) {
            r.ra = to_global[r.ra >> 26][r.ra & 0x3FFFFFFu];
            r.rb = to_global[r.rb >> 26][r.rb & 0x3FFFFFFu];
            cnt[t][r.ra]++;
            cnt[t][r.rb]++;
        }
    });
    std::vector<uint64_t> range_sum(T + 1, 0);
    parallel_for(T, [&](int t) {  // totals per read range
        const uint64_t g0 = n_reads * t / T, g1 = n_reads * (t + 1) / T;
        uint64_t s = 0;
        for (uint64_t g = g0; g < g1; ++g)
            for (int ct = 0; ct < T; ++ct) s += cnt[ct][g];
        range_sum[t + 1] = s;
    });
    for (int t = 0; t < T; ++t) range_sum[t + 1] += range_sum[t];
    parallel_for(T, [&](int t) {  // rowptr and, in cnt[ct][g], where thread ct's intervals of read g start
        const uint64_t g0 = n_reads * t / T, g1 = n_reads * (t + 1) / T;
        uint64_t at = range_sum[t];
        for (uint64_t g = g0; g < g1; ++g) {
            rowptr[g] = (uint32_t)at;
            for (int ct = 0; ct < T; ++ct) {
                const uint32_t c0 = cnt[ct][g];
                cnt[ct][g] = (uint32_t)at;
                at += c0;
            }
            const uint64_t l = ids->length[g];
            len32[g] = l > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)l;
        }
    });
    rowptr[n_reads] = (uint32_t)(2 * n_rec);
    lap("counts and row pointers");
    // ---- 5. fill ----
    parallel_for(T, [&](int t) {
        std::vector<uint32_t> &cur = cnt[t];
        for (const Rec &r : chunks[t].recs) {
            uint32_t a = cur[r.ra]++;
            iv[2 * (size_t)a] = r.ba;
            iv[2 * (size_t)a + 1] = r.ea;
            a = cur[r.rb]++;
            iv[2 * (size_t)a] = r.bb;
            iv[2 * (size_t)a + 1] = r.eb;
        }
    });
    lap("fill");
    ids->n_reads = (uint32_t)n_reads;
    ids->n_iv = 2 * n_rec;
    return true;
}

}  // namespace yb
