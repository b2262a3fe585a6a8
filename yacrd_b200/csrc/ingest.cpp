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
#include <string.h>

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

}  // namespace

bool ingest_buffer(const char *text, size_t n, int format, AddFn add, void *sink, IngestError *err) {
    const bool paf = format == 'p';
    const char delim = paf ? '\t' : ' ';
    const int need = paf ? 9 : 12;
    const char *p = text, *const end = text + n;
    uint64_t line = 0;
    Field f[12];
    while (p < end) {
        // one record: split at most `need` leading fields, then skip to the terminator
        int nf = 0;
        const char *q = p;
        const char *fs = p;
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
        // terminator: "\r\n", "\n" or "\r"
        if (q < end) {
            if (*q == '\r' && q + 1 < end && q[1] == '\n') ++q;
            ++q;
        }
        p = q;
        if (empty) continue;
        ++line;
        uint64_t la, ba, ea, lb, bb, eb;
        const Field *ida, *idb;
        bool ok = nf >= need;
        if (ok) {
            if (paf) {  // io.rs:24-34
                ida = &f[0];
                idb = &f[5];
                ok = parse_unsigned(f[1], UINT64_MAX, &la) && parse_unsigned(f[2], UINT32_MAX, &ba) &&
                     parse_unsigned(f[3], UINT32_MAX, &ea) && is_single_char(f[4]) &&
                     parse_unsigned(f[6], UINT64_MAX, &lb) && parse_unsigned(f[7], UINT32_MAX, &bb) &&
                     parse_unsigned(f[8], UINT32_MAX, &eb);
            } else {  // io.rs:37-50
                uint64_t shared;
                ida = &f[0];
                idb = &f[1];
                ok = is_rust_float(f[2]) && parse_unsigned(f[3], UINT64_MAX, &shared) &&
                     is_single_char(f[4]) && parse_unsigned(f[5], UINT32_MAX, &ba) &&
                     parse_unsigned(f[6], UINT32_MAX, &ea) && parse_unsigned(f[7], UINT64_MAX, &la) &&
                     is_single_char(f[8]) && parse_unsigned(f[9], UINT32_MAX, &bb) &&
                     parse_unsigned(f[10], UINT32_MAX, &eb) && parse_unsigned(f[11], UINT64_MAX, &lb);
            }
        }
        if (!ok) {
            err->code = YB_ERR_READING;
            err->line = line;
            err->message = std::string("Reading of the file at record ") + std::to_string(line) +
                           " impossible, file in " + (paf ? "paf" : "m4") + " format";
            return false;
        }
        // mod.rs:108-109 / 140-141: A first, then B
        if (!add(sink, ida->p, ida->n, (uint32_t)ba, (uint32_t)ea, la) ||
            !add(sink, idb->p, idb->n, (uint32_t)bb, (uint32_t)eb, lb)) {
            err->code = YB_ERR_NOMEM;
            err->line = line;
            err->message = "out of memory while ingesting";
            return false;
        }
    }
    return true;
}

}  // namespace yb
