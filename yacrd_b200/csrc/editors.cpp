// editors.cpp — the post-detection editors of yacrd on top of the device results:
//   scrubb  (reference src/editor/scrubbing.rs:34-236)  every bad region of a read is cut out
//   filter  (src/editor/filter.rs:34-228)               reads (or overlap records) that are not NotBad are dropped
//   extract (src/editor/extract.rs:34-232)              only those are kept
//   split   (src/editor/split.rs:34-226)                Chimeric reads are cut at their interior bad regions
// Host I/O only: per record they ask the context for (bad regions, length, class) — the class was computed on the
// device by the same editor::type_of_read the reference calls per record (editor/mod.rs:85-100) — and stream the
// sequence file through. Record syntax follows what the reference gets from noodles-fasta 0.45 / noodles-fastq 0.16
// (Cargo.lock): fastq is 4 lines per record, the '+' line is written bare, name and description are separated by the
// first space and written back with one space; fasta sequences are re-wrapped at 80 columns; fasta records cut by
// scrubb / split lose their description, fastq records keep it (scrubbing.rs:139-153 vs 211-224).
// Overlap files (filter / extract only) pass through line by line; csv quoting is not interpreted.
// Compression (util.rs:57-87, niffler): the input's codec is sniffed from its magic number (gzip, bzip2, xz) and the
// output is written with the same codec at level 1, as util.rs:84 asks (codec.cpp).
#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/yacrd_b200.h"
#include "store.hpp"

namespace yb {
namespace {

struct Err {
    int code = YB_OK;
    std::string msg;
    int set(int c, const char *fmt, ...) {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        code = c;
        msg = buf;
        return c;
    }
};

// Buffered line reader: lines end with '\n' (a preceding '\r' is dropped); the last line may lack the terminator.
class LineReader {
  public:
    LineReader(ByteSource *f, size_t cap) : f_(f), buf_(cap < (1u << 16) ? (1u << 16) : cap) {}
    // false at end of input. *line stays valid until the next call.
    bool next(std::string *line) {
        line->clear();
        bool any = false;
        for (;;) {
            if (pos_ == end_) {
                end_ = fill();
                pos_ = 0;
                if (end_ == 0) break;
            }
            any = true;
            const char *p = buf_.data() + pos_;
            const char *nl = static_cast<const char *>(memchr(p, '\n', end_ - pos_));
            if (nl) {
                line->append(p, (size_t)(nl - p));
                pos_ += (size_t)(nl - p) + 1;
                if (!line->empty() && line->back() == '\r') line->pop_back();
                return true;
            }
            line->append(p, end_ - pos_);
            pos_ = end_;
        }
        if (any && !line->empty() && line->back() == '\r') line->pop_back();
        return any;
    }
    bool io_error() const { return bad_; }
  private:
    size_t fill() {
        const long got = f_->read(buf_.data(), buf_.size());
        if (got < 0) bad_ = true;
        return got > 0 ? (size_t)got : 0;
    }
    ByteSource *f_;
    std::vector<char> buf_;
    size_t pos_ = 0, end_ = 0;
    bool bad_ = false;
};

class Out {
  public:
    Out(ByteSink *f, size_t cap) : f_(f), buf_(cap < (1u << 16) ? (1u << 16) : cap) {}
    void put(const char *p, size_t n) {
        if (n > buf_.size() - len_) flush();
        if (n > buf_.size()) {
            ok_ = ok_ && f_->write(p, n);
            return;
        }
        memcpy(buf_.data() + len_, p, n);
        len_ += n;
    }
    void put(const std::string &s) { put(s.data(), s.size()); }
    void put(char ch) { put(&ch, 1); }
    void flush() {
        if (len_) ok_ = ok_ && f_->write(buf_.data(), len_);
        len_ = 0;
    }
    bool ok() const { return ok_; }

  private:
    ByteSink *f_;
    std::vector<char> buf_;
    size_t len_ = 0;
    bool ok_ = true;
};

struct Lookup {
    yb_ctx *ctx;
    const uint32_t *gaps = nullptr;  // (begin, end) pairs
    uint32_t n_gaps = 0;
    uint64_t length = 0;
    uint8_t cls = YB_NOT_BAD;
    int get(const char *id, size_t n) { return yb_get_bad_part(ctx, id, n, &gaps, &n_gaps, &length, &cls); }
};

// first token of `s` in the sense of str::split_ascii_whitespace (scrubbing.rs:181-185)
void first_token(const std::string &s, const char **p, size_t *n) {
    size_t a = 0;
    auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\x0c' || c == '\r'; };
    while (a < s.size() && ws(s[a])) ++a;
    size_t b = a;
    while (b < s.size() && !ws(s[b])) ++b;
    *p = s.data() + a;
    *n = b - a;
}

// The pieces of a read an editor keeps, as (begin, end) positions; `whole` = the record is written unchanged.
struct Cut {
    bool drop = false, whole = false;
    std::vector<uint32_t> pos;  // consecutive pairs
};

void plan_scrubb(const Lookup &r, Cut *c) {  // scrubbing.rs:91-121
    c->pos.clear();
    c->drop = r.cls == YB_NOT_COVERED;
    c->whole = !c->drop && r.n_gaps == 0;
    if (c->drop || c->whole) return;
    std::vector<uint32_t> poss;
    poss.push_back(0);
    for (uint32_t g = 0; g < r.n_gaps; ++g) {
        poss.push_back(r.gaps[2 * g]);
        poss.push_back(r.gaps[2 * g + 1]);
    }
    if (poss.back() != (uint32_t)r.length) poss.push_back((uint32_t)r.length);
    const size_t from = (poss[0] == 0 && poss[1] == 0) ? 2 : 0;
    for (size_t i = from; i + 1 < poss.size(); i += 2) {  // chunks_exact(2): an odd tail is dropped
        c->pos.push_back(poss[i]);
        c->pos.push_back(poss[i + 1]);
    }
}

void plan_split(const Lookup &r, Cut *c) {  // split.rs:93-112
    c->pos.clear();
    c->drop = r.cls == YB_NOT_COVERED;
    c->whole = r.cls == YB_NOT_BAD;
    if (c->drop || c->whole) return;
    c->pos.push_back(0);
    for (uint32_t g = 0; g < r.n_gaps; ++g) {
        if (r.gaps[2 * g] == 0 || r.gaps[2 * g + 1] == (uint32_t)r.length) continue;
        c->pos.push_back(r.gaps[2 * g]);
        c->pos.push_back(r.gaps[2 * g + 1]);
    }
    c->pos.push_back((uint32_t)r.length);
}

void plan(int op, const Lookup &r, Cut *c) {
    if (op == YB_EDIT_SCRUBB) return plan_scrubb(r, c);
    if (op == YB_EDIT_SPLIT) return plan_split(r, c);
    c->pos.clear();
    c->whole = true;
    c->drop = op == YB_EDIT_FILTER ? r.cls != YB_NOT_BAD : r.cls == YB_NOT_BAD;  // filter.rs:92, extract.rs:91
}

void put_wrapped(Out &out, const char *seq, size_t n) {  // noodles fasta writer: 80 bases per line
    for (size_t i = 0; i < n; i += 80) {
        out.put(seq + i, n - i < 80 ? n - i : 80);
        out.put('\n');
    }
}

int edit_fasta(int op, LineReader &in, Out &out, Lookup &look, Err *err) {
    std::string line, name, desc, seq, piece;
    Cut cut;
    bool have = false, pending = in.next(&line);
    while (pending) {
        if (line.empty() && !have) {  // blank lines before a record
            pending = in.next(&line);
            continue;
        }
        if (line[0] != '>') return err->set(YB_ERR_READING, "Reading of the file in fasta format failed: record does not start with '>'");
        // definition: name up to the first whitespace, the rest is the description
        size_t sp = 1;
        while (sp < line.size() && line[sp] != ' ' && line[sp] != '\t') ++sp;
        name.assign(line, 1, sp - 1);
        desc = sp < line.size() ? line.substr(sp + 1) : std::string();
        seq.clear();
        have = true;
        while ((pending = in.next(&line)) && (line.empty() || line[0] != '>')) seq += line;
        if (int rc = look.get(name.data(), name.size())) return rc;
        plan(op, look, &cut);
        if (cut.drop) continue;
        if (cut.whole) {
            out.put('>');
            out.put(name);
            if (!desc.empty()) {
                out.put(' ');
                out.put(desc);
            }
            out.put('\n');
            put_wrapped(out, seq.data(), seq.size());
            continue;
        }
        for (size_t i = 0; i + 1 < cut.pos.size(); i += 2) {
            const uint32_t b = cut.pos[i], e = cut.pos[i + 1];
            if (b > seq.size() || e > seq.size()) {
                fprintf(stderr,
                        "For read %s %s position is larger than read, it's strange check your data. For this read, this split position and "
                        "next are ignore.\n",
                        name.c_str(), op == YB_EDIT_SCRUBB ? "scrubb" : "split");
                break;
            }
            char tag[64];
            snprintf(tag, sizeof tag, "_%u_%u", b, e);
            out.put('>');
            out.put(name);
            out.put(tag, strlen(tag));
            out.put('\n');
            if (e > b) put_wrapped(out, seq.data() + b, e - b);
        }
    }
    if (in.io_error()) return err->set(YB_ERR_READING, "Reading of the file in fasta format failed");
    return YB_OK;
}

int edit_fastq(int op, LineReader &in, Out &out, Lookup &look, Err *err) {
    std::string def, seq, plus, qual, name, desc;
    Cut cut;
    while (in.next(&def)) {
        if (def.empty()) continue;
        if (def[0] != '@' || !in.next(&seq) || !in.next(&plus) || plus.empty() || plus[0] != '+' || !in.next(&qual))
            return err->set(YB_ERR_READING, "Reading of the file in fastq format failed: truncated or malformed record");
        const size_t sp = def.find(' ', 1);
        name.assign(def, 1, sp == std::string::npos ? std::string::npos : sp - 1);
        desc = sp == std::string::npos ? std::string() : def.substr(sp + 1);
        const char *id;
        size_t idn;
        first_token(name, &id, &idn);
        if (int rc = look.get(id, idn)) return rc;
        plan(op, look, &cut);
        if (cut.drop) continue;
        auto put_def = [&](const char *tag) {
            out.put('@');
            out.put(name);
            if (tag) out.put(tag, strlen(tag));
            if (!desc.empty()) {
                out.put(' ');
                out.put(desc);
            }
            out.put('\n');
        };
        if (cut.whole) {
            put_def(nullptr);
            out.put(seq);
            out.put("\n+\n", 3);
            out.put(qual);
            out.put('\n');
            continue;
        }
        for (size_t i = 0; i + 1 < cut.pos.size(); i += 2) {
            const uint32_t b = cut.pos[i], e = cut.pos[i + 1];
            if (b > seq.size() || e > seq.size()) {
                fprintf(stderr,
                        "For read %s %s position is larger than read, it's strange check your data. For this read, this split position and "
                        "next are ignore.\n",
                        name.c_str(), op == YB_EDIT_SCRUBB ? "scrubb" : "split");
                break;
            }
            char tag[64];
            snprintf(tag, sizeof tag, "_%u_%u", b, e);
            put_def(tag);
            const uint32_t qe = e <= qual.size() ? e : (uint32_t)qual.size(), qb = b <= qe ? b : qe;
            out.put(seq.data() + b, e > b ? e - b : 0);
            out.put("\n+\n", 3);
            out.put(qual.data() + qb, qe - qb);
            out.put('\n');
        }
    }
    if (in.io_error()) return err->set(YB_ERR_READING, "Reading of the file in fastq format failed");
    return YB_OK;
}

// filter.rs:139-228 / extract.rs:139-232: a record is kept when both reads are NotBad (filter) or when one is not (extract)
int edit_overlaps(int op, char delim, int col_b, LineReader &in, Out &out, Lookup &look, Err *err) {
    std::string line;
    while (in.next(&line)) {
        if (line.empty()) continue;
        const size_t a_end = line.find(delim);
        size_t b0 = 0, b1 = std::string::npos;
        int col = 0;
        for (size_t p = 0; p <= line.size(); ++p) {
            if (p == line.size() || line[p] == delim) {
                if (col == col_b) {
                    b1 = p;
                    break;
                }
                ++col;
                b0 = p + 1;
            }
        }
        if (a_end == std::string::npos || b1 == std::string::npos)
            return err->set(YB_ERR_READING, "Reading of the file in %s format failed: record has too few columns", delim == '\t' ? "paf" : "m4");
        if (int rc = look.get(line.data(), a_end)) return rc;
        const bool a_ok = look.cls == YB_NOT_BAD;
        if (int rc = look.get(line.data() + b0, b1 - b0)) return rc;
        const bool b_ok = look.cls == YB_NOT_BAD;
        const bool keep = op == YB_EDIT_FILTER ? (a_ok && b_ok) : (!a_ok || !b_ok);
        if (keep) {
            out.put(line);
            out.put('\n');
        }
    }
    if (in.io_error()) return err->set(YB_ERR_READING, "Reading of the overlap file failed");
    return YB_OK;
}

}  // namespace

int run_editor(yb_ctx *ctx, int op, const char *input_path, const char *output_path, size_t buffer_size, std::string *error) {
    static const char *const kOpName[] = {"scrubbing", "filter", "extract", "split"};
    Err err;
    auto done = [&](int rc) {
        if (rc != YB_OK && error && !err.msg.empty()) *error = err.msg + " (Filename: " + input_path + ")";
        return rc;
    };
    if (op < YB_EDIT_SCRUBB || op > YB_EDIT_SPLIT) return done(err.set(YB_ERR_INVALID_ARGUMENT, "unknown editor %d", op));
    const int t = yb_file_type(input_path);
    if (t == 0 || t == 'o') return done(err.set(YB_ERR_UNKNOWN_FORMAT, "Format detection for '%s' file not possible", input_path));
    if (t == 'y' || ((t == 'p' || t == 'm') && (op == YB_EDIT_SCRUBB || op == YB_EDIT_SPLIT)))
        return done(err.set(YB_ERR_WRONG_FORMAT, "Can't run %s on %s file %s", kOpName[op], t == 'y' ? "yacrd" : (t == 'p' ? "paf" : "m4"),
                            input_path));
    Codec codec = kPlain;
    std::string why;
    ByteSource *fi = open_source(input_path, &codec, &why);
    if (!fi) return done(err.set(YB_ERR_CANT_READ_FILE, "%s", why.c_str()));
    ByteSink *fo = open_sink(output_path, codec, &why);  // the output keeps the input's compression (util.rs:84)
    if (!fo) {
        delete fi;
        return done(err.set(YB_ERR_CANT_WRITE_FILE, "%s", why.c_str()));
    }
    LineReader in(fi, buffer_size);
    Out out(fo, buffer_size);
    Lookup look{ctx};
    int rc;
    if (t == 'a') rc = edit_fasta(op, in, out, look, &err);
    else if (t == 'q') rc = edit_fastq(op, in, out, look, &err);
    else rc = edit_overlaps(op, t == 'p' ? '\t' : ' ', t == 'p' ? 5 : 1, in, out, look, &err);
    out.flush();
    delete fi;
    const bool wrote = fo->close() && out.ok();
    delete fo;
    if (rc == YB_OK && !wrote) rc = err.set(YB_ERR_WRITING, "Writing of the file %s failed", output_path);
    if (rc != YB_OK && err.msg.empty()) return rc;  // the context already holds the message (yb_get_bad_part)
    return done(rc);
}

}  // namespace yb
