// codec.cpp — compressed input / output of the host side (reference src/util.rs:57-87: niffler sniffs the magic number of
// the input, the editors' output keeps the input's compression at level 1).
//   gzip   zlib (linked)
//   bzip2  libbz2.so.1.0, xz  liblzma.so.5: this image ships the runtime libraries without their headers, so the few
//          entry points are bound with dlopen and the (stable, public) lzma_stream layout is restated below. A machine
//          without the libraries gets a CantReadFile error that says so.
#include <dlfcn.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <vector>

#include "../../include/yacrd_b200.h"
#include "store.hpp"

namespace yb {
namespace {

// ---- plain / gzip ------------------------------------------------------------------------------
struct FileSource : ByteSource {
    FILE *f;
    explicit FileSource(FILE *fh) : f(fh) {}
    ~FileSource() override { fclose(f); }
    long read(void *buf, size_t n) override {
        const size_t got = fread(buf, 1, n, f);
        return got == 0 && ferror(f) ? -1 : (long)got;
    }
};
struct FileSink : ByteSink {
    FILE *f;
    explicit FileSink(FILE *fh) : f(fh) {}
    ~FileSink() override {
        if (f) fclose(f);
    }
    bool write(const void *p, size_t n) override { return fwrite(p, 1, n, f) == n; }
    bool close() override {
        const bool ok = fclose(f) == 0;
        f = nullptr;
        return ok;
    }
};
struct GzSource : ByteSource {
    gzFile g;
    explicit GzSource(gzFile h) : g(h) { gzbuffer(g, 1u << 18); }
    ~GzSource() override { gzclose(g); }
    long read(void *buf, size_t n) override {
        const int got = gzread(g, buf, (unsigned)(n > (1u << 30) ? (1u << 30) : n));
        if (got > 0) return got;
        int errnum = Z_OK;
        gzerror(g, &errnum);  // a stream that ends early reads as a short file unless this is asked
        return (got < 0 || (errnum != Z_OK && errnum != Z_STREAM_END)) ? -1 : 0;
    }
};
struct GzSink : ByteSink {
    gzFile g;
    explicit GzSink(gzFile h) : g(h) {}
    ~GzSink() override {
        if (g) gzclose(g);
    }
    bool write(const void *p, size_t n) override {
        const char *c = static_cast<const char *>(p);
        while (n) {
            const unsigned part = n > (1u << 30) ? (1u << 30) : (unsigned)n;
            if (gzwrite(g, c, part) != (int)part) return false;
            c += part;
            n -= part;
        }
        return true;
    }
    bool close() override {
        const bool ok = gzclose(g) == Z_OK;
        g = nullptr;
        return ok;
    }
};

// ---- bzip2 through dlopen ------------------------------------------------------------------------
struct Bz2Api {
    void *(*open)(const char *, const char *) = nullptr;
    int (*read)(void *, void *, int) = nullptr;
    int (*write)(void *, void *, int) = nullptr;
    void (*close)(void *) = nullptr;
    bool ok = false;
    Bz2Api() {
        void *h = dlopen("libbz2.so.1.0", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libbz2.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!h) return;
        open = reinterpret_cast<decltype(open)>(dlsym(h, "BZ2_bzopen"));
        read = reinterpret_cast<decltype(read)>(dlsym(h, "BZ2_bzread"));
        write = reinterpret_cast<decltype(write)>(dlsym(h, "BZ2_bzwrite"));
        close = reinterpret_cast<decltype(close)>(dlsym(h, "BZ2_bzclose"));
        ok = open && read && write && close;
    }
};
const Bz2Api &bz2() {
    static const Bz2Api api;
    return api;
}
struct Bz2Source : ByteSource {
    void *b;
    explicit Bz2Source(void *h) : b(h) {}
    ~Bz2Source() override { bz2().close(b); }
    long read(void *buf, size_t n) override { return (long)bz2().read(b, buf, (int)(n > (1u << 30) ? (1u << 30) : n)); }
};
struct Bz2Sink : ByteSink {
    void *b;
    explicit Bz2Sink(void *h) : b(h) {}
    ~Bz2Sink() override {
        if (b) bz2().close(b);
    }
    bool write(const void *p, size_t n) override {
        char *c = const_cast<char *>(static_cast<const char *>(p));
        while (n) {
            const int part = n > (1u << 30) ? (1 << 30) : (int)n;
            if (bz2().write(b, c, part) != part) return false;
            c += part;
            n -= (size_t)part;
        }
        return true;
    }
    bool close() override {
        bz2().close(b);
        b = nullptr;
        return true;
    }
};

// ---- xz through dlopen (lzma_stream as in lzma/base.h of xz 5.x) ---------------------------------
struct LzmaStream {
    const uint8_t *next_in;
    size_t avail_in;
    uint64_t total_in;
    uint8_t *next_out;
    size_t avail_out;
    uint64_t total_out;
    const void *allocator;
    void *internal;
    void *reserved_ptr1, *reserved_ptr2, *reserved_ptr3, *reserved_ptr4;
    uint64_t reserved_int1, reserved_int2;
    size_t reserved_int3, reserved_int4;
    int reserved_enum1, reserved_enum2;
};
constexpr int kLzmaRun = 0, kLzmaFinish = 3, kLzmaOk = 0, kLzmaStreamEnd = 1;
struct LzmaApi {
    int (*stream_decoder)(LzmaStream *, uint64_t, uint32_t) = nullptr;
    int (*easy_encoder)(LzmaStream *, uint32_t, int) = nullptr;
    int (*code)(LzmaStream *, int) = nullptr;
    void (*end)(LzmaStream *) = nullptr;
    bool ok = false;
    LzmaApi() {
        void *h = dlopen("liblzma.so.5", RTLD_NOW | RTLD_LOCAL);
        if (!h) return;
        stream_decoder = reinterpret_cast<decltype(stream_decoder)>(dlsym(h, "lzma_stream_decoder"));
        easy_encoder = reinterpret_cast<decltype(easy_encoder)>(dlsym(h, "lzma_easy_encoder"));
        code = reinterpret_cast<decltype(code)>(dlsym(h, "lzma_code"));
        end = reinterpret_cast<decltype(end)>(dlsym(h, "lzma_end"));
        ok = stream_decoder && easy_encoder && code && end;
    }
};
const LzmaApi &lzma() {
    static const LzmaApi api;
    return api;
}
struct XzSource : ByteSource {
    FILE *f;
    LzmaStream s;
    std::vector<uint8_t> in;
    bool eof = false, done = false, bad = false;
    explicit XzSource(FILE *fh) : f(fh), in(1u << 18) {
        memset(&s, 0, sizeof s);
        bad = lzma().stream_decoder(&s, UINT64_MAX, 0x08u /* LZMA_CONCATENATED */) != kLzmaOk;
    }
    ~XzSource() override {
        lzma().end(&s);
        fclose(f);
    }
    long read(void *buf, size_t n) override {
        if (bad) return -1;
        s.next_out = static_cast<uint8_t *>(buf);
        s.avail_out = n;
        while (s.avail_out && !done) {
            if (s.avail_in == 0 && !eof) {
                s.next_in = in.data();
                s.avail_in = fread(in.data(), 1, in.size(), f);
                if (s.avail_in == 0) eof = true;
            }
            const int rc = lzma().code(&s, eof ? kLzmaFinish : kLzmaRun);
            if (rc == kLzmaStreamEnd) done = true;
            else if (rc != kLzmaOk) {
                bad = true;
                return -1;
            }
        }
        return (long)(n - s.avail_out);
    }
};
struct XzSink : ByteSink {
    FILE *f;
    LzmaStream s;
    std::vector<uint8_t> out;
    bool bad = false;
    explicit XzSink(FILE *fh) : f(fh), out(1u << 18) {
        memset(&s, 0, sizeof s);
        bad = lzma().easy_encoder(&s, 1u, 4 /* LZMA_CHECK_CRC64 */) != kLzmaOk;
    }
    ~XzSink() override {
        if (f) close();
    }
    bool pump(int action) {
        for (;;) {
            s.next_out = out.data();
            s.avail_out = out.size();
            const int rc = lzma().code(&s, action);
            const size_t got = out.size() - s.avail_out;
            if (got && fwrite(out.data(), 1, got, f) != got) return false;
            if (rc == kLzmaStreamEnd) return true;
            if (rc != kLzmaOk) return false;
            if (action == kLzmaRun && s.avail_in == 0) return true;
        }
    }
    bool write(const void *p, size_t n) override {
        if (bad) return false;
        s.next_in = static_cast<const uint8_t *>(p);
        s.avail_in = n;
        return n == 0 || pump(kLzmaRun);
    }
    bool close() override {
        bool ok = !bad;
        if (ok) {
            s.next_in = nullptr;
            s.avail_in = 0;
            ok = pump(kLzmaFinish);
        }
        lzma().end(&s);
        ok = (fclose(f) == 0) && ok;
        f = nullptr;
        return ok;
    }
};

}  // namespace

ByteSource *open_source(const char *path, Codec *codec, std::string *err) {
    FILE *f = fopen(path, "rb");
    if (!f) {
        *err = std::string("Can't open file ") + path + ": " + strerror(errno);
        return nullptr;
    }
    unsigned char m[6] = {0};
    const size_t got = fread(m, 1, sizeof m, f);
    *codec = kPlain;
    if (got >= 2 && m[0] == 0x1f && m[1] == 0x8b) *codec = kGzip;
    else if (got >= 3 && !memcmp(m, "BZh", 3)) *codec = kBzip2;
    else if (got >= 6 && !memcmp(m, "\xfd" "7zXZ\0", 6)) *codec = kXz;
    if (*codec == kPlain || *codec == kXz) {
        rewind(f);
        if (*codec == kPlain) return new FileSource(f);
        if (!lzma().ok) {
            fclose(f);
            *err = std::string(path) + " is xz-compressed and liblzma.so.5 is not available on this machine";
            return nullptr;
        }
        return new XzSource(f);
    }
    fclose(f);
    if (*codec == kGzip) {
        gzFile g = gzopen(path, "rb");
        if (!g) {
            *err = std::string("Can't open file ") + path + ": " + strerror(errno);
            return nullptr;
        }
        return new GzSource(g);
    }
    if (!bz2().ok) {
        *err = std::string(path) + " is bzip2-compressed and libbz2.so.1.0 is not available on this machine";
        return nullptr;
    }
    void *b = bz2().open(path, "rb");
    if (!b) {
        *err = std::string("Can't open file ") + path;
        return nullptr;
    }
    return new Bz2Source(b);
}

ByteSink *open_sink(const char *path, Codec codec, std::string *err) {  // util.rs:84: compression level 1
    auto cant = [&]() -> ByteSink * {
        *err = std::string("Can't create file ") + path + ": " + strerror(errno);
        return nullptr;
    };
    if (codec == kGzip) {
        gzFile g = gzopen(path, "wb1");
        return g ? static_cast<ByteSink *>(new GzSink(g)) : cant();
    }
    if (codec == kBzip2) {
        if (!bz2().ok) {
            *err = "libbz2.so.1.0 is not available on this machine";
            return nullptr;
        }
        void *b = bz2().open(path, "wb1");
        return b ? static_cast<ByteSink *>(new Bz2Sink(b)) : cant();
    }
    FILE *f = fopen(path, "wb");
    if (!f) return cant();
    if (codec == kXz) {
        if (!lzma().ok) {
            fclose(f);
            *err = "liblzma.so.5 is not available on this machine";
            return nullptr;
        }
        return new XzSink(f);
    }
    return new FileSink(f);
}

}  // namespace yb
