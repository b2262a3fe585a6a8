// synth.cpp — synthetic overlap workloads for BASELINE.json's configs. Host only, no CUDA.
// Built by workload/Makefile into workload/libyacrd_synth.so (declarations: workload/synth.h): measurement and test
// infrastructure, not part of the product library. Both bench arms and the tests load it from there.
//
// Counter-based: every value of read r is a pure function of (seed, r), so a shard can be generated
// on its own and any subset of reads reproduces bit for bit. Reads are assigned to shards by
// mix64(read index) % n_shards (the north star's "read-id hash").
//
// Per read r:
//   length   ONT:    clamp(round(exp(N(9.0, 0.75))), 200, 250000)   (median ~8.1 kb)
//            PacBio: clamp(round(exp(N(9.39, 0.55))), 500, 120000)  (median ~12 kb)
//   k        ONT:    max(1, round(Gamma(shape 2, scale mean/2)))    (over-dispersed, sd ~ mean/sqrt 2)
//            skew:   min(5000, 1 + floor(8 * Pareto(alpha 1.15)))   (heavy tail, rows at the 5 k cap)
//   kind     3 % chimeric (junction j in [0.2,0.8] len, gap w in [1,300]; no interval crosses (j, j+w)),
//            5 % sparsely covered (intervals shorter than 15 % of len), else normal:
//            40 % prefix (s, x) with s in [0,40], 40 % suffix (x, len - t) with t in [0,40], 20 % internal;
//            0.5 % of rows carry an abutting pair (a,m),(m,z) (the zero-length-gap quirk, stack.rs:73-85).
//   All intervals are well formed: 0 <= begin < end <= length. Order within a row is the draw order
//   (unsorted), as a PAF delivers it.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "synth.h"

namespace {

inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct Rng {  // splitmix64 stream keyed by (seed, read, stream)
    uint64_t s;
    Rng(uint64_t seed, uint64_t read, uint64_t stream) : s(mix64(seed ^ mix64(read * 0x9E3779B97F4A7C15ull + stream))) {}
    uint64_t next() { return mix64(s += 0x9E3779B97F4A7C15ull); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }            // [0,1)
    double uni_open() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }  // (0,1)
    uint32_t below(uint32_t n) { return n ? (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32) : 0; }
    uint32_t range(uint32_t lo, uint32_t hi) { return lo + below(hi - lo + 1); }  // inclusive
    double normal() {
        const double u1 = uni_open(), u2 = uni();
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
};

inline uint32_t shard_of(uint32_t read, uint32_t n_shards) { return n_shards <= 1 ? 0 : (uint32_t)(mix64(read) % n_shards); }

inline void plan_read(const yb_synth_spec *sp, uint32_t r, uint32_t *len, uint32_t *k) {
    Rng g(sp->seed, r, 1);
    double L;
    if (sp->profile == YB_SYNTH_PACBIO_SKEW) {
        L = exp(9.39 + 0.55 * g.normal());
        L = std::min(120000.0, std::max(500.0, floor(L + 0.5)));
        const double par = pow(g.uni_open(), -1.0 / 1.15);
        const double kk = 1.0 + floor(8.0 * par);
        *k = (uint32_t)std::min(5000.0, kk);
    } else {
        L = exp(9.0 + 0.75 * g.normal());
        L = std::min(250000.0, std::max(200.0, floor(L + 0.5)));
        const double gam = -(sp->mean_intervals * 0.5) * log(g.uni_open() * g.uni_open());
        *k = (uint32_t)std::min(100000.0, std::max(1.0, floor(gam + 0.5)));
    }
    *len = (uint32_t)L;
}

void fill_read(const yb_synth_spec *sp, uint32_t r, uint32_t len, uint32_t k, uint32_t *iv) {
    Rng g(sp->seed, r, 2);
    const double kind = g.uni();
    const bool abut = g.uni() < 0.005 && k >= 2 && len >= 64;
    uint32_t j = 0, w = 0;
    if (kind < 0.03 && len >= 400) {
        j = (uint32_t)(len * (0.2 + 0.6 * g.uni()));
        w = g.range(1, 300);
        if (j + w + 2 > len) w = 1;
    }
    const bool chim = w != 0, sparse = !chim && kind < 0.08;
    for (uint32_t i = 0; i < k; ++i) {
        uint32_t b, e;
        if (chim) {  // every interval inside [0, j] or [j + w, len]
            const bool left = g.uni() < 0.5;
            const uint32_t lo = left ? 0 : j + w, hi = left ? j : len;
            const double u = g.uni();
            if (u < 0.45) {  // hugs the junction
                if (left) { e = hi; b = lo + g.below(hi - lo); }
                else { b = lo; e = lo + 1 + g.below(hi - lo); }
            } else if (u < 0.9) {  // hugs the read end
                if (left) { b = lo + g.below(std::min(41u, hi - lo)); e = b + 1 + g.below(hi - b); }
                else { e = hi - g.below(std::min(41u, hi - lo)); b = lo + g.below(e - lo); }
            } else {
                b = lo + g.below(hi - lo);
                e = b + 1 + g.below(hi - b);
            }
        } else if (sparse) {
            const uint32_t maxl = std::max(1u, (uint32_t)(0.15 * len));
            const uint32_t l = 1 + g.below(maxl);
            b = g.below(len - l + 1);
            e = b + l;
        } else {
            const double u = g.uni();
            if (u < 0.4) {  // prefix overlap
                b = g.below(std::min(41u, len));
                e = b + 1 + g.below(len - b);
            } else if (u < 0.8) {  // suffix overlap
                e = len - g.below(std::min(41u, len));
                b = g.below(e);
            } else {  // contained / internal
                b = g.below(len);
                e = b + 1 + g.below(len - b);
            }
        }
        iv[2 * i] = b;
        iv[2 * i + 1] = e;
    }
    if (abut) {  // overwrite the first two intervals with an abutting pair
        const uint32_t m = len / 4 + g.below(len / 2);
        const uint32_t a = g.below(m), z = m + 1 + g.below(len - m);
        iv[0] = a;
        iv[1] = m;
        iv[2] = m;
        iv[3] = z;
    }
}

}  // namespace

extern "C" {

uint32_t yb_synth_shard_of(uint32_t read, uint32_t n_shards) { return shard_of(read, n_shards); }

uint32_t yb_synth_count(const yb_synth_spec *sp) {
    if (!sp) return 0;
    if (sp->n_shards <= 1) return sp->n_reads;
    uint32_t n = 0;
    for (uint32_t r = 0; r < sp->n_reads; ++r) n += shard_of(r, sp->n_shards) == sp->shard;
    return n;
}

uint64_t yb_synth_plan(const yb_synth_spec *sp, uint32_t *global_idx, uint32_t *rowptr, uint32_t *length) {
    if (!sp || !rowptr || !length) return 0;
    uint64_t tot = 0;
    uint32_t i = 0;
    for (uint32_t r = 0; r < sp->n_reads; ++r) {
        if (sp->n_shards > 1 && shard_of(r, sp->n_shards) != sp->shard) continue;
        uint32_t len, k;
        plan_read(sp, r, &len, &k);
        if (global_idx) global_idx[i] = r;
        rowptr[i] = (uint32_t)tot;
        length[i] = len;
        tot += k;
        ++i;
    }
    rowptr[i] = (uint32_t)tot;
    return tot;
}

int yb_synth_fill(const yb_synth_spec *sp, const uint32_t *global_idx, const uint32_t *rowptr, const uint32_t *length,
                  uint32_t n_local, uint32_t *iv, int threads) {
    if (!sp || !rowptr || !length || !iv) return -9;  // (invalid argument)
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<uint32_t>((uint32_t)threads, std::max(1u, n_local));
    auto work = [&](uint32_t r0, uint32_t r1) {
        for (uint32_t i = r0; i < r1; ++i)
            fill_read(sp, global_idx ? global_idx[i] : i, length[i], rowptr[i + 1] - rowptr[i], iv + 2 * (size_t)rowptr[i]);
    };
    std::vector<std::thread> pool;
    const uint64_t total = rowptr[n_local];
    uint32_t r = 0;
    for (int t = 0; t < threads; ++t) {
        const uint64_t target = total * (uint64_t)(t + 1) / (uint64_t)threads;
        const uint32_t r0 = r;
        while (r < n_local && (t == threads - 1 || rowptr[r + 1] <= target)) ++r;
        if (t == threads - 1) r = n_local;
        if (r > r0) pool.emplace_back(work, r0, r);
    }
    for (auto &th : pool) th.join();
    return 0;
}

// Synthetic PAF text for the ingestion bench: n_records overlap records between pseudo-random pairs of
// n_reads reads ("read_0000123", per-read length fixed by the read index), 12 columns like minimap2's.
// Returns the bytes needed; writes only when cap is large enough.
uint64_t yb_synth_paf(uint64_t seed, uint32_t n_reads, uint64_t n_records, char *out, uint64_t cap) {
    const uint64_t need = n_records * 96ull + 16;
    if (!out || cap < need || n_reads == 0) return need;
    char *p = out;
    auto put = [&](uint64_t v) {
        char tmp[24];
        int n = 0;
        do {
            tmp[n++] = (char)('0' + v % 10);
            v /= 10;
        } while (v);
        while (n) *p++ = tmp[--n];
    };
    auto put_id = [&](uint32_t r) {
        memcpy(p, "read_", 5);
        p += 5;
        char tmp[8];
        for (int i = 6; i >= 0; --i) {
            tmp[i] = (char)('0' + r % 10);
            r /= 10;
        }
        memcpy(p, tmp, 7);
        p += 7;
    };
    for (uint64_t i = 0; i < n_records; ++i) {
        Rng g(seed, i, 77);
        const uint32_t a = (uint32_t)(g.next() % n_reads), b = (uint32_t)(g.next() % n_reads);
        const uint32_t la = 500 + (uint32_t)(mix64(seed ^ (a * 0x9E3779B97F4A7C15ull)) % 60000);
        const uint32_t lb = 500 + (uint32_t)(mix64(seed ^ (b * 0x9E3779B97F4A7C15ull)) % 60000);
        const uint32_t ba = (uint32_t)(g.next() % (la - 1)), ea = ba + 1 + (uint32_t)(g.next() % (la - ba));
        const uint32_t bb = (uint32_t)(g.next() % (lb - 1)), eb = bb + 1 + (uint32_t)(g.next() % (lb - bb));
        put_id(a); *p++ = '\t'; put(la); *p++ = '\t'; put(ba); *p++ = '\t'; put(ea); *p++ = '\t';
        *p++ = (i & 1) ? '-' : '+'; *p++ = '\t';
        put_id(b); *p++ = '\t'; put(lb); *p++ = '\t'; put(bb); *p++ = '\t'; put(eb); *p++ = '\t';
        put(ea - ba); *p++ = '\t'; put(ea - ba); *p++ = '\t';
        memcpy(p, "255\n", 4);
        p += 4;
    }
    return (uint64_t)(p - out);
}

}  // extern "C"
