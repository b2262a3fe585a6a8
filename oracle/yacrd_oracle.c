/*
 * yacrd_oracle.c — CPU ORACLE (test infrastructure, NOT the product).
 *
 * A literal, line-by-line restatement in plain C of the reference's detect hot path
 * (natir/yacrd 1.0.0 "Magby", Rust):
 *
 *   yo_compute_bad_part  <- src/stack.rs:61-139   (FromOverlap::compute_bad_part)
 *   yo_type_of_read      <- src/editor/mod.rs:85-100
 *   yo_format_line       <- src/editor/mod.rs:61-83,102-107 (report + bad_region_format)
 *   yo_run_csr           <- src/stack.rs:143-162  (compute_all_bad_part batch loop; the rayon
 *                           par_bridge is restated as a static pthread partition over reads)
 *
 * It deliberately keeps the reference's algorithm (lexicographic sort + min-heap sweep of interval
 * ends), so that it is an independent check of the device's closed-form crossing algorithm.
 *
 * PARITY PIN: the Rust reference cannot be compiled in this image (no cargo/rustc), so there is no
 * oracle/_ref. This restatement is pinned against the reference's own vectors instead:
 * stack.rs:312-390 KATs, editor/mod.rs:114-128 KATs and tests/reads.paf -> tests/truth.yacrd
 * (see tests/test_oracle.py, tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library. The product (yacrd_b200/) never does.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define YO_NOT_BAD 0
#define YO_CHIMERIC 1
#define YO_NOT_COVERED 2

/* ---- min-heap of u32 (BinaryHeap<Reverse<u32>>, stack.rs:63-64) ------------------------------ */
typedef struct {
    uint32_t *a;
    size_t n;
} yo_heap;

static void heap_push(yo_heap *h, uint32_t v) {
    size_t i = h->n++;
    h->a[i] = v;
    while (i > 0) {
        size_t p = (i - 1) / 2;
        if (h->a[p] <= h->a[i]) break;
        uint32_t t = h->a[p];
        h->a[p] = h->a[i];
        h->a[i] = t;
        i = p;
    }
}

static void heap_pop(yo_heap *h) {
    h->a[0] = h->a[--h->n];
    size_t i = 0;
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < h->n && h->a[l] < h->a[m]) m = l;
        if (r < h->n && h->a[r] < h->a[m]) m = r;
        if (m == i) break;
        uint32_t t = h->a[m];
        h->a[m] = h->a[i];
        h->a[i] = t;
        i = m;
    }
}

/* lexicographic order on (u32,u32) tuples: ovls.sort_unstable() (stack.rs:66) */
static int cmp_pair(const void *pa, const void *pb) {
    const uint32_t *a = (const uint32_t *)pa, *b = (const uint32_t *)pb;
    if (a[0] != b[0]) return a[0] < b[0] ? -1 : 1;
    if (a[1] != b[1]) return a[1] < b[1] ? -1 : 1;
    return 0;
}

/*
 * stack.rs:61-139. `iv` holds k (begin,end) pairs; `gaps` must have room for k+2 pairs;
 * `scratch` must have room for 3*k+4 u32 (sorted copy + heap). Returns the number of gaps written.
 */
static uint32_t compute_bad_part(const uint32_t *iv, uint32_t k, uint64_t len, uint64_t coverage,
                                 uint32_t *gaps, uint32_t *scratch) {
    uint32_t *ovls = scratch;               /* 2*k */
    yo_heap stack = {scratch + 2 * (size_t)k, 0}; /* k */
    uint32_t *raw = gaps;                   /* raw gaps built in place, cleaned afterwards */
    size_t n_raw = 0;

    memcpy(ovls, iv, sizeof(uint32_t) * 2 * (size_t)k);
    qsort(ovls, k, 2 * sizeof(uint32_t), cmp_pair); /* stack.rs:66 */

    uint32_t first_covered = 0; /* stack.rs:68 */
    uint32_t last_covered = 0;  /* stack.rs:69 */

    for (uint32_t i = 0; i < k; ++i) { /* stack.rs:71 */
        uint32_t b = ovls[2 * i], e = ovls[2 * i + 1];
        while (stack.n > 0) { /* stack.rs:72 */
            uint32_t head = stack.a[0];
            if (head > b) break;                       /* stack.rs:73-75 */
            if (stack.n > coverage) last_covered = head; /* stack.rs:77-79 */
            heap_pop(&stack);                          /* stack.rs:80 */
        }
        if (stack.n <= coverage) { /* stack.rs:83 */
            if (last_covered != 0) { /* stack.rs:84-85 */
                raw[2 * n_raw] = last_covered;
                raw[2 * n_raw + 1] = b;
                ++n_raw;
            } else {
                first_covered = b; /* stack.rs:87 */
            }
        }
        heap_push(&stack, e); /* stack.rs:90 */
    }

    while (stack.n > coverage) { /* stack.rs:93 */
        last_covered = stack.a[0]; /* stack.rs:94-100 */
        if ((uint64_t)last_covered >= len) break; /* stack.rs:101-103 */
        heap_pop(&stack);                          /* stack.rs:104 */
    }

    if (first_covered != 0) { /* stack.rs:107-109: gaps.insert(0, (0, first_covered)) */
        memmove(raw + 2, raw, sizeof(uint32_t) * 2 * n_raw);
        raw[0] = 0;
        raw[1] = first_covered;
        ++n_raw;
    }
    if ((uint64_t)last_covered != len) { /* stack.rs:111-113 */
        raw[2 * n_raw] = last_covered;
        raw[2 * n_raw + 1] = (uint32_t)len; /* `len as u32` */
        ++n_raw;
    }
    if (n_raw == 0) return 0; /* stack.rs:115-117 */

    /* clean overlapped bad region, stack.rs:119-138 — in place: the write index never passes i */
    size_t n_clean = 0;
    uint32_t begin = raw[0], end = raw[1];
    for (size_t i = 0; i + 1 < n_raw; ++i) { /* gaps.windows(2) */
        uint32_t g1b = raw[2 * i], g1e = raw[2 * i + 1];
        uint32_t g2b = raw[2 * i + 2], g2e = raw[2 * i + 3];
        if (g1b == g2b) { /* stack.rs:127-129 */
            begin = g1b;
            end = g1e > g2e ? g1e : g2e;
        } else { /* stack.rs:130-134 */
            raw[2 * n_clean] = begin;
            raw[2 * n_clean + 1] = end;
            ++n_clean;
            begin = g2b;
            end = g2e;
        }
    }
    raw[2 * n_clean] = begin; /* stack.rs:136 */
    raw[2 * n_clean + 1] = end;
    ++n_clean;
    return (uint32_t)n_clean;
}

uint32_t yo_compute_bad_part(const uint32_t *iv, uint32_t k, uint64_t len, uint64_t coverage,
                             uint32_t *gaps) {
    uint32_t *scratch = (uint32_t *)malloc(sizeof(uint32_t) * (3 * (size_t)k + 4));
    uint32_t n = compute_bad_part(iv, k, len, coverage, gaps, scratch);
    free(scratch);
    return n;
}

/* editor/mod.rs:85-100. u32 wrapping sum (Cargo.toml:40 overflow-checks=false), f64 ratio,
 * NotCovered tested first, strict '>'. */
int yo_type_of_read(uint64_t length, const uint32_t *bads, uint32_t n_bads, double not_covered) {
    uint32_t bad_region_len = 0;
    for (uint32_t i = 0; i < n_bads; ++i) bad_region_len += bads[2 * i + 1] - bads[2 * i];
    if ((double)bad_region_len / (double)length > not_covered) return YO_NOT_COVERED;
    for (uint32_t i = 0; i < n_bads; ++i)
        if (bads[2 * i] != 0 && bads[2 * i + 1] != (uint32_t)length) return YO_CHIMERIC;
    return YO_NOT_BAD;
}

static const char *const yo_names[3] = {"NotBad", "Chimeric", "NotCovered"}; /* editor/mod.rs:51-58 */

const char *yo_type_name(int t) { return (t >= 0 && t < 3) ? yo_names[t] : "?"; }

/* editor/mod.rs:72-79 + 102-107: "{type}\t{id}\t{len}\t{len,beg,end;...}\n". Returns bytes written
 * (excluding NUL) or -1 if `cap` is too small. */
long yo_format_line(const char *id, uint64_t length, const uint32_t *bads, uint32_t n_bads,
                    double not_covered, char *out, size_t cap) {
    int t = yo_type_of_read(length, bads, n_bads, not_covered);
    size_t w = 0;
    int r = snprintf(out, cap, "%s\t%s\t%llu\t", yo_names[t], id, (unsigned long long)length);
    if (r < 0 || (size_t)r >= cap) return -1;
    w = (size_t)r;
    for (uint32_t i = 0; i < n_bads; ++i) {
        uint32_t b = bads[2 * i], e = bads[2 * i + 1];
        r = snprintf(out + w, cap - w, "%s%u,%u,%u", i ? ";" : "", (uint32_t)(e - b), b, e);
        if (r < 0 || (size_t)r >= cap - w) return -1;
        w += (size_t)r;
    }
    if (w + 2 > cap) return -1;
    out[w++] = '\n';
    out[w] = 0;
    return (long)w;
}

/* ---- batch driver (stack.rs:143-162) -------------------------------------------------------- */
typedef struct {
    const uint64_t *rowptr;
    const uint32_t *iv;
    const uint32_t *len;
    uint32_t r0, r1, max_k;
    uint64_t coverage;
    double not_covered;
    uint8_t *cls;
    uint32_t *cnt;
    uint32_t *padded;
    uint64_t total;
} yo_job;

static void *yo_worker(void *p) {
    yo_job *j = (yo_job *)p;
    uint32_t *scratch = (uint32_t *)malloc(sizeof(uint32_t) * (3 * (size_t)j->max_k + 4));
    uint64_t tot = 0;
    for (uint32_t r = j->r0; r < j->r1; ++r) {
        uint64_t s = j->rowptr[r];
        uint32_t k = (uint32_t)(j->rowptr[r + 1] - s);
        /* padded slots: read r owns pairs [rowptr[r] + 2r, rowptr[r+1] + 2(r+1)) (at most k+2 gaps) */
        uint32_t *g = j->padded + 2 * (s + 2 * (uint64_t)r);
        uint32_t n = compute_bad_part(j->iv + 2 * s, k, j->len[r], j->coverage, g, scratch);
        j->cnt[r] = n;
        j->cls[r] = (uint8_t)yo_type_of_read(j->len[r], g, n, j->not_covered);
        tot += n;
    }
    free(scratch);
    j->total = tot;
    return NULL;
}

int yo_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/*
 * The timed CPU-baseline region: compute_bad_part + type_of_read for every read (stack.rs:154 +
 * editor/mod.rs:71), reads statically partitioned over `threads` pthreads by interval count.
 * `padded` must hold rowptr[n_reads] + 2*n_reads pairs; read r's gaps land at pair index
 * rowptr[r] + 2r, gap_cnt[r] of them. threads <= 0: all online cores. Returns total gaps.
 */
uint64_t yo_run_csr_padded(const uint64_t *rowptr, const uint32_t *iv, const uint32_t *len,
                           uint32_t n_reads, uint64_t coverage, double not_covered, uint8_t *cls,
                           uint32_t *gap_cnt, uint32_t *padded, int threads) {
    if (threads <= 0) threads = yo_max_threads();
    if ((uint32_t)threads > n_reads) threads = n_reads ? (int)n_reads : 1;
    uint32_t max_k = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        uint64_t k = rowptr[r + 1] - rowptr[r];
        if (k > max_k) max_k = (uint32_t)k;
    }
    yo_job *jobs = (yo_job *)calloc((size_t)threads, sizeof(yo_job));
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    /* balance by work ~ intervals + reads */
    uint64_t work_total = rowptr[n_reads] + n_reads;
    uint32_t r = 0;
    for (int t = 0; t < threads; ++t) {
        uint64_t target = work_total * (uint64_t)(t + 1) / (uint64_t)threads;
        uint32_t r0 = r;
        while (r < n_reads && (t == threads - 1 || rowptr[r + 1] + r + 1 <= target)) ++r;
        yo_job j = {rowptr, iv, len, r0, r, max_k, coverage, not_covered, cls, gap_cnt, padded, 0};
        jobs[t] = j;
    }
    for (int t = 1; t < threads; ++t) pthread_create(&tid[t], NULL, yo_worker, &jobs[t]);
    yo_worker(&jobs[0]);
    uint64_t tot = jobs[0].total;
    for (int t = 1; t < threads; ++t) {
        pthread_join(tid[t], NULL);
        tot += jobs[t].total;
    }
    free(jobs);
    free(tid);
    return tot;
}

/*
 * Same, then compacted to a CSR of gaps: gap_ptr[0..n_reads] (exclusive scan of the counts) and
 * `gaps` (pairs; capacity rowptr[n_reads] + 2*n_reads pairs always suffices). Returns total gaps.
 */
uint64_t yo_run_csr(const uint64_t *rowptr, const uint32_t *iv, const uint32_t *len,
                    uint32_t n_reads, uint64_t coverage, double not_covered, uint8_t *cls,
                    uint64_t *gap_ptr, uint32_t *gaps, int threads) {
    uint64_t cap_pairs = rowptr[n_reads] + 2 * (uint64_t)n_reads;
    uint32_t *cnt = (uint32_t *)calloc((size_t)n_reads + 1, sizeof(uint32_t));
    uint32_t *padded = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (cap_pairs ? cap_pairs : 1));
    yo_run_csr_padded(rowptr, iv, len, n_reads, coverage, not_covered, cls, cnt, padded, threads);
    uint64_t tot = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        gap_ptr[r] = tot;
        memcpy(gaps + 2 * tot, padded + 2 * (rowptr[r] + 2 * (uint64_t)r),
               sizeof(uint32_t) * 2 * cnt[r]);
        tot += cnt[r];
    }
    gap_ptr[n_reads] = tot;
    free(padded);
    free(cnt);
    return tot;
}
