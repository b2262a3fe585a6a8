/* A compiled caller of the C ABI (include/yacrd_b200.h), the way a cgo / Rust-FFI / C++ host would bind it: no Python,
 * no ctypes signature table in between. Drives the producer side (Reads2Ovl::add_overlap_and_length,
 * reads2ovl/mod.rs:157), FromOverlap::compute_all_bad_part (stack.rs:143-162), BadPart::get_bad_part (stack.rs:164-169)
 * and the report writer (main.rs:80-84) on the reference's own known-answer tests (stack.rs:312-390) and checks the
 * answers. Built and run by tests/test_abi_compiled.py (-m gpu):
 *     gcc -std=c11 -Wall -Wextra -Werror -I include tests/abi_smoke.c -L yacrd_b200 -lyacrd_b200 -o abi_smoke
 * Exit code 0 = all good; every failure prints a line. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "yacrd_b200.h"

static int failures = 0;
#define CHECK(cond, ...)                         \
    do {                                         \
        if (!(cond)) {                           \
            fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
            fprintf(stderr, __VA_ARGS__);        \
            fprintf(stderr, "\n");               \
            ++failures;                          \
        }                                        \
    } while (0)

typedef struct {
    const char *id;
    uint32_t iv[6][2];
    int n_iv;
    uint32_t want[3][2];
    int n_want;
} kat;

static void add(yb_ctx *c, const char *id, uint32_t b, uint32_t e, uint64_t len) {
    int rc = yb_add_overlap_and_length(c, id, strlen(id), b, e, len);
    CHECK(rc == YB_OK, "yb_add_overlap_and_length(%s) = %d: %s", id, rc, yb_last_error(c));
}

static void expect(yb_ctx *c, const char *id, const uint32_t (*want)[2], int n_want, uint64_t want_len) {
    const uint32_t *gaps = NULL;
    uint32_t n = 12345;
    uint64_t len = 999;
    uint8_t cls = 77;
    int rc = yb_get_bad_part(c, id, strlen(id), &gaps, &n, &len, &cls);
    CHECK(rc == YB_OK, "yb_get_bad_part(%s) = %d: %s", id, rc, yb_last_error(c));
    CHECK((int)n == n_want, "%s: %u bad regions, want %d", id, n, n_want);
    CHECK(len == want_len, "%s: length %llu, want %llu", id, (unsigned long long)len, (unsigned long long)want_len);
    for (int i = 0; i < n_want && i < (int)n; ++i)
        CHECK(gaps[2 * i] == want[i][0] && gaps[2 * i + 1] == want[i][1], "%s: region %d = (%u,%u), want (%u,%u)", id, i,
              gaps[2 * i], gaps[2 * i + 1], want[i][0], want[i][1]);
}

int main(int argc, char **argv) {
    const char *report_path = argc > 1 ? argv[1] : "abi_smoke.yacrd";
    CHECK(strncmp(yb_version(), "1.0.0 Magby", 11) == 0, "version %s", yb_version());

    /* stack.rs:312-369: coverage 0 */
    static const kat kats[] = {
        {"A", {{10, 990}}, 1, {{0, 10}, {990, 1000}}, 2},
        {"B", {{10, 90}}, 1, {{0, 10}, {90, 1000}}, 2},
        {"C", {{10, 490}, {510, 990}}, 2, {{0, 10}, {490, 510}, {990, 1000}}, 3},
        {"D", {{0, 990}}, 1, {{990, 1000}}, 1},
        {"E", {{10, 1000}}, 1, {{0, 10}}, 1},
        {"F", {{0, 490}, {510, 1000}}, 2, {{490, 510}}, 1},
    };
    yb_opts opts;
    memset(&opts, 0, sizeof opts);
    opts.device = -1;
    yb_ctx *c = yb_create(&opts);
    if (!c) {
        fprintf(stderr, "yb_create failed: %s\n", yb_create_error());
        return 2;
    }
    for (size_t q = 0; q < sizeof kats / sizeof kats[0]; ++q)
        for (int i = 0; i < kats[q].n_iv; ++i) add(c, kats[q].id, kats[q].iv[i][0], kats[q].iv[i][1], 1000);
    CHECK(yb_n_reads(c) == 6, "n_reads = %u", yb_n_reads(c));
    int rc = yb_compute_all_bad_part(c, 0, 0.8);
    CHECK(rc == YB_OK, "yb_compute_all_bad_part = %d: %s", rc, yb_last_error(c));
    for (size_t q = 0; q < sizeof kats / sizeof kats[0]; ++q) expect(c, kats[q].id, kats[q].want, kats[q].n_want, 1000);
    /* stack.rs:164-169: an unknown read is not an error: no regions, length 0 */
    expect(c, "nobody", NULL, 0, 0);
    /* editor/mod.rs:114-128: a..f at n = 0.8 */
    static const uint8_t want_cls[6] = {YB_NOT_BAD, YB_NOT_COVERED, YB_CHIMERIC, YB_NOT_BAD, YB_NOT_BAD, YB_CHIMERIC};
    for (uint32_t i = 0; i < 6; ++i) {
        const uint32_t *g;
        uint32_t n;
        uint64_t len;
        uint8_t cls;
        rc = yb_get_bad_part_at(c, i, &g, &n, &len, &cls);
        CHECK(rc == YB_OK && cls == want_cls[i], "read %u: class %u, want %u", i, cls, want_cls[i]);
        const char *id;
        size_t id_len;
        CHECK(yb_read_at(c, i, &id, &id_len) == YB_OK && id_len == 1 && id[0] == kats[i].id[0], "read_at(%u)", i);
    }
    rc = yb_write_report(c, report_path);
    CHECK(rc == YB_OK, "yb_write_report = %d: %s", rc, yb_last_error(c));
    {
        FILE *fh = fopen(report_path, "r");
        CHECK(fh != NULL, "report not written");
        if (fh) {
            char line[256];
            CHECK(fgets(line, sizeof line, fh) != NULL && strcmp(line, "NotBad\tA\t1000\t10,0,10;10,990,1000\n") == 0, "first report line: %s", line);
            fclose(fh);
        }
    }
    yb_stats st;
    CHECK(yb_get_stats(c, &st) == YB_OK && st.n_reads == 6 && st.n_intervals == 8 && st.kernel_launches > 0, "stats");

    /* stack.rs:372-390: coverage 2, next batch on the same context (stack.rs:148-161 loops over batches) */
    CHECK(yb_reset(c) == YB_OK, "yb_reset");
    static const uint32_t cov2[6][2] = {{0, 425}, {0, 450}, {0, 475}, {525, 1000}, {550, 1000}, {575, 1000}};
    for (int i = 0; i < 6; ++i) add(c, "A", cov2[i][0], cov2[i][1], 1000);
    rc = yb_compute_all_bad_part(c, 2, 0.8);
    CHECK(rc == YB_OK, "yb_compute_all_bad_part(c=2) = %d: %s", rc, yb_last_error(c));
    static const uint32_t want2[1][2] = {{425, 575}};
    expect(c, "A", want2, 1, 1000);
    yb_destroy(c);
    if (failures == 0) printf("abi_smoke ok\n");
    return failures ? 1 : 0;
}
