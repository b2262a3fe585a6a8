/* synth.h - the synthetic workload generator of BASELINE.json's configs: measurement and test
 * infrastructure, host only, built into workload/libyacrd_synth.so. Not part of the product library. */
#ifndef YACRD_SYNTH_H
#define YACRD_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic workload generator (BASELINE.json configs). Host only. ---------- */
typedef struct yb_synth_spec {
    uint64_t seed;          /* 20261017 */
    uint32_t n_reads;       /* GLOBAL number of reads of the workload */
    uint32_t shard;         /* this shard: reads with yb_synth_shard_of(read, n_shards) == shard */
    uint32_t n_shards;      /* 0 or 1: no sharding */
    uint32_t profile;       /* YB_SYNTH_* */
    double mean_intervals;  /* per read (ignored by the skewed profile) */
} yb_synth_spec;
#define YB_SYNTH_ONT 0u          /* ONT lengths, over-dispersed interval counts */
#define YB_SYNTH_PACBIO_SKEW 1u  /* PacBio Sequel lengths, Pareto interval counts capped at 5000 */
uint32_t yb_synth_shard_of(uint32_t read, uint32_t n_shards); /* read-id hash sharding */
uint32_t yb_synth_count(const yb_synth_spec *spec);           /* reads in this shard */
/* Pass 1: global_idx[0..n_local) (may be NULL), rowptr[0..n_local] and length[0..n_local)
 * (caller-allocated, n_local = yb_synth_count). Returns the shard's total intervals. */
uint64_t yb_synth_plan(const yb_synth_spec *spec, uint32_t *global_idx, uint32_t *rowptr,
                       uint32_t *length);
/* Pass 2: fill iv (pairs) for the rows planned by pass 1. threads <= 0: all cores. */
int yb_synth_fill(const yb_synth_spec *spec, const uint32_t *global_idx, const uint32_t *rowptr,
                  const uint32_t *length, uint32_t n_local, uint32_t *iv, int threads);

/* Synthetic PAF text for the ingestion bench (tools/bench_ingest.py): n_records records between pseudo-random pairs
 * of n_reads reads. Returns the bytes written, or the bytes needed when out is NULL / cap is too small. */
uint64_t yb_synth_paf(uint64_t seed, uint32_t n_reads, uint64_t n_records, char *out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
