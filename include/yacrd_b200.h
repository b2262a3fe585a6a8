/*
 * yacrd_b200.h — C ABI of the B200-native detect path of yacrd (coverage pile-up, bad-region
 * segmentation, Chimeric / NotCovered / NotBad classification).
 *
 * The reference (natir/yacrd 1.0.0, Rust) has no FFI; its seam for this path is two trait objects,
 * `Box<dyn Reads2Ovl>` (src/reads2ovl/mod.rs:43-163) and `Box<dyn BadPart>` (src/stack.rs:35-41),
 * driven by src/main.rs:42-84. Every entry point below names the trait item it replaces, so that a
 * ~60-line `impl Reads2Ovl` + `impl BadPart` shim is all a maintainer adds.
 *
 * Conventions
 *  - plain pointers and sizes; no C++ or torch types cross this boundary; nothing throws.
 *  - every function that can fail returns an int: YB_OK (0) or a negative yb_status mirroring
 *    src/error.rs:30-92; yb_last_error(ctx) gives the message.
 *  - a yb_ctx is single-threaded (the traits take &mut self); use one ctx per GPU / per thread.
 *  - ids are byte strings (not NUL-terminated), copied on first sight; reads are indexed densely
 *    in first-seen order.
 *  - output views (gaps, ids, bitmaps) are BORROWED: valid until the next mutating call on the ctx
 *    or yb_destroy — the same lifetime `get_bad_part` gives its `&(Vec<(u32,u32)>, usize)`.
 *  - the compute path is CUDA (sm_100a) only. There is no CPU fallback: without a usable device
 *    yb_create fails with YB_ERR_CUDA.
 */
#ifndef YACRD_B200_H
#define YACRD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YB_VERSION "1.0.0 Magby (b200)" /* src/cli.rs:35 version string + backend tag */

typedef enum yb_status {
    YB_OK = 0,
    YB_ERR_CANT_READ_FILE = -1,        /* error.rs: CantReadFile */
    YB_ERR_CANT_WRITE_FILE = -2,       /* error.rs: CantWriteFile */
    YB_ERR_UNKNOWN_FORMAT = -3,        /* error.rs: UnableToDetectFileFormat */
    YB_ERR_WRONG_FORMAT = -4,          /* error.rs: CantRunOperationOnFile (fasta/fastq/yacrd given as overlaps) */
    YB_ERR_READING = -5,               /* error.rs: ReadingErrorNoFilename (record does not deserialize) */
    YB_ERR_WRITING = -6,               /* error.rs: WritingErrorNoFilename */
    YB_ERR_CORRUPT_REPORT = -7,        /* error.rs: CorruptYacrdReport */
    YB_ERR_MALFORMED_INTERVAL = -8,    /* reserved (round 1 rejected begin >= end / end > length; such rows are now computed
                                          with the reference's own heap sweep, see yb_stats.n_malformed_intervals) */
    YB_ERR_INVALID_ARGUMENT = -9,
    YB_ERR_STATE = -10,                /* call order violated (e.g. results queried before compute) */
    YB_ERR_TOO_LARGE = -11,            /* > 2^32-16 intervals or reads in one context, a read longer than 2^31-1, or (yb_upload) a
                                          one-shot batch of more than about 2.8e9 intervals + reads: use yb_set_chunk_intervals */
    YB_ERR_CUDA = -12,
    YB_ERR_NOMEM = -13
} yb_status;

/* src/editor/mod.rs:43-59 ReadType; also the 2-bit code in the class bitmap */
typedef enum yb_read_type { YB_NOT_BAD = 0, YB_CHIMERIC = 1, YB_NOT_COVERED = 2 } yb_read_type;

typedef struct yb_ctx yb_ctx;

typedef struct yb_opts {
    int32_t device;            /* CUDA ordinal; -1 = current device */
    uint32_t read_buffer_size; /* --read-buffer-size (cli.rs:57-59), used by yb_init_file; 0 = 8192 */
    uint32_t flags;            /* YB_FLAG_* */
    uint32_t ingest_threads;   /* host threads of yb_init_file / yb_init_buffer; 0 = all cores (at most 32) */
} yb_opts;

#define YB_FLAG_KEEP_HOST_INTERVALS 1u /* keep arrival-order intervals so yb_overlap() works after upload */
#define YB_FLAG_LAZY_DEVICE 4u /* yb_create does not touch CUDA: the device (context creation, about half a second) is opened
                                by the first call that needs it (yb_upload, yb_compute_*, yb_peer_*, yb_init_report*). With
                                yb_device_warmup on another thread, a driver parses its input while CUDA starts up. */
#define YB_FLAG_HOST_ONLY 2u /* producer side only (ingestion, interning, CSR freeze, Reads2Ovl queries): no device is
                                touched; yb_upload / yb_compute_* fail with YB_ERR_CUDA. There is still no CPU pile-up. */

typedef struct yb_stats {
    uint64_t n_reads;
    uint64_t n_intervals;
    uint64_t n_gaps;          /* after compute */
    uint64_t n_not_bad, n_chimeric, n_not_covered;
    uint64_t max_intervals_per_read;
    uint64_t n_reads_warp, n_reads_cta, n_reads_huge; /* reserved (always 0) */
    uint64_t kernel_launches; /* cumulative count of this library's kernel launches */
    uint64_t h2d_bytes, d2h_bytes; /* cumulative */
    uint64_t n_malformed_intervals; /* intervals with begin >= end or end > length in the last batch: accepted, like the
                                       reference does (stack.rs:61-139 has no such test) */
    uint64_t n_literal_reads;       /* reads holding one: computed by the literal heap sweep kernel, not the closed form */
} yb_stats;

/* ---- lifecycle ------------------------------------------------------------------------------ */
/* FullMemory::new (fullmemory.rs:36) + FromOverlap::new (stack.rs:51-59). NULL on failure
 * (yb_create_error() tells why). */
yb_ctx *yb_create(const yb_opts *opts);
const char *yb_create_error(void);
void yb_destroy(yb_ctx *ctx);
/* Empties the store and the results but keeps all host/device buffers, ready for the next batch
 * (the reference's compute_all_bad_part loops over get_overlaps batches, stack.rs:148-161). */
int yb_reset(yb_ctx *ctx);
const char *yb_last_error(const yb_ctx *ctx);
/* First CUDA call of the process for `device` (< 0: the current one): safe from any thread, any number of times. */
int yb_device_warmup(int device);
const char *yb_version(void);
const char *yb_type_name(int read_type); /* ReadType::as_str, editor/mod.rs:51-58 */

/* ---- producer side: trait Reads2Ovl (reads2ovl/mod.rs:43-163) ------------------------------- */
/* add_overlap_and_length, mod.rs:157 / fullmemory.rs:82-90: first-seen length wins. */
int yb_add_overlap_and_length(yb_ctx *ctx, const char *id, size_t id_len, uint32_t begin,
                              uint32_t end, uint64_t length);
/* add_overlap, mod.rs:154 / fullmemory.rs:64-72 (length stays 0 until add_length). */
int yb_add_overlap(yb_ctx *ctx, const char *id, size_t id_len, uint32_t begin, uint32_t end);
/* add_length, mod.rs:155 / fullmemory.rs:74-76: sets unconditionally. */
int yb_add_length(yb_ctx *ctx, const char *id, size_t id_len, uint64_t length);
/* Bulk variant for pre-interned input (one get_overlaps batch, stack.rs:149): read r owns
 * iv[2*rowptr[r] .. 2*rowptr[r+1]) as (begin,end) pairs. With ids (ids[r] / id_lens[r], id_lens may be
 * NULL for NUL-terminated ids) the rows are merged into the named store like add_overlap_and_length.
 * With ids == NULL the reads are named by their decimal index and the block is appended to the
 * context's pinned CSR directly (an index-named context cannot also hold named reads). Host pointers;
 * the data is copied. */
int yb_add_csr(yb_ctx *ctx, const uint32_t *rowptr, const uint32_t *iv, const uint32_t *length,
               uint32_t n_reads, const char *const *ids, const size_t *id_lens);
/* Zero-copy variant: the context BORROWS the caller's host CSR (index-named reads, rowptr[0] == 0) until
 * yb_destroy; allocate it with yb_host_alloc for full H2D speed. Needs an empty context. */
int yb_bind_csr(yb_ctx *ctx, const uint32_t *rowptr, const uint32_t *iv, const uint32_t *length,
                uint32_t n_reads);
/* Page-locked host memory for buffers handed to yb_bind_csr. */
void *yb_host_alloc(size_t n_bytes);
void yb_host_free(void *p);
/* Reads2Ovl::init, mod.rs:44-81: sniff PAF/M4 by file name (util.rs:39-55) and parse. */
int yb_init_file(yb_ctx *ctx, const char *path);
/* Same over an in-memory buffer; format: 'p' = PAF (tab, io.rs:24-34), 'm' = M4 (space, io.rs:37-50). */
int yb_init_buffer(yb_ctx *ctx, const char *text, size_t n_bytes, int format);
/* util::get_file_type (util.rs:39-55): 'm' m4/mhap, 'p' paf, 'y' yacrd, 'q' fastq, 'a' fasta, 'o' yovl,
 * 0 unknown. */
int yb_file_type(const char *path);
/* Reads2Ovl::length, mod.rs:152 (0 when unknown). */
uint64_t yb_length(const yb_ctx *ctx, const char *id, size_t id_len);
/* Reads2Ovl::overlap, mod.rs:150: arrival-order intervals; empty when unknown. Needs
 * YB_FLAG_KEEP_HOST_INTERVALS once the data has been uploaded. */
int yb_overlap(yb_ctx *ctx, const char *id, size_t id_len, const uint32_t **iv_pairs,
               uint32_t *n_intervals);
/* get_reads (mod.rs:160 / stack.rs:171-173) as an index space in first-seen order. */
uint32_t yb_n_reads(const yb_ctx *ctx);
int yb_read_at(const yb_ctx *ctx, uint32_t idx, const char **id, size_t *id_len);
int64_t yb_read_index(const yb_ctx *ctx, const char *id, size_t id_len); /* -1 when unknown */

/* ---- consumer side: trait BadPart (stack.rs:35-41) ------------------------------------------ */
/* compute_all_bad_part, stack.rs:143-162, fused with type_of_read (editor/mod.rs:85-100):
 * host CSR -> H2D -> sm_100a kernels -> D2H of classes + bad-region lists. `coverage` is `-c`
 * (cli.rs:49-51), `not_coverage` is `-n` (cli.rs:53-55). */
int yb_compute_all_bad_part(yb_ctx *ctx, uint64_t coverage, double not_coverage);
/* Streamed batches: the reference's `-d/--ondisk` mode (reads2ovl/ondisk.rs, cli.rs:61-70) bounds the working set by
 * flushing its overlap store every --ondisk-buffer-size bytes; here the bound applies to the device and buys overlap.
 * With n_intervals > 0, yb_compute_all_bad_part sends a CSR with more intervals than that through the device in chunks of
 * whole reads (about n_intervals intervals each, a multiple of 1024 reads) on two lanes, each with its own stream and
 * device buffers of chunk size: chunk k + 1 crosses PCIe while chunk k is computed and the results of chunk k - 1 come
 * back. Results, getters and statistics are the same as for the one-shot call (bit-identical). 0 (default) = one shot.
 * Not combined with yb_upload / yb_bind_peers / yb_bind_device_bitmap (those keep the one-shot path). */
int yb_set_chunk_intervals(yb_ctx *ctx, uint32_t n_intervals);
/* get_bad_part, stack.rs:164-169: unknown id => YB_OK with n_gaps = 0, length = 0, cls = NotBad. */
int yb_get_bad_part(yb_ctx *ctx, const char *id, size_t id_len, const uint32_t **gap_pairs,
                    uint32_t *n_gaps, uint64_t *length, uint8_t *cls);
int yb_get_bad_part_at(yb_ctx *ctx, uint32_t idx, const uint32_t **gap_pairs, uint32_t *n_gaps,
                       uint64_t *length, uint8_t *cls);
/* main.rs:80-84 + editor::report (editor/mod.rs:61-83): one line per read, first-seen order. */
int yb_write_report(yb_ctx *ctx, const char *path);
/* Formats read idx's report line (no trailing newline) into out; returns its length or <0. */
int64_t yb_format_report_line(yb_ctx *ctx, uint32_t idx, char *out, size_t cap);
/* 1 byte per read (yb_read_type), and the packed 2-bit-per-read bitmap (read r lives in byte r/4,
 * bits 2*(r%4)..+1) that the multi-GPU path all-gathers. Host views. */
const uint8_t *yb_classes(yb_ctx *ctx, size_t *n);
const uint8_t *yb_class_bitmap(yb_ctx *ctx, size_t *n_bytes);
/* The whole result as a CSR of bad regions: gap_ptr[0..n_reads] and (begin,end) pairs. Host views. */
const uint32_t *yb_gap_ptr(yb_ctx *ctx, size_t *n);
const uint32_t *yb_gaps(yb_ctx *ctx, size_t *n_pairs);
/* The post-detection editors (main.rs:87-117) over the computed results: editor::scrubbing (editor/scrubbing.rs:34),
 * editor::filter (editor/filter.rs:34), editor::extract (editor/extract.rs:34), editor::split (editor/split.rs:34).
 * input_path: fasta / fastq (all four) or paf / m4 (filter and extract only), type by util.rs:39-55; the output keeps
 * the input's format. Needs yb_compute_all_bad_part first (the class of a read is the one computed with that call's
 * not_coverage, which is what main.rs passes to the editors too). Errors mirror error.rs: CantRunOperationOnFile ->
 * YB_ERR_WRONG_FORMAT, UnableToDetectFileFormat -> YB_ERR_UNKNOWN_FORMAT, ReadingError -> YB_ERR_READING. */
typedef enum yb_editor { YB_EDIT_SCRUBB = 0, YB_EDIT_FILTER = 1, YB_EDIT_EXTRACT = 2, YB_EDIT_SPLIT = 3 } yb_editor;
int yb_edit(yb_ctx *ctx, int editor, const char *input_path, const char *output_path);
/* FromReport::new, stack.rs:182-215: load an existing .yacrd report instead of computing. */
int yb_init_report(yb_ctx *ctx, const char *path);
/* A load that fails (YB_ERR_CORRUPT_REPORT; YB_ERR_TOO_LARGE for a length beyond 32 bits) leaves the context empty. */
int yb_init_report_buffer(yb_ctx *ctx, const char *text, size_t n_bytes);

/* ---- staged device API (what yb_compute_all_bad_part is made of) ---------------------------- */
/* Freeze the host store into a CSR (flat (begin,end) buffer + row pointers + lengths) and copy it to HBM;
 * then, once per uploaded CSR and on the device: row statistics, the test 0 <= begin < end <= length of
 * every interval (a read that fails it is computed by the reference's own heap sweep, stack.rs:61-139, instead
 * of the closed form; counts in yb_stats) and the size-class worklist of the rows (16 bytes per read). This is
 * the get_overlaps boundary (stack.rs:149). */
int yb_upload(yb_ctx *ctx);
/* Kernels only; inputs and outputs stay resident in HBM. `stream` is a cudaStream_t (NULL = the
 * context's own stream); the call is asynchronous with respect to the host (the first call after a
 * yb_upload waits for the upload's validation result, 4 bytes). On a caller's stream the step is ordered
 * behind the upload and yb_download behind the step by events; while that stream is being captured into
 * a CUDA graph the events are left out and the caller synchronises around capture and replays. */
int yb_compute_device(yb_ctx *ctx, uint64_t coverage, double not_coverage, void *stream);
/* D2H of classes, bitmap, gap offsets and gaps (synchronises the stream). */
int yb_download(yb_ctx *ctx);
int yb_synchronize(yb_ctx *ctx);
/* Device views for collectives / chaining (valid after yb_compute_device on the same stream). */
void *yb_device_class_bitmap(yb_ctx *ctx, size_t *n_bytes);
/* Make the kernels write the 2-bit bitmap straight into a caller-owned device buffer (e.g. this rank's
 * slot of an in-place all-gather buffer); n_bytes >= 4*ceil(n_reads/16). NULL unbinds. */
int yb_bind_device_bitmap(yb_ctx *ctx, void *device_ptr, size_t n_bytes);
void *yb_stream(yb_ctx *ctx); /* the context's cudaStream_t */

/* ---- peer-memory all-gather of the class bitmap (one process per GPU, same node, NVLink / NVSwitch) -------
 * Instead of a separate collective after the kernels, the ordering kernel stores every word of this rank's 2-bit
 * bitmap straight into slot `rank` of EVERY rank's gather buffer (peer stores over NVLink) and, when its last CTA is
 * done, releases one flag per peer (system scope). Nothing in a step waits for the slowest rank: the gather buffer is
 * [2 x n_ranks x slot_bytes] - step s uses half s & 1 - and a rank starts storing step s only once every peer has
 * finished step s - 1, i.e. (stream order) whatever it read from the half that step s overwrites. The consumer of a
 * step's bitmaps calls yb_peer_wait on its stream first. Buffers are plain cudaMalloc allocations shared through
 * CUDA IPC handles, which the caller exchanges with whatever it has (torch.distributed, MPI, a file). */
#define YB_IPC_HANDLE_BYTES 64
#define YB_MAX_PEERS 16
/* Device buffer (zero-filled) that other processes can map; handle_out receives YB_IPC_HANDLE_BYTES bytes. */
void *yb_peer_alloc(yb_ctx *ctx, size_t n_bytes, void *handle_out);
/* Maps a buffer another process allocated with yb_peer_alloc. */
void *yb_peer_open(yb_ctx *ctx, const void *handle);
int yb_peer_close(yb_ctx *ctx, void *mapped);
int yb_peer_free(yb_ctx *ctx, void *allocated);
/* gather_bufs[p] / flag_bufs[p]: rank p's gather buffer (2 x n_ranks x slot_bytes bytes) and flag buffer (128 bytes,
 * zero-filled; word q = steps rank q has finished, word 31 = steps this rank has finished), as mapped in THIS process
 * (own buffers for p == rank). Every rank must run the same number of detect steps. n_ranks == 0 unbinds. */
int yb_bind_peers(yb_ctx *ctx, void *const *gather_bufs, void *const *flag_bufs, uint32_t n_ranks, uint32_t rank,
                  size_t slot_bytes);
/* Enqueues on `stream` (NULL = the context's) a wait until every rank's slot of this rank's last finished step is
 * complete in this rank's gather buffer (half (steps - 1) & 1). Bounded: a rank that never arrives is reported by the
 * next yb_download as YB_ERR_STATE instead of hanging the GPU. */
int yb_peer_wait(yb_ctx *ctx, void *stream);
void *yb_device_classes(yb_ctx *ctx, size_t *n);
void *yb_device_gap_ptr(yb_ctx *ctx, size_t *n);
void *yb_device_gaps(yb_ctx *ctx, size_t *n_pairs_capacity);
int yb_get_stats(yb_ctx *ctx, yb_stats *out);
/* Measurement aid: runs the once-per-upload kernels (row statistics, interval validation, size-class worklist) again on
 * the resident CSR and returns their device time (CUDA events) in milliseconds. Results of an earlier step are dropped. */
int yb_time_upload_kernels(yb_ctx *ctx, float *ms_out);
/* The device side of a one-shot call on the uploaded CSR, timed with CUDA events on the context's stream: the per-upload
 * kernels again (as yb_time_upload_kernels) and then the first detect step, the one that also tests every interval. The
 * step's results are left for yb_download. With peers bound every rank must make the call (the step all-gathers). */
int yb_time_one_shot(yb_ctx *ctx, uint64_t coverage, double not_coverage, float *ms_upload_kernels, float *ms_first_step);

#ifdef __cplusplus
}
#endif
#endif /* YACRD_B200_H */
