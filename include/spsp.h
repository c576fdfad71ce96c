/*
 * spsp.h -- C ABI of the B200 device layer of supersampler_b200.
 *
 * The reference (TimRouze/supersampler) has no library / FFI surface: its hot
 * path lives inside two executables.  This header is the boundary the
 * north star asks for ("C++ host code calls CUDA through a thin C-ABI
 * layer"): each entry point names the reference code it replaces, so a
 * maintainer of the reference can swap the body of that function for one call.
 * Plain pointers and sizes only; no C++/torch types; no exceptions cross it.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; the message of the
 *     last error on the calling thread is spsp_last_error().
 *   - the caller owns every host pointer; the library owns device memory,
 *     except for the *_device entry points which work on caller-owned device
 *     buffers (used for the device-resident benchmark and for NCCL plumbing).
 *   - one context per (host thread group, GPU); calls on one (context, slot)
 *     pair must be serialised by the caller; different slots map to different
 *     CUDA streams and may be driven from different host threads.
 *   - there is NO CPU fallback: without a CUDA device spsp_create fails.
 *
 * Packed sequence format ("2-bit"): bases A=0 C=1 T=2 G=3 (reference
 * utils.cpp:13-16, code = (c>>1)&3), 16 bases per little-endian uint32 word,
 * first base in the two most significant bits.  Buffers handed to the scan must
 * be readable up to spsp_packed_words(n_bases) words (zero padded).
 */
#ifndef SPSP_H
#define SPSP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPSP_ABI_VERSION 1

typedef struct spsp_ctx spsp_ctx;

/* One selected m-mer ("hit"): an m-mer whose canonical form hashes <= T.
 * 16 bytes.  pos = index of the m-mer's first base in the scanned buffer. */
typedef struct {
    uint64_t pos;
    uint32_t canon;   /* canonical m-mer value = min(fw, revcomp)           */
    uint32_t rev;     /* 1 when canon != forward m-mer (reference is_rev)  */
} spsp_hit;

/* Scan kernel selector (spsp_scan_config). */
enum {
    SPSP_SCAN_AUTO = 0,    /* q-gram filter kernel when profitable, else dense */
    SPSP_SCAN_DENSE = 1,   /* full XXH64 at every position                    */
    SPSP_SCAN_FILTER = 2   /* shared-memory aligned q-gram filter + exact verify */
};

int spsp_abi_version(void);
const char *spsp_last_error(void);
int spsp_device_count(int *n);
/* Initialise the driver and the device's primary context now (~1 s in a fresh process): a CLI calls it on a
 * background thread while it reads its inputs, so that the first spsp_create does not pay for it. */
int spsp_warmup(int device);

/* Words (uint32) a packed buffer of n_bases must provide to the scan,
 * including the zero padding the kernels may read. */
uint64_t spsp_packed_words(uint64_t n_bases);

/* Pinned host memory for the async copies (cudaHostAlloc / cudaFreeHost). */
int spsp_host_alloc(void **p, size_t bytes);
int spsp_host_free(void *p);

/* Context: fixes (k, m, T) like the reference's Subsampler constructor
 * (SubSampler.h:63-88); threshold is compute_threshold's value
 * (SubSampler.cpp:622-631) evaluated by the host.  n_slots = number of
 * independent streams (>=1). */
int spsp_create(int device, int k, int m, uint64_t threshold, int n_slots, spsp_ctx **ctx);
int spsp_destroy(spsp_ctx *ctx);
int spsp_scan_config(spsp_ctx *ctx, int mode);
/* Which filter table the context built for its (m, T): kind 0 = bit table,
 * 1 = byte table of phase masks, 2 = bank-private bit table + exact hash set
 * (m == 11, few selected m-mers), -1 = none (dense kernel only); g = probe
 * stride in bases, n_selected = selected forward m-mers.  Builds the table on
 * first use, like the first scan does. */
int spsp_scan_filter_info(spsp_ctx *ctx, int *kind, int *g, uint64_t *n_selected);

/* ---- sketch stage ------------------------------------------------------
 * Replaces the per-base loop of Subsampler::parse_fasta_test
 * (SubSampler.cpp:367-440: updateM/updateRCM :36-53, canonical m-mer,
 * unrevhash -> XXHash64::hash(&x,8,1312) :64-67, threshold test :405/:443)
 * by its closed form: report every position whose canonical m-mer hashes
 * <= T.  The exact state-machine replay over this sparse list is host code
 * (supersampler_b200/csrc/host/postpass.cpp).
 *
 * submit: async H2D of `packed` (pinned recommended) + kernel + async D2H of
 *         the hits into an internal pinned buffer; returns immediately.
 * collect: waits for the slot, copies the hits (unordered) to hits_out.
 *          If n_hits > cap the call fails with -2 and *n_hits holds the
 *          needed capacity (call again with a larger buffer). */
int spsp_scan_submit(spsp_ctx *ctx, int slot, const uint32_t *packed, uint64_t n_bases);
int spsp_scan_collect(spsp_ctx *ctx, int slot, spsp_hit *hits_out, uint64_t cap, uint64_t *n_hits);

/* Device-resident variant: d_packed / d_hits / d_count are device pointers
 * owned by the caller; d_count (uint64) is zeroed by the call; launches on the
 * slot's stream and returns without synchronising (spsp_sync to wait).
 * Hits beyond cap are counted but not stored. */
int spsp_scan_device(spsp_ctx *ctx, int slot, const uint32_t *d_packed, uint64_t n_bases,
                     spsp_hit *d_hits, uint64_t cap, uint64_t *d_count);
int spsp_sync(spsp_ctx *ctx, int slot);
/* Time of the last scan kernel(s) launched on the slot, CUDA events on the
 * slot's stream (milliseconds); valid after spsp_sync / collect. */
int spsp_scan_last_kernel_ms(spsp_ctx *ctx, int slot, float *ms);
/* Raw stream handle of a slot (cudaStream_t as void*), for callers that
 * interleave their own work (NCCL through torch.distributed). */
int spsp_stream(spsp_ctx *ctx, int slot, void **stream);

/* ---- compare stage -----------------------------------------------------
 * Replaces Comparator::compare_sketches' N-way merge + count_intersection +
 * compute_scores (Comparator.cpp:39-74, :177-287): for every pair of
 * sketches the number of (minimizer bucket, canonical k-mer) entries they
 * share.  The host decodes sketch files (strDecompressor / inject_minimizer /
 * canonize, Comparator.cpp:78-92, :97-154) into per-sketch element lists.
 *
 * Elements of a sketch: distinct (minimizer, canonical k-mer) pairs, sorted by
 * minimizer (any order inside a bucket).  kmer_hi is NULL when k <= 32.
 * sketch_off has n_sketches+1 entries (element offsets into the arrays).
 * load replaces all previously loaded sketches. */
int spsp_cmp_load(spsp_ctx *ctx, uint32_t n_sketches, const uint64_t *sketch_off,
                  const uint32_t *minimizer, const uint64_t *kmer_lo, const uint64_t *kmer_hi);
/* Same with device-resident arrays (after an NCCL all-gather): the context
 * keeps the pointers, the caller keeps them alive until the next load. */
int spsp_cmp_load_device(spsp_ctx *ctx, uint32_t n_sketches, const uint64_t *sketch_off_host,
                         const uint32_t *d_minimizer, const uint64_t *d_kmer_lo,
                         const uint64_t *d_kmer_hi);
/* Intersections for rows [row_begin,row_end) x columns [col_begin,col_end):
 * out[(i-row_begin)*ld + (j-col_begin)] += |K_i ∩ K_j| (row-major uint32,
 * ld >= col_end-col_begin, zeroed by the caller).  With symmetric != 0 only
 * tiles on or above the diagonal are computed (entries with i < j are valid).
 * tile_rank / tile_ranks deal the work units (a column tile x a run of row
 * tiles) round-robin for multi-GPU runs (0 / 1 for a single GPU).  out is HOST
 * memory; load + run synchronise once, when the counts have arrived. */
int spsp_cmp_run(spsp_ctx *ctx, uint32_t row_begin, uint32_t row_end, uint32_t col_begin,
                 uint32_t col_end, int symmetric, uint32_t tile_rank, uint32_t tile_ranks,
                 uint32_t *out, uint64_t ld);
/* Device-output variant: d_out is a device pointer (not zeroed by the call);
 * returns when the counts are in d_out. */
int spsp_cmp_run_device(spsp_ctx *ctx, uint32_t row_begin, uint32_t row_end, uint32_t col_begin,
                        uint32_t col_end, int symmetric, uint32_t tile_rank, uint32_t tile_ranks,
                        uint32_t *d_out, uint64_t ld);
int spsp_cmp_last_kernel_ms(spsp_ctx *ctx, float *ms);

/* ---- sketch stage, whole batch on the device ------------------------------
 * One call = scan + exact post-pass of a batch of inputs (files) packed back
 * to back, entirely on the GPU (csrc/device/postpass.cu; CUB does the generic
 * radix sorts / prefix sums): replaces parse_fasta_test's loop AND
 * handle_superkmer / the writer loop (SubSampler.cpp:243-302, :352-504) for
 * every input at once, and leaves the comparator's decoded elements
 * (Comparator.cpp:97-264) resident on the device for spsp_cmp_load_batch.
 *
 * Records are described by host arrays (n_rec entries, ascending, global base
 * offsets into the packed buffer): [rec_begin, rec_end) and the input each
 * record belongs to (0 <= rec_input < n_inputs, non-decreasing).
 * The three record arrays may also be DEVICE memory (all three): they are then
 * used in place, without validation or copy (read sets of short reads carry
 * millions of records per batch; a caller that keeps them resident pays the
 * upload once).
 * Result pointers stay valid until the next batch call on the same slot. */
typedef struct {
    const uint8_t *body;         /* sketch bytes after the header line, inputs back to back   */
    const uint64_t *body_off;    /* [n_inputs + 1] byte range of each input inside body       */
    const uint64_t *selected;    /* [n_inputs] selected k-mer occurrences (header field 3)    */
    const uint64_t *elem_off;    /* [n_inputs + 1] compare elements (distinct canonical k-mers per bucket) */
    uint64_t n_hits, n_elems;
    float scan_ms, post_ms;      /* CUDA events on the slot's stream */
} spsp_batch_result;

/* packed: HOST buffer (pinned recommended), copied by the call. */
int spsp_sketch_batch(spsp_ctx *ctx, int slot, const uint32_t *packed, uint64_t n_bases, const uint64_t *rec_begin,
                      const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec, uint32_t n_inputs,
                      unsigned abundance, spsp_batch_result *res);
/* d_packed: caller-owned DEVICE buffer of spsp_packed_words(n_bases) words. */
int spsp_sketch_batch_device(spsp_ctx *ctx, int slot, const uint32_t *d_packed, uint64_t n_bases,
                             const uint64_t *rec_begin, const uint64_t *rec_end, const uint32_t *rec_input,
                             uint64_t n_rec, uint32_t n_inputs, unsigned abundance, spsp_batch_result *res);
/* Staged variant for host pipelines that pack inputs concurrently: reserve the
 * slot's device buffer once (total_words uint32 words, synchronises the slot),
 * let any number of host threads upload finished regions (async H2D on the
 * slot's stream; host_words pinned for true overlap; thread-safe), then run the
 * batch on what was uploaded.  Regions never uploaded hold stale bytes: they
 * must not be covered by a record. */
int spsp_batch_reserve(spsp_ctx *ctx, int slot, uint64_t total_words);
int spsp_batch_upload(spsp_ctx *ctx, int slot, uint64_t word_off, const uint32_t *host_words, uint64_t n_words);
int spsp_sketch_batch_staged(spsp_ctx *ctx, int slot, uint64_t n_bases, const uint64_t *rec_begin,
                             const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec, uint32_t n_inputs,
                             unsigned abundance, spsp_batch_result *res);
/* ---- device-side ingest (optional front end of the staged batch) -----------
 * Replaces getLineFasta + clean_dna (utils.cpp:706-718, :675-702) and the 2-bit
 * packing (utils.cpp:13-16) for inputs whose raw FASTA text is sent to the GPU
 * as it is (csrc/device/ingest.cu): the first line of an input and every line
 * that starts with '>' is a header; on the other lines ACGTacgt are bases,
 * everything else is deleted.  Used when host cores are scarce (several GPUs
 * per box) or records are short (read sets): the host only moves bytes.
 *
 *   reserve : device text buffer of text_bytes (grow-only; synchronises when it grows)
 *   upload  : async H2D of a slice of text to byte offset byte_off (host_text
 *             pinned for true overlap; callable from any thread) on text lane
 *             `lane` (0 / 1; < 0: the lanes alternate).  Text has two copy streams
 *             of its own, beside the two that carry packed words.
 *   pack    : for n_text inputs -- text [text_off[i], +text_len[i]) (text_off a
 *             multiple of 16), region of the staged batch buffer starting at
 *             word word_off[i] (spsp_batch_reserve'd, spsp_packed_words(text_len[i])
 *             words, ascending), batch input index input_index[i] (ascending) --
 *             runs the three ingest passes; the record table (every header line
 *             starts a record, empty and short records included) stays on the
 *             slot and is merged into the next spsp_sketch_batch_staged call,
 *             whose host record arrays then describe the host-packed inputs only.
 *             n_bases_out / n_rec_out (may be NULL): cleaned bases / records per input.
 *             One host synchronisation (the table is sized from the count). */
int spsp_batch_text_reserve(spsp_ctx *ctx, int slot, uint64_t text_bytes);
int spsp_batch_text_upload(spsp_ctx *ctx, int slot, int lane, uint64_t byte_off, const uint8_t *host_text,
                           uint64_t n_bytes);
/* Waits until the copies queued on text lane `lane` (0 / 1; < 0: both) have finished: lets a caller keep a
 * bounded number of text uploads in flight. */
int spsp_batch_upload_wait(spsp_ctx *ctx, int slot, int lane);
/* *idle = 1 when nothing is queued or running on text lane `lane` (0 / 1): the previous raw input has arrived. */
int spsp_batch_upload_idle(spsp_ctx *ctx, int slot, int lane, int *idle);
int spsp_batch_text_pack(spsp_ctx *ctx, int slot, uint32_t n_text, const uint64_t *text_off, const uint64_t *text_len,
                         const uint64_t *word_off, const uint32_t *input_index, uint64_t *n_bases_out,
                         uint64_t *n_rec_out);
/* k-mers of every input of the batch just sketched on the slot (sum over its records of at least k bases of
 * length - k + 1: read_kmer, SubSampler.cpp:343-347), from the record table the device used. */
int spsp_batch_record_kmers(spsp_ctx *ctx, int slot, uint32_t n_inputs, uint64_t *kmers_out);
/* CUDA-event time of the last pack's kernels (milliseconds). */
int spsp_batch_text_last_ms(spsp_ctx *ctx, int slot, float *ms);
/* Diagnostics / tests: the record table the last pack left on the slot (cap
 * entries per array, -2 with *n_rec set when too small), and a copy of words of
 * the slot's staged batch buffer. */
int spsp_batch_text_records(spsp_ctx *ctx, int slot, uint64_t *rec_begin, uint64_t *rec_end, uint32_t *rec_input,
                            uint64_t cap, uint64_t *n_rec);
int spsp_batch_download(spsp_ctx *ctx, int slot, uint64_t word_off, uint32_t *host_words, uint64_t n_words);

/* Load the compare stage with the elements the last batch on `slot` left on the
 * device (one sketch per input): the sketch -> compare hand-off without files. */
int spsp_cmp_load_batch(spsp_ctx *ctx, int slot);
/* Copy the last batch's elements to the host (n_elems entries each; kmer_hi may
 * be NULL when k <= 32) and/or return the device pointers (any argument may be NULL). */
int spsp_batch_elements(spsp_ctx *ctx, int slot, uint32_t *minimizer, uint64_t *kmer_lo, uint64_t *kmer_hi,
                        const uint32_t **d_minimizer, const uint64_t **d_kmer_lo, const uint64_t **d_kmer_hi);

/* ---- dense totals ---------------------------------------------------------
 * The part of the reference's loop that looks at EVERY k-mer, not only the
 * selected ones (csrc/device/dense.cu): rolling canonical m-mer hash at every
 * position, warp-shuffle sliding-window minimum over the k-m+1 m-mers of each
 * k-mer, and an exact parallel replay of the minimizer state machine's
 * super-k-mer boundaries (SubSampler.cpp:374-398, :401, :429-431, :441-454).
 * Per input: total_superkmers = total_superkmer_number of print_stat
 * (SubSampler.cpp:633-665), selected_kmers = number of k-mers whose minimizer
 * hashes <= T (must equal the sketch header's third field).  Records as in
 * spsp_sketch_batch; output arrays have n_inputs entries (either may be NULL);
 * kernel_ms (may be NULL) = CUDA-event time of the kernels.  Synchronous. */
int spsp_dense_stats(spsp_ctx *ctx, int slot, const uint32_t *packed, uint64_t n_bases, const uint64_t *rec_begin,
                     const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec, uint32_t n_inputs,
                     uint64_t *total_superkmers, uint64_t *selected_kmers, float *kernel_ms);
/* d_packed: caller-owned DEVICE buffer. */
int spsp_dense_stats_device(spsp_ctx *ctx, int slot, const uint32_t *d_packed, uint64_t n_bases,
                            const uint64_t *rec_begin, const uint64_t *rec_end, const uint32_t *rec_input,
                            uint64_t n_rec, uint32_t n_inputs, uint64_t *total_superkmers, uint64_t *selected_kmers,
                            float *kernel_ms);
/* On what spsp_batch_upload staged on the slot (the batch just sketched).  rec_begin == NULL with n_rec == 0:
 * the record table of that batch as the device saw it (host-packed, ingested on the device, or both merged). */
int spsp_dense_stats_staged(spsp_ctx *ctx, int slot, uint64_t n_bases, const uint64_t *rec_begin,
                            const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec, uint32_t n_inputs,
                            uint64_t *total_superkmers, uint64_t *selected_kmers, float *kernel_ms);

/* ---- multi-GPU compare (one process per GPU) --------------------------------
 * The compare stage's only exchange step, on NCCL over NVLink (libnccl.so.2 is
 * loaded at run time; single-GPU callers never need it).  Rank 0 creates an id
 * (spsp_nccl_unique_id, 128 bytes), hands it to every rank by any side channel
 * (torch.distributed broadcast, MPI, a file), every rank calls spsp_nccl_init
 * (at most 64 ranks).  A context owns its communicator, so exchanges of
 * different contexts may be in flight at the same time.
 *
 * spsp_cmp_exchange: every rank brings n_local sketches -- the first q_local
 * of them queries (Comparator.cpp:7-21, 512-515: `-q` lists go first) -- as
 * device-resident element arrays plus a host size list.  The union, in the
 * order [queries of rank 0, 1, ... | references of rank 0, 1, ...], is compared
 * all-vs-all (symmetric != 0; then q_local == n_local) or queries x all
 * (Comparator.cpp:328-359).  One payload all-gather into fixed-capacity slots
 * (counts travel inside the payload), the tile plan and the join on the device,
 * the owned tiles sent to rank 0, one host synchronisation.  Rank 0 receives
 * inter_out (rows x columns, row-major, ld >= columns, overwritten; all-vs-all:
 * entries i < j valid); every rank receives sizes_out[columns] (|K_i|, may be
 * NULL), *n_rows and *n_cols.  cap_rows / cap_cols = capacity of the outputs
 * (-2 with the dimensions set when too small).  Collective: every rank must
 * call it with the same `symmetric`; argument errors that only one rank can
 * see travel in the payload, so no rank is left waiting.
 *
 * spsp_cmp_exchange_batch: the same, all-vs-all, for the sketches the last
 * batch on `slot` left on every rank's device. */
int spsp_nccl_unique_id(uint8_t *id128);
int spsp_nccl_init(spsp_ctx *ctx, const uint8_t *id128, int rank, int world);
int spsp_cmp_exchange(spsp_ctx *ctx, uint32_t n_local, uint32_t q_local, const uint64_t *sizes_local,
                      const uint32_t *d_minimizer, const uint64_t *d_kmer_lo, const uint64_t *d_kmer_hi, int symmetric,
                      uint32_t *inter_out, uint64_t ld, uint64_t *sizes_out, uint32_t cap_rows, uint32_t cap_cols,
                      uint32_t *n_rows, uint32_t *n_cols, float *kernel_ms);
int spsp_cmp_exchange_batch(spsp_ctx *ctx, int slot, uint32_t *inter_out, uint64_t ld, uint64_t *sizes_out,
                            uint32_t cap_sketches, uint32_t *n_total, float *kernel_ms);

/* Number of kernels this library launched on the context since creation. */
int spsp_launch_count(spsp_ctx *ctx, uint64_t *n);

#ifdef __cplusplus
}
#endif
#endif /* SPSP_H */
