/*
 * spsp_host.h -- C entry points of the C++ host layer (libspsp_host.so).
 *
 * The host layer mirrors the reference's two classes (Subsampler,
 * Comparator) and its two main() functions; these wrappers exist so that the
 * drop-in CLIs, the tests and bench.py (ctypes) all run the same code.
 * Functions marked [cpu] need no GPU (host logic only); everything else needs
 * a CUDA device and fails loudly without one.
 * Buffers returned through out-pointers are malloc'ed: release with spsph_free.
 */
#ifndef SPSP_HOST_H
#define SPSP_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "spsp.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *spsph_last_error(void);
void spsph_free(void *p);

/* The reference's executables as functions (SubSampler.cpp:667-803,
 * Comparator.cpp:464-521): same flags, defaults, outputs in the CWD. */
int spsph_sub_sampler_main(int argc, char **argv);
int spsph_comparator_main(int argc, char **argv);
/* [cpu] The reference's sortCSV helper (sort_csv.cpp:26-122): argv = {prog, matrix.csv[.gz], out.csv, names.txt}. */
int spsph_sort_csv_main(int argc, char **argv);

/* [cpu] compute_threshold (SubSampler.cpp:622-631); s is the float-parsed -s. */
uint64_t spsph_threshold(int k, int m, double s);

/* [cpu] getLineFasta + clean_dna + 2-bit packing of records >= min_len.
 * words has spsp_packed_words(n_bases) entries, rec_off n_rec+1 entries. */
int spsph_pack_fasta(const uint8_t *fasta, size_t n, uint32_t min_len, uint32_t **words, uint64_t *n_bases,
                     uint64_t **rec_off, uint64_t *n_rec);

/* [cpu] Exact post-pass: hits (any order) on a packed buffer -> sketch bytes
 * (before gzip).  The GPU path feeds it the scan kernel's output. */
int spsph_postpass(const uint32_t *packed, const uint64_t *rec_off, uint64_t n_rec, const spsp_hit *hits,
                   uint64_t n_hits, int k, int m, double s, unsigned abundance, uint8_t **out, size_t *out_len,
                   uint64_t *selected_kmers);

/* [cpu] Sketch bytes -> distinct (minimizer, canonical k-mer) elements. */
int spsph_decode_sketch(const uint8_t *sketch, size_t n, int *k, int *m, uint64_t *n_elems, uint32_t **minimizer,
                        uint64_t **kmer_lo, uint64_t **kmer_hi);

/* [cpu] CSV text of print_containment (jaccard=0) / print_jaccard (jaccard=1).
 * inter is row-major with leading dimension n; full_rows=0: pair (i<j) at
 * inter[i*n+j]; full_rows=1: row i complete. */
int spsph_format_csv(const char *const *names, uint32_t n, uint32_t query_size, const uint32_t *inter, int full_rows,
                     const uint64_t *sizes, int jaccard, unsigned precision, double min_threshold, uint8_t **out,
                     size_t *out_len);

/* [cpu] The same text written to `path` as a gzip file, streamed: row blocks are formatted and compressed by
 * `threads` workers and written in order as consecutive gzip members (memory holds one wave of blocks, not the
 * whole matrix as text).  text_bytes (may be NULL) = CSV size before compression. */
int spsph_write_csv_gz(const char *path, const char *const *names, uint32_t n, uint32_t query_size, const uint32_t *inter,
                       int full_rows, const uint64_t *sizes, int jaccard, unsigned precision, double min_threshold,
                       int threads, uint64_t *text_bytes);

/* GPU: sketch n FASTA texts held in memory with `threads` host workers on
 * `device`; out[i]/out_len[i] receive the sketch bytes (before gzip).
 * timings (may be NULL): [0]=pack s, [1]=scan s (submit..collect), [2]=post-pass s,
 * summed over inputs; launches (may be NULL) = kernels launched. */
int spsph_sketch_buffers(int device, int k, int m, double s, unsigned abundance, int scan_mode, uint32_t n,
                         const uint8_t *const *fasta, const size_t *len, int threads, uint8_t **out, size_t *out_len,
                         double *timings, uint64_t *launches);

/* Persistent sketcher: one device context + `threads` workers (one stream
 * each), reused across calls -- what a long-running caller keeps around. */
typedef struct spsph_sketcher spsph_sketcher;
int spsph_sketcher_create(int device, int k, int m, double s, unsigned abundance, int scan_mode, int threads,
                          spsph_sketcher **out);
int spsph_sketcher_destroy(spsph_sketcher *sk);
/* Same contract as spsph_sketch_buffers. */
int spsph_sketcher_run(spsph_sketcher *sk, uint32_t n, const uint8_t *const *fasta, const size_t *len, uint8_t **out,
                       size_t *out_len, double *timings, uint64_t *launches);
/* The sketcher's device context (for device-resident scans through spsp.h). */
spsp_ctx *spsph_sketcher_ctx(spsph_sketcher *sk);

/* [cpu] Post-pass of a batch: `packed` holds n_inputs inputs back to back,
 * input i starting at base offset base_off[i] (a multiple of 64) with
 * rec_off[rec_first[i] .. rec_first[i+1]) as its n_rec_i + 1 record offsets
 * (relative to base_off[i]; rec_first has n_inputs + 1 entries); hits are positions in the whole buffer (any order, e.g. one
 * scan launch over everything).  Hits in the padding between inputs are
 * ignored.  `threads` host workers; out[i]/out_len[i] as above. */
int spsph_postpass_batch(const uint32_t *packed, uint32_t n_inputs, const uint64_t *base_off, const uint64_t *n_bases,
                         const uint64_t *rec_off, const uint64_t *rec_first, const spsp_hit *hits, uint64_t n_hits,
                         int k, int m, double s, unsigned abundance, int threads, uint8_t **out, size_t *out_len);

/* GPU: all-vs-all (query_size == n) or query-vs-all compare of n sketches held
 * in memory.  inter: rows x n uint32 (rows = n or query_size), sizes: n. */
int spsph_compare_buffers(int n_gpus, uint32_t n, uint32_t query_size, const uint8_t *const *sketch, const size_t *len,
                          uint32_t *inter, uint64_t *sizes, int *full_rows, float *kernel_ms, uint64_t *launches);

/* Persistent comparer: keeps its device context(s) between calls. */
typedef struct spsph_comparer spsph_comparer;
int spsph_comparer_create(int n_gpus, int threads, spsph_comparer **out);
int spsph_comparer_destroy(spsph_comparer *c);
/* Same contract as spsph_compare_buffers; timings (may be NULL): [0] = decode s, [1] = device s. */
int spsph_comparer_run(spsph_comparer *c, uint32_t n, uint32_t query_size, const uint8_t *const *sketch,
                       const size_t *len, uint32_t *inter, uint64_t *sizes, int *full_rows, float *kernel_ms,
                       uint64_t *launches, double *timings);

/* Batch pipeline (csrc/host/pipeline.h): the GPU-shaped `sub_sampler -f`.  Host
 * threads clean + pack all inputs into one pinned staging buffer (async H2D per
 * finished input), one scan + device post-pass yields every sketch, and the
 * compare stage starts from the elements left on the device. */
typedef struct spsph_pipeline spsph_pipeline;
int spsph_pipeline_create(int device, int k, int m, double s, unsigned abundance, int threads, spsph_pipeline **out);
int spsph_pipeline_destroy(spsph_pipeline *p);
spsp_ctx *spsph_pipeline_ctx(spsph_pipeline *p);
/* Upper bound of one device batch in bases (default 2^30); larger jobs run as several batches. */
int spsph_pipeline_set_max_batch_bases(spsph_pipeline *p, uint64_t bases);
/* Who cleans + packs the inputs (csrc/host/pipeline.h, enum Ingest): 0 = host threads, 1 = the device (raw text
 * over PCIe, ingest kernels), 2 = both (the default, or what the environment variable SPSP_INGEST =
 * host|device|auto names): inputs that look like read sets go over as text, the others are packed by the workers,
 * which -- when there are few of them -- also send inputs from the back of the queue as text whenever a text
 * lane is idle.  In modes 1 / 2 text
 * held in PINNED memory is copied asynchronously and must stay valid until the job's device half has run
 * (spsph_pipeline_sketch / _finish returned); pageable text is staged by the copy call itself. */
int spsph_pipeline_set_ingest(spsph_pipeline *p, int mode);
/* Input i is fasta[i]/len[i] (FASTA text in memory) or, when fasta is NULL or
 * fasta[i] is NULL, the FASTA(.gz) file paths[i].  out/out_len as in
 * spsph_sketch_buffers; ok[i] = 0 for an unopenable file (may be NULL).
 * stats (may be NULL, 14 doubles): prep s, pack s, device s, assemble s, scan ms,
 * post-pass ms, hits, compare elements, H2D bytes, D2H bytes, batches, bases,
 * inputs ingested on the device, ingest kernel ms. */
int spsph_pipeline_sketch(spsph_pipeline *p, uint32_t n, const uint8_t *const *fasta, const size_t *len,
                          const char *const *paths, uint8_t **out, size_t *out_len, int *ok, double *stats,
                          uint64_t *launches);
/* spsph_pipeline_sketch in two halves, to overlap jobs on two pipelines: _pack prepares, packs and queues the
 * H2D copies (returns when the host work is done; the inputs may be released then); _finish runs the device
 * phase and delivers the results of that job (same n). */
int spsph_pipeline_pack(spsph_pipeline *p, uint32_t n, const uint8_t *const *fasta, const size_t *len,
                        const char *const *paths);
int spsph_pipeline_finish(spsph_pipeline *p, uint32_t n, uint8_t **out, size_t *out_len, int *ok, double *stats,
                          uint64_t *launches);
/* Element offsets (n + 1 entries) of the last spsph_pipeline_sketch call; *on_device = 1 when the
 * elements are still resident on the GPU (single batch: spsp_batch_elements on slot 0 of spsph_pipeline_ctx). */
int spsph_pipeline_elem_off(spsph_pipeline *p, uint64_t *off, int *on_device);
/* Compare the sketches of the last spsph_pipeline_sketch call (contract of
 * spsph_compare_buffers: inter rows x n with rows = n or query_size). */
int spsph_pipeline_compare(spsph_pipeline *p, uint32_t query_size, uint32_t *inter, uint64_t *sizes, int *full_rows,
                           float *kernel_ms, uint64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* SPSP_HOST_H */
