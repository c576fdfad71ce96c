// Device-side FASTA ingest: raw text in HBM -> cleaned 2-bit regions of a batch + its record table.
// Declarations shared by ingest.cu and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spsp {

constexpr int ING_THREADS = 1024;                    // one CTA = one tile
constexpr int ING_TILE = ING_THREADS * 16;           // bytes of text per tile (16 per thread)

// One text input of a batch.  text_off is a multiple of 16 (so every thread's 16 bytes are one aligned load);
// the region of the packed buffer starts at word word_off (a 64-base boundary, like every input of a batch).
struct IngestInput {
    uint64_t text_off, text_len;      // bytes in the device text buffer
    uint64_t word_off;                // first word of the input's region in the packed buffer
    uint64_t tile0;                   // first tile of this input (tiles never span inputs)
    uint32_t input;                   // index of the input in the batch (rec_input)
    uint32_t pad;
};

struct IngestTile {                   // written by the summary pass
    uint32_t cnt_pre;                 // bases before the tile's first line start (count only if the line is sequence)
    uint32_t cnt_post;                // bases in sequence lines from the first line start on
    uint32_t n_hdr;                   // header lines that start in the tile (= records)
    uint32_t flags;                   // bit 0: the tile holds a line start; bit 1: its last line start is a header
};

struct IngestCarry {                  // written by the carry pass: what a tile needs to know about the text before it
    uint64_t base_prefix;             // cleaned bases of the input before this tile
    uint32_t hdr_prefix;              // records of the input that start before this tile
    uint32_t in_header;               // 1: the tile starts inside a header line
};

struct IngestTotals {                 // per input, device + host copy
    uint64_t n_bases, n_rec, rec_base;
};

cudaError_t launch_ingest_summary(const uint8_t *d_text, const IngestInput *d_in, uint32_t n_in, uint64_t n_tiles,
                                  IngestTile *d_tiles, cudaStream_t st);
// One warp per input walks its tiles; then one CTA turns the per-input record counts into rec_base and
// writes the grand totals {bases, records} to d_grand[0..1].
cudaError_t launch_ingest_carry(const IngestInput *d_in, uint32_t n_in, const IngestTile *d_tiles, IngestCarry *d_carry,
                                IngestTotals *d_tot, uint64_t *d_grand, cudaStream_t st);
// Writes the 2-bit codes (regions must be zero) and the records: rec r of input i is
// [rec_begin, rec_end) in batch coordinates (16 * word_off + cleaned offset), rec_input = IngestInput::input.
cudaError_t launch_ingest_write(const uint8_t *d_text, const IngestInput *d_in, uint32_t n_in, uint64_t n_tiles,
                                const IngestCarry *d_carry, const IngestTotals *d_tot, uint32_t *d_packed,
                                uint64_t *d_rec_begin, uint64_t *d_rec_end, uint32_t *d_rec_input, cudaStream_t st);
// Merge two ascending record tables (by begin; the inputs' regions are disjoint) into one.
cudaError_t launch_record_merge(const uint64_t *a_begin, const uint64_t *a_end, const uint32_t *a_input, uint64_t na,
                                const uint64_t *b_begin, const uint64_t *b_end, const uint32_t *b_input, uint64_t nb,
                                uint64_t *o_begin, uint64_t *o_end, uint32_t *o_input, cudaStream_t st);

// d_kmers[input] += k-mers of the input's records (records of at least k bases: length - k + 1); d_kmers zeroed by the caller.
cudaError_t launch_record_kmers(const uint64_t *rec_begin, const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec,
                                uint32_t k, unsigned long long *d_kmers, cudaStream_t st);

}  // namespace spsp
