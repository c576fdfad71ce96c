// Host-side sequence ingest: FASTA(.gz) -> cleaned, 2-bit packed records.
// Mirrors the reference's getLineFasta + clean_dna (utils.cpp:706-718, 675-702)
// byte for byte in *behaviour*; the implementation is a streaming state machine
// that packs straight into (pinned) memory for the async H2D copy.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace spsp_host {

// Growable buffer of packed words; pinned (cudaHostAlloc through the C ABI)
// when `pinned` is set, plain malloc otherwise (CPU-only tests).
class WordBuf {
public:
    explicit WordBuf(bool pinned) : pinned_(pinned) {}
    ~WordBuf();
    WordBuf(const WordBuf &) = delete;
    WordBuf &operator=(const WordBuf &) = delete;
    void reserve(uint64_t words);
    // Use caller-owned memory of fixed size (a region of a batch staging buffer);
    // reserve() beyond it throws.  detach() before the owner frees it.
    void attach(uint32_t *p, uint64_t words);
    void detach();
    uint32_t *data() { return p_; }
    const uint32_t *data() const { return p_; }
    uint64_t capacity() const { return cap_; }

private:
    uint32_t *p_ = nullptr;
    uint64_t cap_ = 0;
    bool pinned_;
    bool external_ = false;
};

// All records of one input that are at least `min_len` bases long, packed back
// to back (A0 C1 T2 G3, 16 bases per u32, first base in the MSBs).
struct PackedInput {
    explicit PackedInput(bool pinned) : words(pinned) {}
    WordBuf words;
    uint64_t n_bases = 0;
    std::vector<uint64_t> rec_off;     // n_rec + 1 base offsets
    uint64_t dropped_records = 0;      // records shorter than min_len
    void clear() { n_bases = 0; rec_off.assign(1, 0); dropped_records = 0; }
};

// Streaming FASTA cleaner/packer.  feed() any number of byte chunks, then finish().
class FastaPacker {
public:
    FastaPacker(PackedInput &out, uint32_t min_len);
    void feed(const uint8_t *p, size_t n);
    void finish();
    // Number of leading output words that are final already (in output bit order, never rewritten):
    // lets a caller copy a long input to the device while the rest is still being packed.
    uint64_t commit();

private:
    void end_record();
    void feed_piece(const uint8_t *p, size_t n);
    PackedInput &out_;
    std::vector<uint32_t> scratch_;    // words [converted_, word_idx_) before the sweep into the output buffer
    uint32_t min_len_;
    enum State { HEADER, LINE_START, SEQ } state_ = HEADER;
    // Pending bases not yet written: base j of the run at bits 2j..2j+1 (little-endian
    // order, what PEXT compaction produces); converted to the output format (first
    // base in the MSBs of each u32) when 32 bases are complete.
    uint64_t acc_ = 0;
    int fill_ = 0;                     // bits in acc_ (< 64, even); bits above are zero
    uint64_t word_idx_ = 0;            // next u32 word to write (even until finish())
    uint64_t converted_ = 0;           // words already swept into output bit order
    uint64_t rec_start_ = 0;           // base offset where the current record began
    // checkpoint of the pending bases at rec_start_, to drop short records
    uint64_t ck_acc_ = 0;
    int ck_fill_ = 0;
    uint64_t ck_word_idx_ = 0;
    bool any_input_ = false;
};

// Read a whole file (gzip auto-detected by magic, like the reference's zstr,
// include/zstr.hpp:154-167) and pack it.  Returns false if it cannot be opened.
bool pack_fasta_file(const std::string &path, uint32_t min_len, PackedInput &out, uint64_t *file_bytes = nullptr);
void pack_fasta_buffer(const uint8_t *p, size_t n, uint32_t min_len, PackedInput &out);

// Whole (possibly gzip) file -> bytes.  Returns false if it cannot be opened.
bool read_file_maybe_gz(const std::string &path, std::vector<uint8_t> &out);
bool write_gz(const std::string &path, const uint8_t *p, size_t n, int level);

inline unsigned base_at(const uint32_t *w, uint64_t i)
{
    return (w[i >> 4] >> (30 - 2 * (i & 15))) & 3u;
}

}  // namespace spsp_host
