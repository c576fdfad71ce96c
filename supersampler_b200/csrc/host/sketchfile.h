// Sketch file (after gunzip) -> compare-stage elements, and CSV output.
// Format: SURVEY.md App. B / reference SubSampler.cpp:459-504 (writer) and
// Comparator.cpp:23-37, :78-92, :97-154, :177-264, :291-323 (reader).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace spsp_host {

// Distinct (bucket minimizer, canonical k-mer) pairs of one sketch, buckets in
// file order (ascending minimizer), k-mers sorted inside a bucket.
struct SketchElems {
    int k = 0, m = 0;
    std::vector<uint32_t> minim;
    std::vector<uint64_t> klo, khi;      // khi only filled when k > 32
    uint64_t size() const { return minim.size(); }
};

// Returns false (with *err) when the header cannot be parsed.
bool decode_sketch(const uint8_t *p, size_t n, SketchElems &out, std::string *err);

// print_containment / print_jaccard (Comparator.cpp:362-408, :412-460).
// inter(i,j) must return |K_i ∩ K_j| for i < j.
void format_csv(const std::vector<std::string> &names, uint32_t query_size, const uint32_t *inter, uint64_t ld,
                bool row_major_full, const std::vector<uint64_t> &sizes, bool jaccard, unsigned precision,
                double min_threshold, std::vector<uint8_t> &out);
// The same text streamed into a gzip file: row blocks are formatted and compressed by `threads` workers and written
// in order as consecutive gzip members, so memory holds one wave of blocks instead of the whole matrix as text
// (10^4 sketches all-vs-all are ~10^8 cells).  text_bytes (may be null) = size of the CSV before compression.
bool write_csv_gz(const std::string &path, const std::vector<std::string> &names, uint32_t query_size, const uint32_t *inter,
                  uint64_t ld, bool row_major_full, const std::vector<uint64_t> &sizes, bool jaccard, unsigned precision,
                  double min_threshold, int threads, uint64_t *text_bytes);

}  // namespace spsp_host
