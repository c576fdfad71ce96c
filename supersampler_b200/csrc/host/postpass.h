// Exact post-pass of the sketch stage: from the sparse hit list of the scan
// kernel to the bytes of the reference's sketch file.
//
//   hits -> replay of the minimizer state machine over hits only
//           (SubSampler.cpp:352-454 incl. regular_minimizer_pos :81-169)
//        -> selected super-k-mer pieces -> oriented k-mers per minimizer
//           bucket in emission order (handle_superkmer :243-302)
//        -> first-occurrence de-duplication with uint8 counts
//        -> greedy reconstruction (:512-620) -> serialisation (:459-504).
//
// Non-hit m-mers can never win against a hit (their hash is > T >= any hit's)
// and pieces whose minimizer is not selected are never written, so replaying
// the machine on hits alone yields the same pieces as the dense loop.
#pragma once
#include <cstdint>
#include <vector>

#include "spsp.h"

namespace spsp_host {

struct SketchParams {
    int k = 31, m = 11;
    double s = 1000.0;           // as parsed by the reference: (double)stof(arg)
    uint64_t threshold = 0;
    unsigned abundance = 1;
};

struct SketchStats {
    uint64_t records = 0, bases = 0;
    uint64_t selected_kmers = 0, selected_superkmers = 0, maximal_superkmers = 0;
    uint64_t buckets = 0, distinct_kmers = 0, out_superkmers = 0, out_maximal = 0;
    uint64_t hits = 0, hits_used = 0;
};

// SubSampler.cpp:622-631 / SubSampler.h:79-83 (x87 long double on purpose).
uint64_t compute_threshold(int k, int m, double s);

// packed/rec_off: output of the FASTA packer (records >= k only).
// hits: scan output on the same packed buffer, any order; hits that straddle a
// record boundary are ignored here.  Appends the sketch bytes (before gzip).
void build_sketch(const uint32_t *packed, const std::vector<uint64_t> &rec_off, std::vector<spsp_hit> &hits,
                  const SketchParams &prm, std::vector<uint8_t> &out, SketchStats *stats);

}  // namespace spsp_host
