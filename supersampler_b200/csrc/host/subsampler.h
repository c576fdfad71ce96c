// Host mirror of the reference's class Subsampler (SubSampler.h:29-105):
// same constructor arguments, parse_fasta_test(input, prefix), print_stat(),
// compute_threshold().  The per-base loop runs on the GPU through the C ABI.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "postpass.h"
#include "seqio.h"
#include "session.h"

namespace spsp_host {

// SubSampler.cpp:171-221 extract_name / get_out_name: directory stripped, name
// cut at the first '.', prefix prepended.
std::string get_out_name(const std::string &path, const std::string &prefix);

class Subsampler {
public:
    // Same argument order as the reference constructor (SubSampler.h:63);
    // `session`/`slot` say which device context and stream this instance uses.
    Subsampler(uint64_t k, uint64_t minimizer_size, double subsampling_rate, uint64_t cores, unsigned type,
               unsigned abundance, std::shared_ptr<DeviceSession> session, int slot);
    ~Subsampler();

    // SubSampler.cpp:306-510.  Writes <prefix><stem>.gz into the CWD; an
    // unopenable input prints a message and returns (reference :313-322).
    void parse_fasta_test(const std::string &input_file, const std::string &output_prefix);
    // In-memory variant: FASTA text -> sketch bytes (before gzip).
    void sketch_buffer(const uint8_t *fasta, size_t n, std::vector<uint8_t> &sketch);
    void print_stat();                                   // SubSampler.cpp:633-665
    uint64_t compute_threshold(double sampling_rate);    // SubSampler.cpp:622-631

    uint64_t k, minimizer_size, coreNumber, selection_threshold, abundance;
    double subsampling_rate;
    unsigned type;
    std::string subsampled_file;
    SketchStats stats;
    double t_pack = 0, t_scan = 0, t_post = 0, t_write = 0;   // seconds, last input
    int gzip_level = 6;
    bool want_dense_stats = false;     // also compute print_stat's totals over every k-mer (dense kernels)
    uint64_t total_superkmers = 0;     // total_superkmer_number of the last input (when want_dense_stats)

private:
    void sketch_packed(std::vector<uint8_t> &sketch);
    std::shared_ptr<DeviceSession> session_;
    int slot_;
    PackedInput input_;
    std::vector<spsp_hit> hits_;
};

// The reference's main() (SubSampler.cpp:667-803) as a callable: same getopt
// string, defaults and messages; -g N (accepted and ignored by the reference's
// getopt string) selects the number of GPUs.  Returns the process exit code.
int sub_sampler_main(int argc, char **argv);

}  // namespace spsp_host
