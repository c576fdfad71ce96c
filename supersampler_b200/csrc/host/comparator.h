// Host mirror of the reference's class Comparator (Comparator.h:17-46):
// getfilesname, compare_sketches(query_size), print_containment, print_jaccard.
// The N-way merge + colour counting runs on the GPU through the C ABI.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <memory>

#include "session.h"
#include "sketchfile.h"

namespace spsp_host {

class Comparator {
public:
    Comparator(unsigned precision, double min_threshold);   // Comparator.h:28-33

    void getfilesname(const std::string &fof, std::vector<std::string> &result);   // Comparator.cpp:7-21
    void compare_sketches(unsigned size_query);                                    // Comparator.cpp:39-74
    void print_containment(const std::string &outfile);                            // Comparator.cpp:362-408
    void print_jaccard(const std::string &outfile);                                // Comparator.cpp:412-460
    // In-memory: sketches as byte strings (after gunzip); names go to the CSV header.
    void compare_buffers(const std::vector<std::string> &names, const std::vector<const uint8_t *> &data,
                         const std::vector<size_t> &len, unsigned size_query);
    void csv(bool jaccard, std::vector<uint8_t> &out) const;

    uint64_t k = 0, m = 0, nb_files = 0, query_size = 0, precision;
    double min_threshold;
    std::vector<std::string> files_names;
    std::vector<uint64_t> nb_kmer_seen_infile;      // |K_i|
    std::vector<uint32_t> score;                    // query_size x nb_files (row i = query i) or upper triangle
    bool full_rows = false;                         // true in query mode
    int n_gpus = 1, n_threads = 0;
    double t_load = 0, t_compare = 0;               // seconds
    float kernel_ms = 0;
    uint64_t launches = 0;

private:
    void run_device(std::vector<SketchElems> &sk);
    std::vector<std::shared_ptr<DeviceSession>> sessions_;   // one per GPU, reused while (k, m) stay the same
    uint64_t sess_k_ = 0, sess_m_ = 0;
};

int comparator_main(int argc, char **argv);          // Comparator.cpp:464-521
int sort_csv_main(int argc, char **argv);            // sort_csv.cpp:115-122 (host-only helper)

}  // namespace spsp_host
