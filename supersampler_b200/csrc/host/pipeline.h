// Batch pipeline of the sketch stage: the GPU-shaped form of the reference's
// `sub_sampler -f` loop (SubSampler.cpp:761-801).  Where the reference gives
// every file to one OpenMP thread that runs the whole per-base loop, here host
// threads only clean + pack their files (getLineFasta / clean_dna,
// utils.cpp:675-718) straight into one pinned staging buffer, each finished
// region is copied to the device asynchronously while the other files are
// still being packed, and ONE scan + device post-pass (include/spsp.h,
// spsp_sketch_batch_staged) produces the sketch bytes of every file.  The
// compare stage can start from the elements the batch left on the device.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "seqio.h"
#include "session.h"

namespace spsp_host {

struct BatchSource {
    const uint8_t *data = nullptr;    // FASTA text in memory ...
    size_t len = 0;
    std::string path;                 // ... or a FASTA(.gz) file when data == nullptr
};

struct BatchStats {
    double prep_s = 0, pack_s = 0, device_s = 0, assemble_s = 0;   // wall seconds, summed over batches
    double scan_ms = 0, post_ms = 0;                               // CUDA events
    uint64_t hits = 0, elems = 0, bases = 0, batches = 0, h2d_bytes = 0, d2h_bytes = 0;
    uint64_t text_inputs = 0;                                      // inputs that went to the device as raw text
    double ingest_ms = 0;                                          // their clean + pack kernels (CUDA events)
};

// Who cleans + packs an input (getLineFasta / clean_dna, utils.cpp:675-718):
//   HOST    host threads (AVX2 / AVX-512 packer) -> 2-bit words -> H2D (0.25 B per base over PCIe)
//   DEVICE  raw text -> H2D -> ingest kernels (csrc/device/ingest.cu); the host only moves bytes
//   AUTO    both at once on one work queue: the workers pack inputs from the front; whenever one of them finds a
//           text lane idle (the previous raw input has arrived) it first sends an input from the back of the
//           queue as raw text, so the split follows what the box can do (few cores per GPU: about half of the
//           inputs go over as text).  With more than 11 workers the queue is host-only: the host lane then runs
//           into the host's memory bandwidth and raw text only competes for the PCIe link (measured, see
//           pipeline.cpp).  Inputs that look like read sets (a header line every few hundred bytes) always go over
//           as text: the host packer crawls on them, the ingest kernels do not care.  The default.
enum class Ingest { HOST = 0, DEVICE = 1, AUTO = 2 };

// Persistent worker threads (the pack phase runs every few milliseconds: no thread start-up per batch).
class WorkerPool {
public:
    explicit WorkerPool(int threads);
    ~WorkerPool();
    WorkerPool(const WorkerPool &) = delete;
    WorkerPool &operator=(const WorkerPool &) = delete;
    // Runs fn(i) for i in [0, n); the caller takes part; rethrows the first exception.
    void run(size_t n, const std::function<void(size_t)> &fn);
    int size() const { return n_threads_; }

private:
    struct Impl;
    Impl *impl_;
    int n_threads_;
};

class BatchSketcher {
public:
    // Uses slot 0 of `session` (the compare stage's stream, so the hand-off needs no extra sync).
    BatchSketcher(std::shared_ptr<DeviceSession> session, int k, int m, double s, unsigned abundance, int threads);
    ~BatchSketcher();
    BatchSketcher(const BatchSketcher &) = delete;
    BatchSketcher &operator=(const BatchSketcher &) = delete;

    // Sketch bytes (before gzip) of every source, in order.  ok[i] = 0 for a file that
    // cannot be opened (its sketch stays empty), like SubSampler.cpp:313-322.
    void run(const std::vector<BatchSource> &src, std::vector<std::vector<uint8_t>> &sketches, std::vector<char> &ok);
    // The same in two halves, so that a caller can overlap them across jobs (another BatchSketcher can pack the
    // next job while this one's device phase runs): begin() prepares, packs and queues the copies and returns
    // as soon as the host work is done; finish() runs the device phase and delivers what run() delivers.
    // `src` must stay valid until begin() returns.  A job that needs several batches is run entirely by begin().
    void begin(const std::vector<BatchSource> &src);
    void finish(std::vector<std::vector<uint8_t>> &sketches, std::vector<char> &ok);

    // All-vs-all (query_size >= n) or query-vs-all compare of the sketches of the last
    // run(), starting from the elements the batch left on the device (several batches:
    // from their host copies).  inter: rows x n, sizes: n (Comparator.cpp:39-74 semantics).
    void compare_last(unsigned query_size, std::vector<uint32_t> &inter, std::vector<uint64_t> &sizes, bool &full_rows,
                      float *kernel_ms);

    // Element offsets (n + 1 entries) of the last run(); on the device only when it was one batch.
    const std::vector<uint64_t> &elem_off() const { return elem_off_; }
    bool elems_on_device() const { return elems_on_device_; }

    BatchStats stats;                          // last run()
    uint64_t max_batch_bases = 1ull << 30;     // upper bound of one batch (bases incl. padding)
    uint64_t max_batch_occurrences = 1ull << 27;   // ... and of its expected selected k-mer occurrences
    std::vector<uint64_t> selected;            // header field 3 of every source, last run()
    // print_stat totals that need every k-mer (SubSampler.cpp:633-665): filled by run() when dense_stats is
    // set, by the dense minimizer machine on the batch still staged on the device (spsp_dense_stats_staged).
    bool dense_stats = false;
    Ingest ingest = Ingest::AUTO;              // or what SPSP_INGEST (host|device|auto) names
    std::vector<uint64_t> total_kmers, total_superkmers;
    double dense_ms = 0;

private:
    struct Prepared;
    struct Job;
    void pack_batch(const std::vector<BatchSource> &src, std::vector<Prepared> &prep, size_t first, size_t last);
    void device_batch(std::vector<Prepared> &prep, size_t first, size_t last, std::vector<std::vector<uint8_t>> &sketches,
                      bool keep_host_elems);
    std::unique_ptr<Job> job_;                 // between begin() and finish()
    std::shared_ptr<DeviceSession> session_;
    WorkerPool pool_;
    int k_, m_, threads_;
    double s_;
    unsigned abundance_;
    bool dbg_no_upload_ = false;
    uint32_t *stage_ = nullptr;                // pinned staging buffer (grow-only)
    uint64_t stage_words_ = 0;
    // elements of the last run
    bool elems_on_device_ = false;
    uint32_t n_last_ = 0;
    std::vector<uint64_t> elem_off_;
    std::vector<uint32_t> h_minim_;
    std::vector<uint64_t> h_klo_, h_khi_;
};

// Runs fn(i) for i in [0, n) on up to `threads` host threads (the caller is one of them).
void parallel_for(int threads, size_t n, const std::function<void(size_t)> &fn);

}  // namespace spsp_host
