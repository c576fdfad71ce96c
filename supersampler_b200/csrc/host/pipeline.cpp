#include "pipeline.h"

#include "postpass.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <functional>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace spsp_host {

using clk = std::chrono::steady_clock;
static double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

void parallel_for(int threads, size_t n, const std::function<void(size_t)> &fn)
{
    if (n == 0) return;
    size_t nw = (size_t)(threads < 1 ? 1 : threads);
    if (nw > n) nw = n;
    std::atomic<size_t> next{0};
    std::mutex mu;
    std::string error;
    auto work = [&]() {
        try {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= n) break;
                fn(i);
            }
        } catch (const std::exception &e) {
            std::lock_guard<std::mutex> g(mu);
            if (error.empty()) error = e.what();
            next.store(n);
        }
    };
    std::vector<std::thread> pool;
    for (size_t w = 1; w < nw; w++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (!error.empty()) throw std::runtime_error(error);
}

struct WorkerPool::Impl {
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> threads;
    const std::function<void(size_t)> *fn = nullptr;
    size_t n = 0;
    std::atomic<size_t> next{0};
    uint64_t generation = 0;
    int active = 0;
    bool stop = false;
    std::string error;

    void drain()
    {
        try {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= n) break;
                (*fn)(i);
            }
        } catch (const std::exception &e) {
            std::lock_guard<std::mutex> g(mu);
            if (error.empty()) error = e.what();
            next.store(n);
        }
    }
    void worker()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
            }
            drain();
            {
                std::lock_guard<std::mutex> g(mu);
                if (--active == 0) cv_done.notify_all();
            }
        }
    }
};

WorkerPool::WorkerPool(int threads) : impl_(new Impl()), n_threads_(threads < 1 ? 1 : threads)
{
    for (int t = 1; t < n_threads_; t++) impl_->threads.emplace_back([this] { impl_->worker(); });
}

WorkerPool::~WorkerPool()
{
    {
        std::lock_guard<std::mutex> g(impl_->mu);
        impl_->stop = true;
    }
    impl_->cv_work.notify_all();
    for (auto &t : impl_->threads) t.join();
    delete impl_;
}

void WorkerPool::run(size_t n, const std::function<void(size_t)> &fn)
{
    if (n == 0) return;
    Impl &p = *impl_;
    {
        std::lock_guard<std::mutex> g(p.mu);
        p.fn = &fn; p.n = n; p.next.store(0); p.error.clear();
        p.active = (int)p.threads.size();
        p.generation++;
    }
    p.cv_work.notify_all();
    p.drain();
    std::unique_lock<std::mutex> lk(p.mu);
    p.cv_done.wait(lk, [&] { return p.active == 0; });
    if (!p.error.empty()) throw std::runtime_error(p.error);
}

// Two pinned chunks per worker thread for text read from files (device-side ingest).
namespace {
// Does this text look like a read set?  (header lines every few hundred bytes: the host packer's block loops
// fall back to its byte-wise state machine at every '>', 0.6 GB/s per thread against 7-10 on genomes, while the
// ingest kernels do not care: C4's 1 Gbp read sets go 1.7 -> 33 Gbp/s from files)
bool looks_like_reads(const uint8_t *p, size_t n)
{
    n = std::min<size_t>(n, 64u << 10);
    size_t headers = 0;
    for (size_t i = 0; i + 1 < n; i++) headers += (p[i] == '\n') & (p[i + 1] == '>');
    return n >= 4096 && headers * 4096 > n;              // more than one record per 4 KB
}
struct TextRing {
    uint8_t *buf[2] = {nullptr, nullptr};
    size_t cap = 0;
    void ensure(size_t bytes)
    {
        if (bytes <= cap) return;
        for (auto &b : buf) {
            if (b) spsp_host_free(b);
            void *v = nullptr;
            if (spsp_host_alloc(&v, bytes) != 0) throw_spsp("spsp_host_alloc");
            b = static_cast<uint8_t *>(v);
        }
        cap = bytes;
    }
    ~TextRing() { for (auto &b : buf) if (b) spsp_host_free(b); }
};
}  // namespace

struct BatchSketcher::Prepared {
    uint64_t len = 0;                  // upper bound of the number of bases (= text bytes)
    bool ok = true, from_file = false, gz = false;
    std::vector<uint8_t> text;         // inflated gzip input
    // filled by the pack phase
    uint64_t word_off = 0, n_bases = 0;
    std::vector<uint64_t> rec_off;
    // device-side ingest
    bool raw = false;                  // went to the device as text
    bool reads = false;                // short records: AUTO sends it as text
    uint64_t text_off = 0;             // where in the device text buffer (multiple of 16)
};

struct BatchSketcher::Job {
    std::vector<Prepared> prep;
    std::vector<std::vector<uint8_t>> sketches;
    std::vector<char> ok;
    size_t first = 0, last = 0;
    bool device_pending = false;
    uint64_t total_words = 0;
    std::chrono::steady_clock::time_point t_packed;
};

BatchSketcher::BatchSketcher(std::shared_ptr<DeviceSession> session, int k, int m, double s, unsigned abundance, int threads)
    : session_(std::move(session)), pool_(threads), k_(k), m_(m), threads_(threads < 1 ? 1 : threads), s_(s),
      abundance_(abundance)
{
    if (const char *e = getenv("SPSP_INGEST"))
        ingest = !strcmp(e, "device") ? Ingest::DEVICE : !strcmp(e, "auto") ? Ingest::AUTO : Ingest::HOST;
}

BatchSketcher::~BatchSketcher()
{
    if (stage_) spsp_host_free(stage_);
}

static bool file_is_gzip(int fd)
{
    unsigned char mg[2] = {0, 0};
    return pread(fd, mg, 2, 0) == 2 && mg[0] == 0x1f && mg[1] == 0x8b;
}

void BatchSketcher::run(const std::vector<BatchSource> &src, std::vector<std::vector<uint8_t>> &sketches, std::vector<char> &ok)
{
    begin(src);
    finish(sketches, ok);
}

void BatchSketcher::begin(const std::vector<BatchSource> &src)
{
    job_.reset(new Job());
    std::vector<std::vector<uint8_t>> &sketches = job_->sketches;
    std::vector<char> &ok = job_->ok;
    std::vector<Prepared> &prep = job_->prep;
    stats = BatchStats();
    const size_t n = src.size();
    sketches.assign(n, std::vector<uint8_t>());
    ok.assign(n, 1);
    selected.assign(n, 0);
    total_kmers.assign(n, 0);
    total_superkmers.assign(n, 0);
    dense_ms = 0;
    elem_off_.assign(1, 0);
    h_minim_.clear(); h_klo_.clear(); h_khi_.clear();
    elems_on_device_ = false;
    n_last_ = (uint32_t)n;
    if (n == 0) return;

    // ---- prepare: sizes.  A gzip input's size is unknown before it is inflated: its ISIZE trailer (exact for a
    // single-member file below 4 GB) or a multiple of the compressed size stands in until its batch is cut;
    // inflating happens batch by batch, so host memory holds the text of about one batch, not of the whole job
    auto t0 = clk::now();
    prep.assign(n, Prepared());
    pool_.run(n, [&](size_t i) {
        Prepared &p = prep[i];
        if (src[i].data) { p.len = src[i].len; p.reads = looks_like_reads(src[i].data, src[i].len); return; }
        p.from_file = true;
        int fd = open(src[i].path.c_str(), O_RDONLY);
        struct stat st;
        if (fd < 0 || fstat(fd, &st) != 0 || S_ISDIR(st.st_mode)) {
            if (fd >= 0) close(fd);
            p.ok = false;
            return;
        }
        p.len = (uint64_t)st.st_size;
        if (file_is_gzip(fd)) {
            p.gz = true;
            unsigned char t4[4] = {0, 0, 0, 0};
            uint64_t isize = 0;
            if (st.st_size >= 18 && pread(fd, t4, 4, st.st_size - 4) == 4)
                isize = (uint64_t)t4[0] | ((uint64_t)t4[1] << 8) | ((uint64_t)t4[2] << 16) | ((uint64_t)t4[3] << 24);
            p.len = std::max<uint64_t>(isize, 3 * (uint64_t)st.st_size);
        } else {
            uint8_t head[64u << 10];
            const ssize_t r = pread(fd, head, sizeof head, 0);
            p.reads = r > 0 && looks_like_reads(head, (size_t)r);
        }
        close(fd);
    });
    stats.prep_s = secs(t0, clk::now());
    for (size_t i = 0; i < n; i++) ok[i] = prep[i].ok ? 1 : 0;

    // ---- batches of consecutive inputs: bounded by bases and by the expected number of selected k-mer
    // occurrences (the device post-pass keeps ~150 bytes per occurrence: dense sampling, -s close to 1, must
    // not turn a big batch into an allocation the GPU cannot serve)
    const double p_hit = (double)compute_threshold(k_, m_, s_) / 18446744073709551616.0;
    const double occ_per_base = std::min(1.0, 1.5 * (double)(k_ - m_ + 1) * p_hit);
    const uint64_t occ_limit_bases = (uint64_t)std::min(9.0e18, (double)max_batch_occurrences / std::max(occ_per_base, 1e-12));
    const uint64_t batch_limit = std::max<uint64_t>(4096, std::min(max_batch_bases, occ_limit_bases));
    auto cut = [&](size_t first) {                              // end of the batch that starts at `first`
        uint64_t acc = 0;
        size_t i = first;
        for (; i < n; i++) {
            const uint64_t need = 16 * spsp_packed_words(prep[i].ok ? prep[i].len : 0);
            if (i > first && acc + need > batch_limit) break;
            acc += need;
        }
        return i;
    };
    size_t first = 0;
    while (first < n) {
        size_t last = cut(first);
        // inflate this batch's gzip inputs (now their sizes are exact), then cut again
        auto ti = clk::now();
        pool_.run(last - first, [&](size_t j) {
            Prepared &p = prep[first + j];
            if (!p.ok || !p.gz) return;
            p.gz = false;
            if (!read_file_maybe_gz(src[first + j].path, p.text)) { p.ok = false; p.len = 0; return; }
            p.from_file = false;
            p.len = p.text.size();
            p.reads = looks_like_reads(p.text.data(), p.text.size());
        });
        stats.prep_s += secs(ti, clk::now());
        for (size_t i = first; i < last; i++) ok[i] = prep[i].ok ? 1 : 0;
        last = std::min(last, cut(first));                      // (inputs cut off here keep their text for the next batch)
        stats.batches++;
        if (first == 0 && last == n) {
            // the usual case: host half now, device half in finish()
            pack_batch(src, prep, 0, n);
            job_->first = 0; job_->last = n;
            job_->device_pending = true;
            return;
        }
        pack_batch(src, prep, first, last);
        device_batch(prep, first, last, sketches, true);
        first = last;
    }
}

void BatchSketcher::finish(std::vector<std::vector<uint8_t>> &sketches, std::vector<char> &ok)
{
    if (!job_) throw std::runtime_error("BatchSketcher::finish without begin");
    if (job_->device_pending) {
        device_batch(job_->prep, job_->first, job_->last, job_->sketches, false);
        elems_on_device_ = true;
    }
    sketches.swap(job_->sketches);
    ok.swap(job_->ok);
    job_.reset();
}

void BatchSketcher::pack_batch(const std::vector<BatchSource> &src, std::vector<Prepared> &prep, size_t first, size_t last)
{
    spsp_ctx *ctx = session_->ctx();
    const size_t nb = last - first;
    dbg_no_upload_ = getenv("SPSP_PIPE_NO_UPLOAD") != nullptr;      // measurement aid: pack only (results are garbage)
    auto t0 = clk::now();
    // AUTO with many pack workers is HOST: measured on B200 boxes, 16 workers already run into the host's memory
    // bandwidth and raw text then only competes for the PCIe link (96.7 vs 95.3 Gbp/s); with 4 workers per GPU
    // (8 GPUs on a 32-core box) the mixed queue takes 33 to 63 Gbp/s.  SPSP_AUTO_MAX_WORKERS moves the switch.
    static const int auto_max_workers = getenv("SPSP_AUTO_MAX_WORKERS") ? atoi(getenv("SPSP_AUTO_MAX_WORKERS")) : 11;
    Ingest mode = ingest;
    const bool auto_mixed = mode == Ingest::AUTO && pool_.size() <= auto_max_workers;     // opportunistic raw sends
    bool any_reads = false;
    if (mode == Ingest::AUTO)
        for (size_t i = first; i < last; i++) any_reads = any_reads || (prep[i].ok && prep[i].reads);
    if (mode == Ingest::AUTO && !auto_mixed && !any_reads) mode = Ingest::HOST;
    uint64_t total_words = 0, total_text = 0;
    for (size_t i = first; i < last; i++) {
        Prepared &p = prep[i];
        p.word_off = total_words;
        p.raw = false;
        total_words += spsp_packed_words(p.ok ? p.len : 0);
        p.text_off = total_text;
        if (p.ok) {
            total_text += (p.len + 15) & ~(uint64_t)15;

        }
    }
    const uint64_t n_total = 16 * total_words;               // one scan covers every region
    const uint64_t need_words = spsp_packed_words(n_total);
    if (mode != Ingest::DEVICE && need_words > stage_words_) {
        if (stage_) spsp_host_free(stage_);
        stage_ = nullptr; stage_words_ = 0;
        void *v = nullptr;
        const uint64_t cap = need_words + need_words / 8;
        if (spsp_host_alloc(&v, cap * sizeof(uint32_t)) != 0) throw_spsp("spsp_host_alloc");
        stage_ = static_cast<uint32_t *>(v);
        stage_words_ = cap;
    }
    if (spsp_batch_reserve(ctx, 0, need_words) != 0) throw_spsp("spsp_batch_reserve");
    if (mode != Ingest::HOST) {
        if (spsp_batch_text_reserve(ctx, 0, total_text) != 0) throw_spsp("spsp_batch_text_reserve");
    }

    // AUTO: called by the pack workers between slices -- a copy lane with nothing queued means the link has room:
    // send an input from the back of the queue as raw text (asynchronous; the worker goes on packing)
    std::mutex qmu;
    size_t q_lo = 0, q_hi = nb;
    const std::vector<size_t> *q_map = nullptr;          // queue position -> input of the batch
    std::atomic<int> lane_claim[2];
    lane_claim[0] = 0; lane_claim[1] = 0;
    std::function<void(size_t, int)> send_input_fn;
    auto feed_link = [&]() {
        if (mode != Ingest::AUTO || !auto_mixed) return;
        for (int lane = 0; lane < 2; lane++) {
            int idle = 0;
            if (lane_claim[lane].load(std::memory_order_relaxed)) continue;
            if (spsp_batch_upload_idle(ctx, 0, lane, &idle) != 0) throw_spsp("spsp_batch_upload_idle");
            if (!idle || lane_claim[lane].exchange(1)) continue;
            size_t j = 0;
            bool got = false;
            {
                // (text that still sits in a file would have to be read into pinned memory by this worker first,
                // which costs about what packing it costs: files go over as text in DEVICE mode only)
                std::lock_guard<std::mutex> g(qmu);
                if (q_lo < q_hi && !prep[first + (*q_map)[q_hi - 1]].from_file) { j = (*q_map)[--q_hi]; got = true; }
            }
            if (got) send_input_fn(j, lane);
            lane_claim[lane].store(0);
        }
    };
    // ---- host lane: clean + pack one whole input into its region and queue the copy
    auto pack_input = [&](size_t j) {
        Prepared &p = prep[first + j];
        const BatchSource &sc = src[first + j];
        PackedInput in(false);
        in.words.attach(stage_ + p.word_off, spsp_packed_words(p.ok ? p.len : 0));
        uint64_t uploaded = 0;                              // words of this region already queued for the copy
        auto upload_to = [&](uint64_t words) {
            if (words <= uploaded) return;
            if (dbg_no_upload_) { uploaded = words; return; }
            if (spsp_batch_upload(ctx, 0, p.word_off + uploaded, stage_ + p.word_off + uploaded, words - uploaded) != 0)
                throw_spsp("spsp_batch_upload");
            uploaded = words;
        };
        {
            // feed in slices: what is final after a slice goes to the device while the next one is packed
            const size_t SLICE = 1u << 20;
            FastaPacker pk(in, (uint32_t)k_);
            if (!p.ok) {
            } else if (!p.from_file) {
                const uint8_t *d = sc.data ? sc.data : p.text.data();
                const size_t len = sc.data ? sc.len : p.text.size();
                for (size_t off = 0; off < len; off += SLICE) {
                    pk.feed(d + off, std::min(SLICE, len - off));
                    if (off + SLICE < len) upload_to(pk.commit());
                    feed_link();
                }
            } else {
                int fd = open(sc.path.c_str(), O_RDONLY);
                if (fd < 0) throw std::runtime_error("cannot reopen " + sc.path);
                std::vector<uint8_t> buf(SLICE);
                uint64_t got = 0;
                while (got < p.len) {                        // never more than the size the region was cut for
                    ssize_t r = read(fd, buf.data(), (size_t)std::min<uint64_t>(buf.size(), p.len - got));
                    if (r <= 0) break;
                    pk.feed(buf.data(), (size_t)r);
                    got += (uint64_t)r;
                    if (got < p.len) upload_to(pk.commit());
                    feed_link();
                }
                close(fd);
            }
            pk.finish();
        }
        p.n_bases = in.n_bases;
        p.rec_off.swap(in.rec_off);
        upload_to(spsp_packed_words(p.n_bases));
        in.words.detach();
        std::vector<uint8_t>().swap(p.text);
    };
    // ---- device lane, text in a file: bytes [off0, off0 + len0) pass through two pinned chunks of this worker, one
    // per text lane; chunk c is refilled once the copies queued on its lane have left (no staging buffer of the
    // size of the file: pinned memory is slow to allocate, 13 MB were measured at 16 ms).  Pieces of one file are
    // independent (the device needs no sequential state from the host), so a large file is read by many workers.
    auto send_file_piece = [&](size_t j, uint64_t off0, uint64_t len0) {
        Prepared &p = prep[first + j];
        const size_t SLICE = 4u << 20;
        int fd = open(src[first + j].path.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("cannot reopen " + src[first + j].path);
        static thread_local TextRing ring;
        ring.ensure(SLICE);
        uint64_t got = 0;
        for (int c = 0; got < len0; c ^= 1) {
            if (spsp_batch_upload_wait(ctx, 0, c) != 0) { close(fd); throw_spsp("spsp_batch_upload_wait"); }
            const ssize_t r = pread(fd, ring.buf[c], (size_t)std::min<uint64_t>(SLICE, len0 - got), (off_t)(off0 + got));
            if (r <= 0) { close(fd); throw std::runtime_error(src[first + j].path + " changed while it was read"); }
            if (spsp_batch_text_upload(ctx, 0, c, p.text_off + off0 + got, ring.buf[c], (uint64_t)r) != 0) {
                close(fd);
                throw_spsp("spsp_batch_text_upload");
            }
            got += (uint64_t)r;
        }
        close(fd);
    };
    // ---- device lane: the raw text of one input goes to the device text buffer (copy lane `lane`)
    auto send_input = [&](size_t j, int lane) {
        Prepared &p = prep[first + j];
        const BatchSource &sc = src[first + j];
        p.raw = true;
        if (!p.ok || !p.len) return;
        const size_t SLICE = 4u << 20;
        if (!p.from_file) {
            const uint8_t *d = sc.data ? sc.data : p.text.data();
            for (uint64_t off = 0; off < p.len; off += SLICE)
                if (spsp_batch_text_upload(ctx, 0, lane, p.text_off + off, d + off, std::min<uint64_t>(SLICE, p.len - off)) != 0)
                    throw_spsp("spsp_batch_text_upload");
            if (!sc.data) {                                  // inflated text: must outlive the copy
                if (spsp_batch_upload_wait(ctx, 0, lane) != 0) throw_spsp("spsp_batch_upload_wait");
                std::vector<uint8_t>().swap(p.text);
            }
            return;
        }
        send_file_piece(j, 0, p.len);
    };

    send_input_fn = send_input;
    if (mode == Ingest::HOST) {
        pool_.run(nb, pack_input);
    } else if (mode == Ingest::DEVICE) {
        // work items: inputs in memory as they are, files in pieces of 32 MB
        struct Item { size_t j; uint64_t off, len; bool file; };
        std::vector<Item> items;
        const uint64_t PIECE = 32u << 20;
        for (size_t j = 0; j < nb; j++) {
            Prepared &p = prep[first + j];
            if (p.ok && p.from_file && p.len) {
                p.raw = true;
                for (uint64_t off = 0; off < p.len; off += PIECE) items.push_back({j, off, std::min(PIECE, p.len - off), true});
            } else {
                items.push_back({j, 0, 0, false});
            }
        }
        pool_.run(items.size(), [&](size_t i) {
            if (items[i].file) send_file_piece(items[i].j, items[i].off, items[i].len);
            else send_input(items[i].j, -1);
        });
    } else {
        // inputs that look like read sets go over as text (files in pieces, like DEVICE); the others through one
        // queue with two ends: every worker packs inputs from the front; raw sends take inputs from the back
        struct Item { size_t j; uint64_t off, len; bool file; };
        std::vector<Item> items;
        std::vector<size_t> rest;
        const uint64_t PIECE = 32u << 20;
        for (size_t j = 0; j < nb; j++) {
            Prepared &p = prep[first + j];
            if (!(p.ok && p.reads && p.len)) { rest.push_back(j); continue; }
            if (p.from_file) {
                p.raw = true;
                for (uint64_t off = 0; off < p.len; off += PIECE) items.push_back({j, off, std::min(PIECE, p.len - off), true});
            } else {
                items.push_back({j, 0, 0, false});
            }
        }
        q_lo = 0; q_hi = rest.size();
        q_map = &rest;
        std::atomic<size_t> next_item{0};
        pool_.run((size_t)std::max(1, pool_.size()), [&](size_t) {
            for (;;) {
                const size_t i = next_item.fetch_add(1);
                if (i >= items.size()) break;
                if (items[i].file) send_file_piece(items[i].j, items[i].off, items[i].len);
                else send_input(items[i].j, -1);
            }
            for (;;) {
                feed_link();
                size_t j;
                {
                    std::lock_guard<std::mutex> g(qmu);
                    if (q_lo >= q_hi) break;
                    j = rest[q_lo++];
                }
                pack_input(j);
            }
        });
    }
    auto t1 = clk::now();
    stats.pack_s += secs(t0, t1);
    job_->total_words = total_words;
    job_->t_packed = t1;
}

void BatchSketcher::device_batch(std::vector<Prepared> &prep, size_t first, size_t last,
                                 std::vector<std::vector<uint8_t>> &sketches, bool keep_host_elems)
{
    spsp_ctx *ctx = session_->ctx();
    const size_t nb = last - first;
    const uint64_t total_words = job_->total_words, n_total = 16 * total_words;
    auto t1 = clk::now();                                 // (not t_packed: the caller may have waited in between)

    // ---- inputs that went over as text: clean + pack them on the device (their records stay there)
    {
        std::vector<uint64_t> t_off, t_len, w_off, nb_out;
        std::vector<uint32_t> idx;
        for (size_t i = first; i < last; i++) {
            const Prepared &p = prep[i];
            if (!p.raw || !p.ok) continue;
            t_off.push_back(p.text_off); t_len.push_back(p.len); w_off.push_back(p.word_off); idx.push_back((uint32_t)(i - first));
        }
        if (!idx.empty()) {
            nb_out.resize(idx.size());
            if (spsp_batch_text_pack(ctx, 0, (uint32_t)idx.size(), t_off.data(), t_len.data(), w_off.data(), idx.data(),
                                     nb_out.data(), nullptr) != 0)
                throw_spsp("spsp_batch_text_pack");
            for (size_t j = 0; j < idx.size(); j++) {
                prep[first + idx[j]].n_bases = nb_out[j];
                stats.h2d_bytes += t_len[j];
            }
            float ms = 0;
            spsp_batch_text_last_ms(ctx, 0, &ms);
            stats.ingest_ms += ms;
            stats.text_inputs += idx.size();
        }
    }
    // ---- records of the host-packed inputs, ascending
    std::vector<uint64_t> rb, re;
    std::vector<uint32_t> ri;
    for (size_t i = first; i < last; i++) {
        const Prepared &p = prep[i];
        const uint64_t base = 16 * p.word_off;
        stats.bases += p.n_bases;
        if (p.raw) continue;
        for (size_t r = 0; r + 1 < p.rec_off.size(); r++) {
            rb.push_back(base + p.rec_off[r]);
            re.push_back(base + p.rec_off[r + 1]);
            ri.push_back((uint32_t)(i - first));
        }
        stats.h2d_bytes += spsp_packed_words(p.n_bases) * 4;
    }
    spsp_batch_result res{};
    if (spsp_sketch_batch_staged(ctx, 0, n_total, rb.data(), re.data(), ri.data(), rb.size(), (uint32_t)nb, abundance_,
                                 &res) != 0)
        throw_spsp("spsp_sketch_batch_staged");
    if (dense_stats) {
        std::vector<uint64_t> tot(nb ? nb : 1), sel(nb ? nb : 1), km(nb ? nb : 1, 0);
        float ms = 0;
        bool any_raw = false;
        for (size_t j = 0; j < nb; j++) any_raw = any_raw || prep[first + j].raw;
        // (a batch with inputs ingested on the device: the record table lives there, merged by the sketch call)
        if (any_raw ? spsp_dense_stats_staged(ctx, 0, n_total, nullptr, nullptr, nullptr, 0, (uint32_t)nb, tot.data(), sel.data(), &ms)
                    : spsp_dense_stats_staged(ctx, 0, n_total, rb.data(), re.data(), ri.data(), rb.size(), (uint32_t)nb, tot.data(),
                                              sel.data(), &ms))
            throw_spsp("spsp_dense_stats_staged");
        if (any_raw && spsp_batch_record_kmers(ctx, 0, (uint32_t)nb, km.data()) != 0) throw_spsp("spsp_batch_record_kmers");
        dense_ms += ms;
        for (size_t j = 0; j < nb; j++) {
            const Prepared &p = prep[first + j];
            total_superkmers[first + j] = tot[j];
            if (any_raw) total_kmers[first + j] += km[j];
            else
                for (size_t r = 0; r + 1 < p.rec_off.size(); r++) total_kmers[first + j] += p.rec_off[r + 1] - p.rec_off[r] - (uint64_t)k_ + 1;
            if (sel[j] != res.selected[j])      // two independent device paths must agree on the selected k-mers
                throw std::runtime_error("dense and sparse sketch paths disagree on the number of selected k-mers");
        }
    }
    auto t2 = clk::now();
    stats.device_s += secs(t1, t2);
    stats.scan_ms += res.scan_ms; stats.post_ms += res.post_ms;
    stats.hits += res.n_hits; stats.elems += res.n_elems;
    stats.h2d_bytes += rb.size() * 20;
    stats.d2h_bytes += res.body_off[nb] + nb * 24;

    // ---- header line + body (SubSampler.cpp:459-460)
    for (size_t j = 0; j < nb; j++) {
        if (!prep[first + j].ok) continue;
        char hdr[96];
        const int hl = snprintf(hdr, sizeof hdr, "%d %d %llu %f\n", 2 * k_ - m_, m_, (unsigned long long)res.selected[j], s_);
        std::vector<uint8_t> &o = sketches[first + j];
        const uint64_t b0 = res.body_off[j], b1 = res.body_off[j + 1];
        o.resize((size_t)hl + (size_t)(b1 - b0));
        memcpy(o.data(), hdr, (size_t)hl);
        if (b1 > b0) memcpy(o.data() + hl, res.body + b0, (size_t)(b1 - b0));
        selected[first + j] = res.selected[j];
    }
    const uint64_t e0 = elem_off_.back();
    for (size_t j = 0; j < nb; j++) elem_off_.push_back(e0 + res.elem_off[j + 1]);
    if (keep_host_elems && res.n_elems) {
        const size_t old = h_minim_.size(), ne = (size_t)res.n_elems;
        h_minim_.resize(old + ne); h_klo_.resize(old + ne);
        if (k_ > 32) h_khi_.resize(old + ne);
        if (spsp_batch_elements(ctx, 0, h_minim_.data() + old, h_klo_.data() + old, k_ > 32 ? h_khi_.data() + old : nullptr,
                                nullptr, nullptr, nullptr) != 0)
            throw_spsp("spsp_batch_elements");
        stats.d2h_bytes += ne * (k_ > 32 ? 20 : 12);
    }
    stats.assemble_s += secs(t2, clk::now());
}

void BatchSketcher::compare_last(unsigned query_size, std::vector<uint32_t> &inter, std::vector<uint64_t> &sizes,
                                 bool &full_rows, float *kernel_ms)
{
    spsp_ctx *ctx = session_->ctx();
    const uint32_t n = n_last_;
    sizes.assign(n, 0);
    for (uint32_t i = 0; i < n; i++) sizes[i] = elem_off_[i + 1] - elem_off_[i];
    if (query_size > n) query_size = n;
    full_rows = query_size < n;
    const uint32_t rows = full_rows ? query_size : n;
    inter.assign((size_t)rows * n, 0);
    if (kernel_ms) *kernel_ms = 0;
    if (rows == 0) return;
    if (elems_on_device_) {
        if (spsp_cmp_load_batch(ctx, 0) != 0) throw_spsp("spsp_cmp_load_batch");
    } else {
        const uint32_t one32 = 0; const uint64_t one64 = 0;
        if (spsp_cmp_load(ctx, n, elem_off_.data(), h_minim_.empty() ? &one32 : h_minim_.data(),
                          h_klo_.empty() ? &one64 : h_klo_.data(), k_ > 32 ? (h_khi_.empty() ? &one64 : h_khi_.data()) : nullptr) != 0)
            throw_spsp("spsp_cmp_load");
    }
    if (spsp_cmp_run(ctx, 0, rows, 0, n, full_rows ? 0 : 1, 0, 1, inter.data(), n) != 0) throw_spsp("spsp_cmp_run");
    if (kernel_ms) spsp_cmp_last_kernel_ms(ctx, kernel_ms);
}

}  // namespace spsp_host
