#include "comparator.h"

#include <getopt.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <thread>

#include "seqio.h"
#include "session.h"

namespace spsp_host {

using clk = std::chrono::steady_clock;
static double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

Comparator::Comparator(unsigned p, double mt) : precision(p), min_threshold(mt) {}

void Comparator::getfilesname(const std::string &fof, std::vector<std::string> &result)
{
    std::vector<uint8_t> raw;
    if (!read_file_maybe_gz(fof, raw)) {
        std::cout << "Can't open " << fof << std::endl;
        return;
    }
    size_t b = 0;
    while (b <= raw.size()) {
        size_t e = b;
        while (e < raw.size() && raw[e] != '\n') e++;
        if (e - b > 2) result.emplace_back(raw.begin() + (ptrdiff_t)b, raw.begin() + (ptrdiff_t)e);   // :17
        b = e + 1;
    }
}

static unsigned worker_count(int requested, size_t jobs)
{
    unsigned n = requested > 0 ? (unsigned)requested : std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    if ((size_t)n > jobs) n = (unsigned)std::max<size_t>(jobs, 1);
    return n;
}

void Comparator::compare_sketches(unsigned size_query)
{
    auto t0 = clk::now();
    // Open + inflate + decode every sketch (the reference keeps all of them open
    // and walks them bucket by bucket; here they are decoded in parallel).
    // Unopenable files are dropped from the comparison like Comparator.cpp:45-50
    // drops them from input_files.
    const size_t n = files_names.size();
    std::vector<SketchElems> sk(n);
    std::vector<char> ok(n, 0);
    std::atomic<size_t> next{0};
    unsigned nw = worker_count(n_threads, n);
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < nw; w++)
        pool.emplace_back([&]() {
            std::vector<uint8_t> raw;
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= n) break;
                if (!read_file_maybe_gz(files_names[i], raw)) continue;
                std::string err;
                if (decode_sketch(raw.data(), raw.size(), sk[i], &err)) ok[i] = 1;
            }
        });
    for (auto &t : pool) t.join();
    std::vector<SketchElems> kept;
    std::vector<std::string> names;
    unsigned q = 0;
    for (size_t i = 0; i < n; i++) {
        if (!ok[i]) {
            std::cout << "Problem with file opening" << std::endl;
            continue;
        }
        if (i < size_query) q++;
        kept.push_back(std::move(sk[i]));
        names.push_back(files_names[i]);
    }
    files_names = names;      // keeps the CSV aligned with the sketches that were compared
    query_size = q;
    t_load = secs(t0, clk::now());
    trace("sketches read + decoded");
    run_device(kept);
    trace("compared");
    std::cout << "kmers evaluated are of length: " << k << " minimizer size is " << m << std::endl;
    std::cout << "Comparisons done" << std::endl;
}

void Comparator::compare_buffers(const std::vector<std::string> &names, const std::vector<const uint8_t *> &data,
                                 const std::vector<size_t> &len, unsigned size_query)
{
    auto t0 = clk::now();
    files_names = names;
    query_size = size_query;
    const size_t n = names.size();
    std::vector<SketchElems> sk(n);
    std::vector<std::string> errs(n);
    std::atomic<size_t> next{0};
    unsigned nw = worker_count(n_threads, n);
    auto work = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n) break;
            std::string err;
            if (!decode_sketch(data[i], len[i], sk[i], &err)) errs[i] = err.empty() ? "undecodable" : err;
        }
    };
    std::vector<std::thread> pool;
    for (unsigned w = 1; w < nw; w++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    for (size_t i = 0; i < n; i++)
        if (!errs[i].empty()) throw std::runtime_error("sketch " + names[i] + ": " + errs[i]);
    t_load = secs(t0, clk::now());
    run_device(sk);
}

void Comparator::run_device(std::vector<SketchElems> &sk)
{
    auto t0 = clk::now();
    const uint32_t n = (uint32_t)sk.size();
    nb_files = n;
    if (query_size > n) query_size = n;
    nb_kmer_seen_infile.assign(n, 0);
    score.clear();
    if (n == 0) return;
    // last header wins, like get_header_info (Comparator.cpp:23-37)
    k = (uint64_t)sk[n - 1].k; m = (uint64_t)sk[n - 1].m;
    for (uint32_t i = 0; i < n; i++)
        if ((uint64_t)sk[i].k != k || (uint64_t)sk[i].m != m)
            throw std::runtime_error("sketches were built with different k/m");
    std::vector<uint64_t> off(n + 1, 0);
    for (uint32_t i = 0; i < n; i++) {
        nb_kmer_seen_infile[i] = sk[i].size();
        off[i + 1] = off[i] + sk[i].size();
    }
    const uint64_t E = off[n];
    const bool hi = k > 32;
    std::vector<uint32_t> minim(E ? E : 1);
    std::vector<uint64_t> klo(E ? E : 1), khi(hi ? (E ? E : 1) : 0);
    for (uint32_t i = 0; i < n; i++) {
        if (!sk[i].size()) continue;
        memcpy(minim.data() + off[i], sk[i].minim.data(), sk[i].size() * 4);
        memcpy(klo.data() + off[i], sk[i].klo.data(), sk[i].size() * 8);
        if (hi) memcpy(khi.data() + off[i], sk[i].khi.data(), sk[i].size() * 8);
        SketchElems().minim.swap(sk[i].minim);
    }
    full_rows = query_size < n;
    const uint32_t rows = full_rows ? (uint32_t)query_size : n;
    score.assign((size_t)rows * n, 0);
    if (rows == 0) return;
    int ndev = 0;
    if (spsp_device_count(&ndev) != 0 || ndev == 0) throw std::runtime_error("no CUDA device available");
    int g = std::min(n_gpus, ndev);
    if (g < 1) g = 1;
    std::vector<std::string> errors((size_t)g);
    std::vector<std::vector<uint32_t>> part((size_t)g);
    std::vector<float> kms((size_t)g, 0.f);
    std::vector<uint64_t> nl((size_t)g, 0);
    if (sess_k_ != k || sess_m_ != m || (int)sessions_.size() != g) {
        sessions_.clear();
        for (int dev = 0; dev < g; dev++) sessions_.push_back(std::make_shared<DeviceSession>(dev, (int)k, (int)m, 0, 1));
        sess_k_ = k; sess_m_ = m;
    }
    std::vector<uint64_t> l0((size_t)g, 0);
    for (int dev = 0; dev < g; dev++) l0[(size_t)dev] = sessions_[(size_t)dev]->launches();
    auto work = [&](int dev) {
        try {
            DeviceSession &session = *sessions_[(size_t)dev];
            if (spsp_cmp_load(session.ctx(), n, off.data(), minim.data(), klo.data(), hi ? khi.data() : nullptr) != 0)
                throw_spsp("spsp_cmp_load");
            std::vector<uint32_t> &out = dev == 0 ? score : part[(size_t)dev];
            if (dev != 0) out.assign((size_t)rows * n, 0);
            if (spsp_cmp_run(session.ctx(), 0, rows, 0, n, full_rows ? 0 : 1, (uint32_t)dev, (uint32_t)g, out.data(), n) != 0)
                throw_spsp("spsp_cmp_run");
            spsp_cmp_last_kernel_ms(session.ctx(), &kms[(size_t)dev]);
            nl[(size_t)dev] = session.launches() - l0[(size_t)dev];
        } catch (const std::exception &e) {
            errors[(size_t)dev] = e.what();
        }
    };
    std::vector<std::thread> pool;
    for (int dev = 1; dev < g; dev++) pool.emplace_back(work, dev);
    work(0);
    for (auto &t : pool) t.join();
    for (const auto &e : errors)
        if (!e.empty()) throw std::runtime_error(e);
    for (int dev = 1; dev < g; dev++)
        for (size_t i = 0; i < score.size(); i++) score[i] += part[(size_t)dev][i];
    kernel_ms = 0;
    launches = 0;
    for (int dev = 0; dev < g; dev++) { kernel_ms = std::max(kernel_ms, kms[(size_t)dev]); launches += nl[(size_t)dev]; }
    t_compare = secs(t0, clk::now());
}

void Comparator::csv(bool jaccard, std::vector<uint8_t> &out) const
{
    format_csv(files_names, (uint32_t)query_size, score.data(), nb_files, full_rows, nb_kmer_seen_infile, jaccard,
               (unsigned)precision, min_threshold, out);
}

void Comparator::print_containment(const std::string &outfile)
{
    std::cout << "Containement index dump " << std::endl;
    const int nt = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    if (!write_csv_gz(outfile, files_names, (uint32_t)query_size, score.data(), nb_files, full_rows, nb_kmer_seen_infile, false,
                      (unsigned)precision, min_threshold, nt, nullptr))
        std::cout << "Can't write " << outfile << std::endl;
}

void Comparator::print_jaccard(const std::string &outfile)
{
    std::cout << "Jackard index dump" << std::endl;
    const int nt = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    if (!write_csv_gz(outfile, files_names, (uint32_t)query_size, score.data(), nb_files, full_rows, nb_kmer_seen_infile, true,
                      (unsigned)precision, min_threshold, nt, nullptr))
        std::cout << "Can't write " << outfile << std::endl;
}

int comparator_main(int argc, char **argv)
{
    std::string inputfof, query, output_name("results");
    unsigned p = 6, gpus = 1, threads = 0;
    double min_threshold = 0;
    int ch;
    optind = 1;
    try {
        while ((ch = getopt(argc, argv, "hdag:q:k:m:n:s:t:b:e:f:i:p:o:")) != -1) {
            switch (ch) {
            case 'f': inputfof = optarg; break;
            case 'q': query = optarg; break;
            case 'p': p = (unsigned)std::stoi(optarg); break;
            case 'm': min_threshold = std::stod(optarg); break;
            case 'o': output_name = optarg; break;
            case 'g': gpus = (unsigned)std::stoi(optarg); break;      // ignored by the reference
            case 't': threads = (unsigned)std::stoi(optarg); break;   // ignored by the reference
            }
        }
    } catch (const std::exception &e) {
        std::cout << "Bad argument: " << e.what() << std::endl;
        return 1;
    }
    if (inputfof.empty()) {
        std::cout << "Core arguments:" << std::endl
                  << "-f Index file of files (mandatory)" << std::endl
                  << "-q Query file of files (\"\" for all versus all comparison of the index)" << std::endl
                  << "Ouput arguments:" << std::endl
                  << "-m Minimum value to be output (0.0)" << std::endl
                  << "-p Required precision to be output in the CSV (6)" << std::endl
                  << "-o output prefix (results)" << std::endl
                  << "-g Number of GPUs (1)" << std::endl;
        return 0;
    }
    if (gpus <= 1) setenv("CUDA_VISIBLE_DEVICES", "0", 0);
    // driver + context start-up (~1 s) runs beside the reading and decoding of the sketch files
    std::thread warm([gpus] {
        int ndev = 0;
        if (spsp_device_count(&ndev) != 0) return;
        for (int d = 0; d < ndev && d < (int)std::max(1u, gpus); d++) spsp_warmup(d);
        trace("devices warm");
    });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{warm};
    try {
        Comparator comp(p, min_threshold);
        comp.n_gpus = (int)gpus;
        comp.n_threads = (int)threads;
        auto start = clk::now();
        trace("main: arguments parsed");
        if (query.empty()) {
            std::cout << "No query file, I will perform a all versus all comparison" << std::endl;
            comp.getfilesname(inputfof, comp.files_names);
            std::cout << "I found " << comp.files_names.size() << " documents" << std::endl;
            comp.compare_sketches((unsigned)comp.files_names.size());
        } else {
            comp.getfilesname(query, comp.files_names);
            unsigned qs = (unsigned)comp.files_names.size();
            std::cout << "I query " << qs << " file(s) against the bank" << std::endl;
            comp.getfilesname(inputfof, comp.files_names);
            comp.compare_sketches(qs);
        }
        auto middle = clk::now();
        std::cout << "Comparisons lasted " << secs(start, middle) << " sec" << std::endl;
        comp.print_containment(output_name + "_containment.csv.gz");
        comp.print_jaccard(output_name + "_jaccard.csv.gz");
        std::cout << "Jaccard output lasted " << secs(middle, clk::now()) << " sec" << std::endl;
        trace("CSV files written");
    } catch (const std::exception &e) {
        std::cerr << "comparator: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}

}  // namespace spsp_host
