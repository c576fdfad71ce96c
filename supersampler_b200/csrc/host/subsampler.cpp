#include "subsampler.h"

#include "pipeline.h"

#include <getopt.h>
#include <sys/stat.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace spsp_host {

using clk = std::chrono::steady_clock;
static double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

std::string get_out_name(const std::string &path, const std::string &prefix)
{
    size_t begin = path.find_last_of('/');
    begin = begin == std::string::npos ? 0 : begin + 1;
    size_t dot = path.find('.', begin);
    return prefix + path.substr(begin, dot == std::string::npos ? std::string::npos : dot - begin);
}

Subsampler::Subsampler(uint64_t ik, uint64_t im, double rate, uint64_t cores, unsigned itype, unsigned iabundance,
                       std::shared_ptr<DeviceSession> session, int slot)
    : k(ik), minimizer_size(im), coreNumber(cores), abundance(iabundance), subsampling_rate(rate), type(itype),
      session_(std::move(session)), slot_(slot), input_(true)
{
    selection_threshold = compute_threshold(rate);
}

Subsampler::~Subsampler() = default;

uint64_t Subsampler::compute_threshold(double sampling_rate)
{
    return spsp_host::compute_threshold((int)k, (int)minimizer_size, sampling_rate);
}

void Subsampler::sketch_packed(std::vector<uint8_t> &sketch)
{
    auto t0 = clk::now();
    spsp_ctx *ctx = session_->ctx();
    if (spsp_scan_submit(ctx, slot_, input_.words.data(), input_.n_bases) != 0) throw_spsp("spsp_scan_submit");
    uint64_t n = 0;
    if (hits_.size() < 4096) hits_.resize(4096);
    int rc = spsp_scan_collect(ctx, slot_, hits_.data(), hits_.size(), &n);
    if (rc == -2) {
        hits_.resize(n);
        rc = spsp_scan_collect(ctx, slot_, hits_.data(), hits_.size(), &n);
    }
    if (rc != 0) throw_spsp("spsp_scan_collect");
    std::vector<spsp_hit> hv(hits_.begin(), hits_.begin() + (ptrdiff_t)n);
    auto t1 = clk::now();
    SketchParams prm;
    prm.k = (int)k; prm.m = (int)minimizer_size; prm.s = subsampling_rate;
    prm.threshold = selection_threshold; prm.abundance = (unsigned)abundance;
    sketch.clear();
    build_sketch(input_.words.data(), input_.rec_off, hv, prm, sketch, &stats);
    auto t2 = clk::now();
    t_scan = secs(t0, t1);
    t_post = secs(t1, t2);
    total_superkmers = 0;
    if (want_dense_stats) {
        // print_stat's totals over every k-mer: dense minimizer machine on the buffer the scan left on the device
        std::vector<uint64_t> rb(input_.rec_off.begin(), input_.rec_off.end() - 1), re(input_.rec_off.begin() + 1, input_.rec_off.end());
        std::vector<uint32_t> ri(rb.size(), 0);
        uint64_t tot = 0, sel = 0;
        if (spsp_dense_stats_staged(ctx, slot_, input_.n_bases, rb.data(), re.data(), ri.data(), rb.size(), 1, &tot, &sel,
                                    nullptr) != 0)
            throw_spsp("spsp_dense_stats_staged");
        if (sel != stats.selected_kmers)
            throw std::runtime_error("dense and sparse sketch paths disagree on the number of selected k-mers");
        total_superkmers = tot;
    }
}

void Subsampler::sketch_buffer(const uint8_t *fasta, size_t n, std::vector<uint8_t> &sketch)
{
    auto t0 = clk::now();
    pack_fasta_buffer(fasta, n, (uint32_t)k, input_);
    t_pack = secs(t0, clk::now());
    sketch_packed(sketch);
}

void Subsampler::parse_fasta_test(const std::string &input_file, const std::string &output_prefix)
{
    stats = SketchStats();
    auto t0 = clk::now();
    if (!pack_fasta_file(input_file, (uint32_t)k, input_)) {
        std::cout << "Can't open file: " << input_file << std::endl;
        return;
    }
    t_pack = secs(t0, clk::now());
    trace("input packed");
    subsampled_file = get_out_name(input_file, output_prefix) + ".gz";
    std::vector<uint8_t> sketch;
    sketch_packed(sketch);
    trace("scan + post-pass done");
    auto t1 = clk::now();
    if (!write_gz(subsampled_file, sketch.data(), sketch.size(), gzip_level))
        std::cout << "Can't write file: " << subsampled_file << std::endl;
    t_write = secs(t1, clk::now());
}

static std::string with_commas(uint64_t n)
{
    std::string s = std::to_string(n), o;
    for (size_t i = 0; i < s.size(); i++) {
        if (i && (s.size() - i) % 3 == 0) o.push_back(',');
        o.push_back(s[i]);
    }
    return o;
}

void Subsampler::print_stat()
{
    // Totals over every k-mer (total_superkmer_number) come from the dense minimizer
    // machine on the GPU (want_dense_stats); everything else from the exact post-pass.
    if (stats.selected_kmers == 0) {
        std::cout << "No kmer selected ***Crickets noise***" << std::endl;
        return;
    }
    uint64_t total_kmers = stats.bases - stats.records * (k - 1);
    std::cout << "I have seen " << with_commas(total_kmers) << " kmers and I selected "
              << with_commas(stats.selected_kmers) << " kmers" << std::endl;
    std::cout << "After removing duplicate kmers, I selected " << with_commas(stats.distinct_kmers) << " kmers" << std::endl;
    std::cout << "This means a practical subsampling rate of " << (double)total_kmers / stats.selected_kmers
              << " with duplicates" << std::endl;
    std::cout << "This means a practical subsampling rate of " << (double)total_kmers / stats.distinct_kmers
              << " without duplicates" << std::endl;
    if (total_superkmers) {
        std::cout << "I have seen " << with_commas(total_superkmers) << " superkmers and I selected "
                  << with_commas(stats.selected_superkmers) << " superkmers" << std::endl;
        std::cout << "This means a practical subsampling rate of " << (double)total_superkmers / stats.selected_superkmers
                  << " with duplicates" << std::endl;
        std::cout << "This means a mean superkmer size of " << (double)total_kmers / total_superkmers
                  << " kmer per superkmer in the input" << std::endl;
    } else {
        std::cout << "I selected " << with_commas(stats.selected_superkmers) << " superkmers" << std::endl;
    }
    std::cout << "After reconstruction and filtering with abundance, I have selected "
              << with_commas(stats.out_superkmers) << " superkmers" << std::endl;
    std::cout << "This means a mean superkmer size of " << (double)stats.selected_kmers / stats.selected_superkmers
              << " kmer per superkmer with duplicates" << std::endl;
    if (stats.out_superkmers)
        std::cout << "This means a mean superkmer size of " << (double)stats.distinct_kmers / stats.out_superkmers
                  << " kmer per superkmer in the output" << std::endl;
    struct stat sb;
    if (!subsampled_file.empty() && stat(subsampled_file.c_str(), &sb) == 0) {
        std::cout << "Actual output file size is " << with_commas((uint64_t)sb.st_size / 1000) << "KB" << std::endl;
        std::cout << "This mean " << ((double)sb.st_size * 8 / stats.distinct_kmers) << " bits per kmer" << std::endl;
    }
    std::cout << "Minimizer number: " << with_commas(stats.buckets) << std::endl;
    std::cout << "Number of maximal skmer was:       " << with_commas(stats.maximal_superkmers) << std::endl;
    std::cout << "Actual number of maximal skmer is: " << with_commas(stats.out_maximal) << std::endl;
    std::cout << "GPU scan: " << with_commas(stats.hits) << " selected m-mer positions; pack " << t_pack * 1e3
              << " ms, scan " << t_scan * 1e3 << " ms, post-pass " << t_post * 1e3 << " ms, write " << t_write * 1e3
              << " ms" << std::endl;
    std::cout << "\n" << std::endl;
}

int sub_sampler_main(int argc, char **argv)
{
    std::string input, inputfof, output("subsampled_");
    unsigned k = 31, m1 = 11, c = 8, abundance = 1, type = 3, gpus = 1;
    double s = 1000;
    bool verbose = true;
    int ch;
    optind = 1;
    try {
        while ((ch = getopt(argc, argv, "hdg:q:k:m:n:s:t:b:e:f:i:p:v:x:a:")) != -1) {
            switch (ch) {
            case 'i': input = optarg; break;
            case 'f': inputfof = optarg; break;
            case 'k': k = (unsigned)std::stoi(optarg); break;
            case 'm': m1 = (unsigned)std::stoi(optarg); break;
            case 't': c = (unsigned)std::stoi(optarg); break;
            case 's': s = std::stof(optarg); break;          // float on purpose (SubSampler.cpp:699)
            case 'p': output = optarg; break;
            case 'v': verbose = std::stoi(optarg) != 0; break;
            case 'x': type = (unsigned)std::stoi(optarg); break;
            case 'a': abundance = (unsigned)std::stoi(optarg); break;
            case 'g': gpus = (unsigned)std::stoi(optarg); break;
            }
        }
    } catch (const std::exception &e) {
        std::cout << "Bad argument: " << e.what() << std::endl;
        return 1;
    }
    if (input.empty() && inputfof.empty()) {
        std::cout << "Core arguments:" << std::endl
                  << "	-i Input file" << std::endl
                  << "	-f Input file of file" << std::endl
                  << "	-p Output prefix (subsampled)" << std::endl
                  << "	-k Kmer size used  (31) " << std::endl
                  << "	-s Subsampling used  (1000) " << std::endl
                  << "	-t Threads used  (8) " << std::endl
                  << "	-m Minimizer size used  (11, max value is 15) " << std::endl
                  << "	-v Verbose level (1) " << std::endl
                  << "	-a Abundance min (2) " << std::endl
                  << "	-g Number of GPUs (1) " << std::endl;
        return 0;
    }
    if (m1 % 2 == 0) { std::cout << "Minimizer size must be odd" << std::endl; m1++; }
    if (k % 2 == 0) { std::cout << "Kmer size must be odd" << std::endl; k++; }
    if (m1 > 15) { std::cout << "Minimizer size can't be greater than 15." << std::endl; m1 = 15; }
    std::cout << " I use k=" << k << " m=" << m1 << " s=" << s << std::endl;
    std::cout << "Maximal super kmer are of length " << 2 * k - m1 << " or " << k - m1 + 1 << " kmers" << std::endl;
    if (c < 1) c = 1;
    if (gpus < 1) gpus = 1;
    // a process that uses one GPU need not enumerate the others (driver start-up grows with the number of devices)
    if (gpus == 1) setenv("CUDA_VISIBLE_DEVICES", "0", 0);
    try {
        int ndev = 0;
        trace("main: arguments parsed");
        if (spsp_device_count(&ndev) != 0 || ndev == 0) throw std::runtime_error("no CUDA device available");
        trace("device count (driver initialised)");
        if ((int)gpus > ndev) gpus = (unsigned)ndev;
        const uint64_t thr = compute_threshold((int)k, (int)m1, s);
        if (!input.empty()) {
            auto session = std::make_shared<DeviceSession>(0, (int)k, (int)m1, thr, 1);
            trace("device context created");
            Subsampler ss(k, m1, s, c, type, abundance, session, 0);
            ss.want_dense_stats = verbose;
            ss.parse_fasta_test(input, output);
            trace("sketch written");
            if (verbose) ss.print_stat();
            return 0;
        }
        std::vector<std::string> files;
        {
            std::vector<uint8_t> raw;
            if (!read_file_maybe_gz(inputfof, raw)) {
                std::cout << "Can't open file of file " << inputfof << std::endl;
                return 0;
            }
            size_t b = 0;
            while (b <= raw.size()) {
                size_t e = b;
                while (e < raw.size() && raw[e] != '\n') e++;
                if (e - b > 3) files.emplace_back(raw.begin() + (ptrdiff_t)b, raw.begin() + (ptrdiff_t)e);   // :780
                b = e + 1;
            }
        }
        {
            std::ofstream out_fof(get_out_name(inputfof, output) + ".txt");
            for (const auto &f : files) out_fof << get_out_name(f, output) + ".gz\n";
        }
        // Files are dealt round-robin to the GPUs (sketching shards by input file, no collective);
        // per GPU one batch pipeline: -t host threads pack, one scan + device post-pass per batch.
        for (const auto &f : files) std::cout << f << std::endl;
        const unsigned per_gpu = std::max(1u, c / gpus);
        std::vector<std::string> errors(gpus);
        std::mutex cout_mu;
        auto gpu_work = [&](unsigned g) {
            try {
                std::vector<size_t> mine;
                for (size_t i = g; i < files.size(); i += gpus) mine.push_back(i);
                if (mine.empty()) return;
                auto session = std::make_shared<DeviceSession>((int)g, (int)k, (int)m1, thr, 1);
                trace("device context created");
                BatchSketcher bs(session, (int)k, (int)m1, s, abundance, (int)per_gpu);
                bs.dense_stats = verbose;
                std::vector<BatchSource> src(mine.size());
                for (size_t j = 0; j < mine.size(); j++) src[j].path = files[mine[j]];
                std::vector<std::vector<uint8_t>> sk;
                std::vector<char> ok;
                bs.run(src, sk, ok);
                trace("batches sketched");
                parallel_for((int)per_gpu, mine.size(), [&](size_t j) {
                    const std::string &f = files[mine[j]];
                    if (!ok[j]) {
                        std::lock_guard<std::mutex> lk(cout_mu);
                        std::cout << "Can't open file: " << f << std::endl;
                        return;
                    }
                    const std::string name = get_out_name(f, output) + ".gz";
                    if (!write_gz(name, sk[j].data(), sk[j].size(), 6)) {
                        std::lock_guard<std::mutex> lk(cout_mu);
                        std::cout << "Can't write file: " << name << std::endl;
                    }
                });
                if (verbose) {
                    std::lock_guard<std::mutex> lk(cout_mu);
                    for (size_t j = 0; j < mine.size(); j++) {
                        if (!ok[j]) continue;
                        std::cout << files[mine[j]] << ":" << std::endl;
                        if (bs.selected[j] == 0) {
                            std::cout << "No kmer selected ***Crickets noise***" << std::endl;
                            continue;
                        }
                        std::cout << "I have seen " << with_commas(bs.total_kmers[j]) << " kmers and I selected "
                                  << with_commas(bs.selected[j]) << " kmers" << std::endl;
                        std::cout << "This means a practical subsampling rate of "
                                  << (double)bs.total_kmers[j] / (double)bs.selected[j] << " with duplicates" << std::endl;
                        std::cout << "I have seen " << with_commas(bs.total_superkmers[j]) << " superkmers" << std::endl;
                        std::cout << "This means a mean superkmer size of "
                                  << (double)bs.total_kmers[j] / (double)bs.total_superkmers[j] << " kmer per superkmer in the input"
                                  << std::endl;
                        std::cout << "Sketch of " << with_commas(sk[j].size()) << " bytes" << std::endl;
                    }
                    const BatchStats &st = bs.stats;
                    std::cout << "GPU " << g << ": " << with_commas(st.bases) << " bases in " << st.batches << " batch(es), "
                              << with_commas(st.hits) << " selected m-mer positions; pack " << st.pack_s * 1e3 << " ms, device "
                              << st.device_s * 1e3 << " ms (scan kernel " << st.scan_ms << " ms, post-pass " << st.post_ms
                              << " ms, dense totals " << bs.dense_ms << " ms)" << std::endl;
                }
            } catch (const std::exception &e) {
                errors[g] = e.what();
            }
        };
        std::vector<std::thread> pool;
        for (unsigned g = 1; g < gpus; g++) pool.emplace_back(gpu_work, g);
        gpu_work(0);
        for (auto &t : pool) t.join();
        trace("sketch files written");
        for (const auto &e : errors)
            if (!e.empty()) throw std::runtime_error(e);
    } catch (const std::exception &e) {
        std::cerr << "sub_sampler: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}

}  // namespace spsp_host
