// extern "C" wrappers of the host layer (include/spsp_host.h).
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "comparator.h"
#include "postpass.h"
#include "seqio.h"
#include "sketchfile.h"
#include "spsp_host.h"
#include "subsampler.h"

using namespace spsp_host;

static thread_local std::string g_herr;
static int hfail(const std::string &m) { g_herr = m; return -1; }

template <class T>
static T *dup_vec(const T *p, size_t n)
{
    T *o = static_cast<T *>(malloc((n ? n : 1) * sizeof(T)));
    if (n) memcpy(o, p, n * sizeof(T));
    return o;
}

extern "C" const char *spsph_last_error(void) { return g_herr.c_str(); }
extern "C" void spsph_free(void *p) { free(p); }
extern "C" int spsph_sub_sampler_main(int argc, char **argv) { return sub_sampler_main(argc, argv); }
extern "C" int spsph_comparator_main(int argc, char **argv) { return comparator_main(argc, argv); }
extern "C" uint64_t spsph_threshold(int k, int m, double s) { return compute_threshold(k, m, s); }

extern "C" int spsph_pack_fasta(const uint8_t *fasta, size_t n, uint32_t min_len, uint32_t **words, uint64_t *n_bases,
                                uint64_t **rec_off, uint64_t *n_rec)
{
    try {
        PackedInput in(false);
        pack_fasta_buffer(fasta, n, min_len, in);
        *words = dup_vec(in.words.data(), (size_t)spsp_packed_words(in.n_bases));
        *n_bases = in.n_bases;
        *rec_off = dup_vec(in.rec_off.data(), in.rec_off.size());
        *n_rec = in.rec_off.size() - 1;
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_postpass(const uint32_t *packed, const uint64_t *rec_off, uint64_t n_rec, const spsp_hit *hits,
                              uint64_t n_hits, int k, int m, double s, unsigned abundance, uint8_t **out,
                              size_t *out_len, uint64_t *selected_kmers)
{
    try {
        SketchParams prm;
        prm.k = k; prm.m = m; prm.s = s; prm.abundance = abundance; prm.threshold = compute_threshold(k, m, s);
        std::vector<uint64_t> ro(rec_off, rec_off + n_rec + 1);
        std::vector<spsp_hit> hv(hits, hits + n_hits);
        std::vector<uint8_t> o;
        SketchStats st;
        build_sketch(packed, ro, hv, prm, o, &st);
        *out = dup_vec(o.data(), o.size());
        *out_len = o.size();
        if (selected_kmers) *selected_kmers = st.selected_kmers;
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_decode_sketch(const uint8_t *sketch, size_t n, int *k, int *m, uint64_t *n_elems,
                                   uint32_t **minimizer, uint64_t **kmer_lo, uint64_t **kmer_hi)
{
    try {
        SketchElems se;
        std::string err;
        if (!decode_sketch(sketch, n, se, &err)) return hfail(err);
        *k = se.k; *m = se.m; *n_elems = se.size();
        *minimizer = dup_vec(se.minim.data(), se.minim.size());
        *kmer_lo = dup_vec(se.klo.data(), se.klo.size());
        *kmer_hi = se.k > 32 ? dup_vec(se.khi.data(), se.khi.size()) : nullptr;
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_format_csv(const char *const *names, uint32_t n, uint32_t query_size, const uint32_t *inter,
                                int full_rows, const uint64_t *sizes, int jaccard, unsigned precision,
                                double min_threshold, uint8_t **out, size_t *out_len)
{
    try {
        std::vector<std::string> nm(names, names + n);
        std::vector<uint64_t> sz(sizes, sizes + n);
        std::vector<uint8_t> o;
        format_csv(nm, query_size, inter, n, full_rows != 0, sz, jaccard != 0, precision, min_threshold, o);
        *out = dup_vec(o.data(), o.size());
        *out_len = o.size();
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_sketch_buffers(int device, int k, int m, double s, unsigned abundance, int scan_mode, uint32_t n,
                                    const uint8_t *const *fasta, const size_t *len, int threads, uint8_t **out,
                                    size_t *out_len, double *timings, uint64_t *launches)
{
    try {
        if (threads < 1) threads = 1;
        if ((uint32_t)threads > n) threads = (int)(n ? n : 1);
        auto session = std::make_shared<DeviceSession>(device, k, m, compute_threshold(k, m, s), threads);
        if (scan_mode != SPSP_SCAN_AUTO && spsp_scan_config(session->ctx(), scan_mode) != 0)
            return hfail(std::string("spsp_scan_config: ") + spsp_last_error());
        std::atomic<uint32_t> next{0};
        std::vector<std::string> errors((size_t)threads);
        std::vector<double> tp((size_t)threads, 0), ts((size_t)threads, 0), tq((size_t)threads, 0);
        std::vector<std::thread> pool;
        for (int w = 0; w < threads; w++)
            pool.emplace_back([&, w]() {
                try {
                    Subsampler ss((uint64_t)k, (uint64_t)m, s, (uint64_t)threads, 3, abundance, session, w);
                    std::vector<uint8_t> sk;
                    for (;;) {
                        uint32_t i = next.fetch_add(1);
                        if (i >= n) break;
                        ss.sketch_buffer(fasta[i], len[i], sk);
                        out[i] = dup_vec(sk.data(), sk.size());
                        out_len[i] = sk.size();
                        tp[(size_t)w] += ss.t_pack; ts[(size_t)w] += ss.t_scan; tq[(size_t)w] += ss.t_post;
                    }
                } catch (const std::exception &e) { errors[(size_t)w] = e.what(); }
            });
        for (auto &t : pool) t.join();
        for (const auto &e : errors)
            if (!e.empty()) return hfail(e);
        if (timings) {
            timings[0] = timings[1] = timings[2] = 0;
            for (int w = 0; w < threads; w++) { timings[0] += tp[(size_t)w]; timings[1] += ts[(size_t)w]; timings[2] += tq[(size_t)w]; }
        }
        if (launches) *launches = session->launches();
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_compare_buffers(int n_gpus, uint32_t n, uint32_t query_size, const uint8_t *const *sketch,
                                     const size_t *len, uint32_t *inter, uint64_t *sizes, int *full_rows,
                                     float *kernel_ms, uint64_t *launches)
{
    try {
        Comparator comp(6, 0.0);
        comp.n_gpus = n_gpus;
        std::vector<std::string> names(n);
        std::vector<const uint8_t *> data(sketch, sketch + n);
        std::vector<size_t> ln(len, len + n);
        comp.compare_buffers(names, data, ln, query_size);
        if (!comp.score.empty()) memcpy(inter, comp.score.data(), comp.score.size() * sizeof(uint32_t));
        for (uint32_t i = 0; i < n; i++) sizes[i] = comp.nb_kmer_seen_infile[i];
        if (full_rows) *full_rows = comp.full_rows ? 1 : 0;
        if (kernel_ms) *kernel_ms = comp.kernel_ms;
        if (launches) *launches = comp.launches;
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}
