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
#include "pipeline.h"
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
extern "C" int spsph_sort_csv_main(int argc, char **argv) { return sort_csv_main(argc, argv); }
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

struct spsph_sketcher {
    std::shared_ptr<DeviceSession> session;
    std::vector<std::unique_ptr<Subsampler>> workers;
    int k, m;
    double s;
};

extern "C" int spsph_sketcher_create(int device, int k, int m, double s, unsigned abundance, int scan_mode, int threads,
                                     spsph_sketcher **out)
{
    try {
        if (threads < 1) threads = 1;
        if (threads > 64) threads = 64;
        auto sk = std::make_unique<spsph_sketcher>();
        sk->k = k; sk->m = m; sk->s = s;
        sk->session = std::make_shared<DeviceSession>(device, k, m, compute_threshold(k, m, s), threads);
        if (scan_mode != SPSP_SCAN_AUTO && spsp_scan_config(sk->session->ctx(), scan_mode) != 0)
            return hfail(std::string("spsp_scan_config: ") + spsp_last_error());
        for (int w = 0; w < threads; w++)
            sk->workers.emplace_back(new Subsampler((uint64_t)k, (uint64_t)m, s, (uint64_t)threads, 3, abundance,
                                                    sk->session, w));
        *out = sk.release();
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_sketcher_destroy(spsph_sketcher *sk)
{
    delete sk;
    return 0;
}

extern "C" spsp_ctx *spsph_sketcher_ctx(spsph_sketcher *sk) { return sk ? sk->session->ctx() : nullptr; }

extern "C" int spsph_sketcher_run(spsph_sketcher *sk, uint32_t n, const uint8_t *const *fasta, const size_t *len,
                                  uint8_t **out, size_t *out_len, double *timings, uint64_t *launches)
{
    try {
        if (!sk) return hfail("null sketcher");
        const int threads = (int)std::min<size_t>(sk->workers.size(), n ? n : 1);
        uint64_t l0 = sk->session->launches();
        std::atomic<uint32_t> next{0};
        std::vector<std::string> errors((size_t)threads);
        std::vector<double> tp((size_t)threads, 0), ts((size_t)threads, 0), tq((size_t)threads, 0);
        std::vector<std::thread> pool;
        auto work = [&](int w) {
            try {
                Subsampler &ss = *sk->workers[(size_t)w];
                std::vector<uint8_t> o;
                for (;;) {
                    uint32_t i = next.fetch_add(1);
                    if (i >= n) break;
                    ss.sketch_buffer(fasta[i], len[i], o);
                    out[i] = dup_vec(o.data(), o.size());
                    out_len[i] = o.size();
                    tp[(size_t)w] += ss.t_pack; ts[(size_t)w] += ss.t_scan; tq[(size_t)w] += ss.t_post;
                }
            } catch (const std::exception &e) { errors[(size_t)w] = e.what(); }
        };
        for (int w = 1; w < threads; w++) pool.emplace_back(work, w);
        work(0);
        for (auto &t : pool) t.join();
        for (const auto &e : errors)
            if (!e.empty()) return hfail(e);
        if (timings) {
            timings[0] = timings[1] = timings[2] = 0;
            for (int w = 0; w < threads; w++) { timings[0] += tp[(size_t)w]; timings[1] += ts[(size_t)w]; timings[2] += tq[(size_t)w]; }
        }
        if (launches) *launches = sk->session->launches() - l0;
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_write_csv_gz(const char *path, const char *const *names, uint32_t n, uint32_t query_size,
                                  const uint32_t *inter, int full_rows, const uint64_t *sizes, int jaccard, unsigned precision,
                                  double min_threshold, int threads, uint64_t *text_bytes)
{
    try {
        std::vector<std::string> nm(names, names + n);
        std::vector<uint64_t> sz(sizes, sizes + n);
        if (!write_csv_gz(path, nm, query_size, inter, n, full_rows != 0, sz, jaccard != 0, precision, min_threshold, threads,
                          text_bytes))
            return hfail(std::string("cannot write ") + path);
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_sketch_buffers(int device, int k, int m, double s, unsigned abundance, int scan_mode, uint32_t n,
                                    const uint8_t *const *fasta, const size_t *len, int threads, uint8_t **out,
                                    size_t *out_len, double *timings, uint64_t *launches)
{
    if (threads < 1) threads = 1;
    if ((uint32_t)threads > n) threads = (int)(n ? n : 1);
    spsph_sketcher *sk = nullptr;
    int rc = spsph_sketcher_create(device, k, m, s, abundance, scan_mode, threads, &sk);
    if (rc) return rc;
    rc = spsph_sketcher_run(sk, n, fasta, len, out, out_len, timings, launches);
    spsph_sketcher_destroy(sk);
    return rc;
}

extern "C" int spsph_postpass_batch(const uint32_t *packed, uint32_t n_inputs, const uint64_t *base_off,
                                    const uint64_t *n_bases, const uint64_t *rec_off, const uint64_t *rec_first,
                                    const spsp_hit *hits, uint64_t n_hits, int k, int m, double s, unsigned abundance,
                                    int threads, uint8_t **out, size_t *out_len)
{
    try {
        // route every hit to its input (inputs are disjoint, ascending ranges)
        std::vector<std::vector<spsp_hit>> per((size_t)n_inputs);
        for (uint64_t h = 0; h < n_hits; h++) {
            const uint64_t p = hits[h].pos;
            uint32_t lo = 0, hi = n_inputs;          // last input with base_off <= p
            while (hi - lo > 1) { uint32_t mid = (lo + hi) / 2; if (base_off[mid] <= p) lo = mid; else hi = mid; }
            if (n_inputs == 0 || p < base_off[lo] || p + (uint64_t)m > base_off[lo] + n_bases[lo]) continue;
            spsp_hit x = hits[h];
            x.pos -= base_off[lo];
            per[lo].push_back(x);
        }
        SketchParams prm;
        prm.k = k; prm.m = m; prm.s = s; prm.abundance = abundance; prm.threshold = compute_threshold(k, m, s);
        if (threads < 1) threads = 1;
        if ((uint32_t)threads > n_inputs) threads = (int)(n_inputs ? n_inputs : 1);
        std::atomic<uint32_t> next{0};
        std::vector<std::string> errors((size_t)threads);
        auto work = [&](int w) {
            try {
                std::vector<uint8_t> o;
                std::vector<uint64_t> ro;
                for (;;) {
                    uint32_t i = next.fetch_add(1);
                    if (i >= n_inputs) break;
                    ro.assign(rec_off + rec_first[i], rec_off + rec_first[i + 1]);
                    o.clear();
                    build_sketch(packed + base_off[i] / 16, ro, per[i], prm, o, nullptr);
                    out[i] = dup_vec(o.data(), o.size());
                    out_len[i] = o.size();
                }
            } catch (const std::exception &e) { errors[(size_t)w] = e.what(); }
        };
        std::vector<std::thread> pool;
        for (int w = 1; w < threads; w++) pool.emplace_back(work, w);
        work(0);
        for (auto &t : pool) t.join();
        for (const auto &e : errors)
            if (!e.empty()) return hfail(e);
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

struct spsph_comparer {
    Comparator comp{6, 0.0};
};

extern "C" int spsph_comparer_create(int n_gpus, int threads, spsph_comparer **out)
{
    try {
        auto c = std::make_unique<spsph_comparer>();
        c->comp.n_gpus = n_gpus < 1 ? 1 : n_gpus;
        c->comp.n_threads = threads;
        *out = c.release();
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_comparer_destroy(spsph_comparer *c)
{
    delete c;
    return 0;
}

extern "C" int spsph_comparer_run(spsph_comparer *c, uint32_t n, uint32_t query_size, const uint8_t *const *sketch,
                                  const size_t *len, uint32_t *inter, uint64_t *sizes, int *full_rows,
                                  float *kernel_ms, uint64_t *launches, double *timings)
{
    try {
        if (!c) return hfail("null comparer");
        Comparator &comp = c->comp;
        std::vector<std::string> names(n);
        std::vector<const uint8_t *> data(sketch, sketch + n);
        std::vector<size_t> ln(len, len + n);
        comp.compare_buffers(names, data, ln, query_size);
        if (!comp.score.empty()) memcpy(inter, comp.score.data(), comp.score.size() * sizeof(uint32_t));
        for (uint32_t i = 0; i < n; i++) sizes[i] = comp.nb_kmer_seen_infile[i];
        if (full_rows) *full_rows = comp.full_rows ? 1 : 0;
        if (kernel_ms) *kernel_ms = comp.kernel_ms;
        if (launches) *launches = comp.launches;
        if (timings) { timings[0] = comp.t_load; timings[1] = comp.t_compare; }
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_compare_buffers(int n_gpus, uint32_t n, uint32_t query_size, const uint8_t *const *sketch,
                                     const size_t *len, uint32_t *inter, uint64_t *sizes, int *full_rows,
                                     float *kernel_ms, uint64_t *launches)
{
    spsph_comparer *c = nullptr;
    int rc = spsph_comparer_create(n_gpus, 0, &c);
    if (rc) return rc;
    rc = spsph_comparer_run(c, n, query_size, sketch, len, inter, sizes, full_rows, kernel_ms, launches, nullptr);
    spsph_comparer_destroy(c);
    return rc;
}

// ------------------------------------------------------------ batch pipeline

struct spsph_pipeline {
    std::shared_ptr<DeviceSession> session;
    std::unique_ptr<BatchSketcher> bs;
    uint64_t l0 = 0;
    uint32_t n_pending = 0;
};

extern "C" int spsph_pipeline_create(int device, int k, int m, double s, unsigned abundance, int threads,
                                     spsph_pipeline **out)
{
    try {
        auto p = std::make_unique<spsph_pipeline>();
        p->session = std::make_shared<DeviceSession>(device, k, m, compute_threshold(k, m, s), 1);
        p->bs.reset(new BatchSketcher(p->session, k, m, s, abundance, threads));
        *out = p.release();
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_pipeline_destroy(spsph_pipeline *p)
{
    delete p;
    return 0;
}

extern "C" spsp_ctx *spsph_pipeline_ctx(spsph_pipeline *p) { return p ? p->session->ctx() : nullptr; }

extern "C" int spsph_pipeline_set_max_batch_bases(spsph_pipeline *p, uint64_t bases)
{
    if (!p || bases < 4096) return hfail("bad arguments");
    p->bs->max_batch_bases = bases;
    return 0;
}

extern "C" int spsph_pipeline_set_ingest(spsph_pipeline *p, int mode)
{
    if (!p || mode < 0 || mode > 2) return hfail("bad arguments");
    p->bs->ingest = mode == 1 ? Ingest::DEVICE : mode == 2 ? Ingest::AUTO : Ingest::HOST;
    return 0;
}

static int pipeline_sources(uint32_t n, const uint8_t *const *fasta, const size_t *len, const char *const *paths,
                            std::vector<BatchSource> &src)
{
    src.assign(n, BatchSource());
    for (uint32_t i = 0; i < n; i++) {
        if (fasta && fasta[i]) { src[i].data = fasta[i]; src[i].len = len[i]; }
        else if (paths && paths[i]) src[i].path = paths[i];
        else return hfail("input without data and without path");
    }
    return 0;
}

static void pipeline_deliver(spsph_pipeline *p, uint32_t n, std::vector<std::vector<uint8_t>> &sk, std::vector<char> &okv,
                             uint8_t **out, size_t *out_len, int *ok, double *stats)
{
    for (uint32_t i = 0; i < n; i++) {
        out[i] = dup_vec(sk[i].data(), sk[i].size());
        out_len[i] = sk[i].size();
        if (ok) ok[i] = okv[i];
    }
    if (stats) {
        const BatchStats &st = p->bs->stats;
        stats[0] = st.prep_s; stats[1] = st.pack_s; stats[2] = st.device_s; stats[3] = st.assemble_s;
        stats[4] = st.scan_ms; stats[5] = st.post_ms; stats[6] = (double)st.hits; stats[7] = (double)st.elems;
        stats[8] = (double)st.h2d_bytes; stats[9] = (double)st.d2h_bytes; stats[10] = (double)st.batches;
        stats[11] = (double)st.bases; stats[12] = (double)st.text_inputs; stats[13] = st.ingest_ms;
    }
}

extern "C" int spsph_pipeline_sketch(spsph_pipeline *p, uint32_t n, const uint8_t *const *fasta, const size_t *len,
                                     const char *const *paths, uint8_t **out, size_t *out_len, int *ok, double *stats,
                                     uint64_t *launches)
{
    try {
        if (!p) return hfail("null pipeline");
        std::vector<BatchSource> src;
        if (pipeline_sources(n, fasta, len, paths, src)) return -1;
        const uint64_t l0 = p->session->launches();
        std::vector<std::vector<uint8_t>> sk;
        std::vector<char> okv;
        p->bs->run(src, sk, okv);
        pipeline_deliver(p, n, sk, okv, out, out_len, ok, stats);
        if (launches) *launches = p->session->launches() - l0;
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_pipeline_pack(spsph_pipeline *p, uint32_t n, const uint8_t *const *fasta, const size_t *len,
                                   const char *const *paths)
{
    try {
        if (!p) return hfail("null pipeline");
        std::vector<BatchSource> src;
        if (pipeline_sources(n, fasta, len, paths, src)) return -1;
        p->l0 = p->session->launches();
        p->n_pending = n;
        p->bs->begin(src);
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_pipeline_finish(spsph_pipeline *p, uint32_t n, uint8_t **out, size_t *out_len, int *ok, double *stats,
                                     uint64_t *launches)
{
    try {
        if (!p) return hfail("null pipeline");
        if (n != p->n_pending) return hfail("spsph_pipeline_finish: n differs from the packed job");
        std::vector<std::vector<uint8_t>> sk;
        std::vector<char> okv;
        p->bs->finish(sk, okv);
        pipeline_deliver(p, n, sk, okv, out, out_len, ok, stats);
        if (launches) *launches = p->session->launches() - p->l0;
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}

extern "C" int spsph_pipeline_elem_off(spsph_pipeline *p, uint64_t *off, int *on_device)
{
    if (!p || !off) return hfail("bad arguments");
    const std::vector<uint64_t> &e = p->bs->elem_off();
    memcpy(off, e.data(), e.size() * sizeof(uint64_t));
    if (on_device) *on_device = p->bs->elems_on_device() ? 1 : 0;
    return 0;
}

extern "C" int spsph_pipeline_compare(spsph_pipeline *p, uint32_t query_size, uint32_t *inter, uint64_t *sizes,
                                      int *full_rows, float *kernel_ms, uint64_t *launches)
{
    try {
        if (!p) return hfail("null pipeline");
        const uint64_t l0 = p->session->launches();
        std::vector<uint32_t> iv;
        std::vector<uint64_t> sv;
        bool full = false;
        p->bs->compare_last(query_size, iv, sv, full, kernel_ms);
        if (!iv.empty()) memcpy(inter, iv.data(), iv.size() * sizeof(uint32_t));
        if (!sv.empty()) memcpy(sizes, sv.data(), sv.size() * sizeof(uint64_t));
        if (full_rows) *full_rows = full ? 1 : 0;
        if (launches) *launches = p->session->launches() - l0;
        return 0;
    } catch (const std::exception &e) { return hfail(e.what()); }
}
