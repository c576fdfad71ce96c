// C ABI of the device layer (include/spsp.h): contexts, slots (streams),
// buffers and kernel launches.  No algorithmic work happens on the host here
// and there is no CPU fallback: every entry point needs a CUDA device.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/spsp.h"
#include "compare.cuh"
#include "dense.cuh"
#include "ingest.cuh"
#include "postpass.cuh"
#include "scan.cuh"

using namespace spsp;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(-1, std::string(#call) + ": " + cudaGetErrorString(e_));                      \
    } while (0)

// Grow-only device / pinned buffers: steady-state calls never hit cudaMalloc / cudaFree
// (both serialise against the whole device and against driver queries).
static void trace_alloc(const char *what, size_t from, size_t to)
{
    static const bool on = getenv("SPSP_TRACE_ALLOC") != nullptr;
    if (on) fprintf(stderr, "[alloc] %s buffer %zu -> %zu bytes\n", what, from, to);
}
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        trace_alloc("device", cap, bytes);
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        trace_alloc("pinned", cap, bytes);
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Slot {
    cudaStream_t stream = nullptr;
    uint32_t *d_packed = nullptr;
    uint64_t d_packed_words = 0;
    spsp_hit *d_hits = nullptr;
    uint64_t hits_cap = 0;
    unsigned long long *d_count = nullptr;
    unsigned long long *h_count = nullptr;     // pinned
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    uint64_t n_bases = 0;
    bool pending = false;
    bool timed = false;
    // whole-batch path
    // staged uploads alternate between two copy streams: one copy engine moves ~27 GB/s on this platform, two 55
    cudaStream_t up_stream[2] = {nullptr, nullptr};
    cudaEvent_t up_event[2] = {nullptr, nullptr};
    // raw FASTA text (device-side ingest) travels on two streams of its own: its long copies must not hold back the
    // short packed pieces queued behind them, and "lane idle" then means "the previous raw input has arrived"
    cudaStream_t raw_stream[2] = {nullptr, nullptr};
    cudaEvent_t raw_event[2] = {nullptr, nullptr};
    PostpassBuffers *pp = nullptr;
    DenseBuffers *dn = nullptr;
    DevBuf b_rec_begin, b_rec_end, b_rec_input;
    PostpassOut last_batch{};
    uint32_t last_batch_inputs = 0;
    bool has_batch = false;
    // device-side ingest (raw FASTA text -> regions of d_packed + record table)
    DevBuf b_text, b_ing_in, b_ing_tiles, b_ing_carry, b_ing_tot, b_ing_grand;
    DevBuf t_rec_begin, t_rec_end, t_rec_input;      // records of the text inputs of the staged batch
    DevBuf m_rec_begin, m_rec_end, m_rec_input;      // ... merged with the host-packed inputs' records
    PinBuf p_ing;
    uint64_t n_trec = 0;
    // record table of the last staged batch as the device saw it (host-packed, ingested or merged)
    const uint64_t *last_rb = nullptr, *last_re = nullptr;
    const uint32_t *last_ri = nullptr;
    uint64_t last_nrec = 0;
    DevBuf b_kmers;
    uint32_t text_rr = 0;
    cudaEvent_t ing0 = nullptr, ing1 = nullptr;
    bool ing_timed = false;
};

struct spsp_ctx {
    int device = 0, k = 0, m = 0;
    uint64_t thr = 0;
    int mode = SPSP_SCAN_AUTO;
    std::vector<Slot> slots;
    // q-gram filter
    bool filter_ready = false;
    bool filter_tried = false;
    bool filter_profitable = false;
    FilterParams fp{};
    uint32_t *d_table = nullptr, *d_exact = nullptr;
    uint64_t n_selected = 0;
    // compare
    uint32_t n_sketches = 0, n_chunks = 0;
    uint64_t n_elems = 0;
    bool has_hi = false, owns_cmp = false;
    CmpData cmp{};
    DevBuf b_minim, b_klo, b_khi, b_sk_off, b_chunk_off, b_tiles, b_out;
    PinBuf p_stage, p_out, p_units, p_skoff;
    cudaEvent_t cev0 = nullptr, cev1 = nullptr, ev_stage = nullptr;   // ev_stage: last copy out of the pinned staging buffers
    bool cmp_timed = false;
    // multi-GPU exchange (one process per GPU)
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    DevBuf x_hdrsz, x_klo, x_khi, x_min, x_begin, x_end, x_compact, x_units, x_dims, x_tiles, x_recv, x_neg;
    PinBuf xp_hdrsz, xp_out, xp_neg;
    uint64_t x_ncap = 0, x_qcap = 0, x_ecap = 0;      // negotiated slot capacities (identical on every rank)
    uint64_t launches = 0;
    std::mutex mu;
};

static int ensure_packed(struct Slot &s, uint64_t words);
static bool is_device_ptr(const void *p)
{
    cudaPointerAttributes a{};
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice;
}
static int check_records(const uint64_t *rec_begin, const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec,
                         uint64_t n_bases, uint32_t n_inputs, const char *who);

extern "C" int spsp_abi_version(void) { return SPSP_ABI_VERSION; }
extern "C" const char *spsp_last_error(void) { return g_err.c_str(); }

extern "C" int spsp_device_count(int *n)
{
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *n = 0; return fail(-1, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    *n = c;
    return 0;
}

extern "C" int spsp_warmup(int device)
{
    CK(cudaSetDevice(device));
    CK(cudaFree(nullptr));                       // creates the primary context
    return 0;
}

extern "C" uint64_t spsp_packed_words(uint64_t n_bases) { return 4 * ((n_bases + 4 + 63) / 64) + 8; }

extern "C" int spsp_host_alloc(void **p, size_t bytes)
{
    CK(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}
extern "C" int spsp_host_free(void *p)
{
    if (p) CK(cudaFreeHost(p));
    return 0;
}

// Cost model of the filter kernel, in instruction-equivalents per 64-base lane
// chunk, fitted to B200 measurements (profiles/r01_scan_sweep.md): the kernel is
// bound by shared-memory wavefronts, so a probe costs about the same whatever
// its instruction count; positives cost a queue entry plus the exact checks.
//   n_sel  expected number of selected forward m-mers (4^m * T / 2^64)
static double filter_cost(int m, int kind, int g, double n_sel, FilterParams *fp)
{
    int q = m - g + 1;
    if (q < 3) return 1e18;
    int bits = 2 * q, hashed = 0;
    if (kind == 1) {
        if (g != 4 || bits > 16) return 1e18;
    } else if (bits > FILTER_MAX_BITS) {
        bits = FILTER_MAX_BITS; hashed = 1;
    }
    const double n_probe = 64.0 / g;
    const double load = g * n_sel / std::ldexp(1.0, bits);
    const double delta = hashed ? 1.0 - std::exp(-load) : std::min(1.0, load);   // P(probe positive)
    const double lam = delta * n_probe;                                          // positives per lane chunk
    const double p_lane = 1.0 - std::pow(1.0 - delta, n_probe);                  // P(lane queues an entry)
    const double phases = kind == 1 ? 1.0 + 0.75 * delta : (double)g;            // exact checks per positive
    fp->g = g; fp->q = q; fp->bits = kind == 1 ? 2 * q + 3 : bits; fp->hashed = hashed; fp->rep_log2 = 0; fp->kind = kind;
    return n_probe * 5.5 + 30.0 + 20.0 * p_lane + lam * phases * 45.0;
}

// kind 2 (scan_rowbit_kernel): m == 11, few selected m-mers.  One conflict-free lookup per probe; a flagged
// probe costs a parked entry and four exact checks in shared memory.
static double rowbit_cost(int m, double n_sel, FilterParams *fp)
{
    if (m != 11) return 1e18;
    if (n_sel > 4096.0) return 1e18;
    int h = ROWBIT_MIN_HBITS;                                      // level-2 bits: <= 1/256 full where that fits
    while (h < ROWBIT_MAX_HBITS && std::ldexp(1.0, h) < 256.0 * n_sel) h++;
    const double delta = 1.0 - std::exp(-4.0 * n_sel / 32768.0);   // P(probe flagged): 4 phases, 15-bit key
    const double p_lane = 1.0 - std::pow(1.0 - delta, 16.0);
    fp->g = 4; fp->q = 8; fp->bits = 15; fp->hashed = 0; fp->rep_log2 = 5; fp->kind = 2; fp->hbits = h;
    // fitted to 320 Mbp launches at m = 11: 51.5 vs 58 us (byte table) at s = 1000, 59.9 vs 63.7 at 700, 67.0 vs 70.8 at 550,
    // 85 vs 82 at 400: the two cross near 450 selected m-mers
    return 70.0 + 25.0 * p_lane + 16.0 * delta * 72.0;
}

static int build_filter(spsp_ctx *c)
{
    const double p = (double)c->thr / 18446744073709551616.0;
    const double n_sel = std::ldexp(1.0, 2 * c->m) * p;
    const char *ek = getenv("SPSP_FILTER_KIND"), *eg = getenv("SPSP_FILTER_G");   // tuning overrides (experiments)
    for (int allow2 = 1; allow2 >= 0; allow2--) {
        double best = 1e18;
        FilterParams bfp{};
        for (int kind = 0; kind < 3; kind++)
            for (int g : {4, 2, 1}) {
                if (ek && atoi(ek) != kind && !(atoi(ek) == 2 && !allow2)) continue;
                if (eg && atoi(eg) != g && kind == 0) continue;
                if (kind == 2 && (g != 4 || !allow2)) continue;
                FilterParams fp{};
                double cost = kind == 2 ? rowbit_cost(c->m, n_sel, &fp) : filter_cost(c->m, kind, g, n_sel, &fp);
                if (cost < best) { best = cost; bfp = fp; }
            }
        const double dense_cost = 64.0 * 30.0;                 // full hash at every position
        c->filter_profitable = best < dense_cost;
        if (best >= 1e18) { c->filter_ready = false; return 0; }
        c->fp = bfp;
        CK(cudaMalloc(&c->d_table, filter_table_bytes(bfp)));
        if (!c->d_exact) CK(cudaMalloc(&c->d_exact, ((size_t)1 << (2 * c->m)) / 8));
        unsigned long long *d_n = nullptr;
        CK(cudaMalloc(&d_n, sizeof(unsigned long long)));
        CK(launch_filter_build(c->m, c->thr, bfp, c->d_table, c->d_exact, d_n, c->slots[0].stream));
        c->launches++;
        unsigned long long n = 0;
        CK(cudaMemcpyAsync(&n, d_n, sizeof n, cudaMemcpyDeviceToHost, c->slots[0].stream));
        CK(cudaStreamSynchronize(c->slots[0].stream));
        CK(cudaFree(d_n));
        c->n_selected = n;
        c->filter_ready = true;
        return 0;
    }
    return 0;
}

static int ensure_filter(spsp_ctx *c)
{
    std::lock_guard<std::mutex> g(c->mu);
    if (c->filter_tried) return 0;
    c->filter_tried = true;
    return build_filter(c);
}

extern "C" int spsp_create(int device, int k, int m, uint64_t threshold, int n_slots, spsp_ctx **out)
{
    if (!out) return fail(-3, "spsp_create: null out");
    *out = nullptr;
    if (k < 3 || k > 63 || m < 3 || m > 15 || m >= k || !(k & 1) || !(m & 1))
        return fail(-3, "spsp_create: need odd 3<=m<=15, odd m<k<=63");
    if (n_slots < 1 || n_slots > 64) return fail(-3, "spsp_create: n_slots out of range");
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0)
        return fail(-1, "spsp_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= cnt) return fail(-3, "spsp_create: bad device index");
    CK(cudaSetDevice(device));
    // How host threads wait for the device (SPSP_SCHED=spin|yield|block): spinning is fastest when every waiting
    // thread has a core of its own; with several ranks per box and few cores per rank the waiters must yield.
    if (const char *sch = getenv("SPSP_SCHED")) {
        unsigned fl = !strcmp(sch, "yield") ? cudaDeviceScheduleYield : !strcmp(sch, "block") ? cudaDeviceScheduleBlockingSync
                    : !strcmp(sch, "spin") ? cudaDeviceScheduleSpin : cudaDeviceScheduleAuto;
        cudaSetDeviceFlags(fl);
        cudaGetLastError();
    }
    spsp_ctx *c = new spsp_ctx();
    c->device = device; c->k = k; c->m = m; c->thr = threshold;
    c->slots.resize(n_slots);
    for (auto &s : c->slots) {
        CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CK(cudaMalloc(&s.d_count, sizeof(unsigned long long)));
        CK(cudaHostAlloc(&s.h_count, sizeof(unsigned long long), cudaHostAllocDefault));
        CK(cudaEventCreate(&s.ev0));
        CK(cudaEventCreate(&s.ev1));
        CK(cudaEventCreate(&s.ev2));
        CK(cudaEventCreate(&s.ing0));
        CK(cudaEventCreate(&s.ing1));
        for (int u = 0; u < 2; u++) {
            CK(cudaStreamCreateWithFlags(&s.up_stream[u], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&s.up_event[u], cudaEventDisableTiming));
            CK(cudaStreamCreateWithFlags(&s.raw_stream[u], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&s.raw_event[u], cudaEventDisableTiming));
        }
    }
    CK(cudaEventCreate(&c->cev0));
    CK(cudaEventCreate(&c->cev1));
    CK(cudaEventCreateWithFlags(&c->ev_stage, cudaEventDisableTiming));
    *out = c;                        // the q-gram filter table is built on first scan
    return 0;
}

static void free_cmp(spsp_ctx *c)
{
    c->b_minim.release(); c->b_klo.release(); c->b_khi.release(); c->b_sk_off.release();
    c->b_chunk_off.release(); c->b_tiles.release(); c->b_out.release();
    c->p_stage.release(); c->p_out.release(); c->p_units.release(); c->p_skoff.release();
    c->owns_cmp = false;
    c->n_sketches = 0;
}

static void nccl_release(spsp_ctx *c);

extern "C" int spsp_destroy(spsp_ctx *c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    for (auto &s : c->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        cudaFree(s.d_packed); cudaFree(s.d_hits); cudaFree(s.d_count);
        cudaFreeHost(s.h_count);
        if (s.ev0) cudaEventDestroy(s.ev0);
        if (s.ev1) cudaEventDestroy(s.ev1);
        if (s.ev2) cudaEventDestroy(s.ev2);
        if (s.ing0) cudaEventDestroy(s.ing0);
        if (s.ing1) cudaEventDestroy(s.ing1);
        for (int u = 0; u < 2; u++) {
            if (s.up_stream[u]) { cudaStreamSynchronize(s.up_stream[u]); cudaStreamDestroy(s.up_stream[u]); }
            if (s.up_event[u]) cudaEventDestroy(s.up_event[u]);
            if (s.raw_stream[u]) { cudaStreamSynchronize(s.raw_stream[u]); cudaStreamDestroy(s.raw_stream[u]); }
            if (s.raw_event[u]) cudaEventDestroy(s.raw_event[u]);
        }
        if (s.pp) postpass_buffers_destroy(s.pp);
        if (s.dn) dense_buffers_destroy(s.dn);
        s.b_rec_begin.release(); s.b_rec_end.release(); s.b_rec_input.release();
        s.b_text.release(); s.b_ing_in.release(); s.b_ing_tiles.release(); s.b_ing_carry.release(); s.b_ing_tot.release();
        s.b_ing_grand.release(); s.t_rec_begin.release(); s.t_rec_end.release(); s.t_rec_input.release();
        s.m_rec_begin.release(); s.m_rec_end.release(); s.m_rec_input.release(); s.p_ing.release(); s.b_kmers.release();
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    free_cmp(c);
    nccl_release(c);
    cudaFree(c->d_table);
    cudaFree(c->d_exact);
    if (c->cev0) cudaEventDestroy(c->cev0);
    if (c->cev1) cudaEventDestroy(c->cev1);
    if (c->ev_stage) cudaEventDestroy(c->ev_stage);
    delete c;
    return 0;
}

extern "C" int spsp_scan_config(spsp_ctx *c, int mode)
{
    if (!c) return fail(-3, "null ctx");
    if (mode < SPSP_SCAN_AUTO || mode > SPSP_SCAN_FILTER) return fail(-3, "spsp_scan_config: bad mode");
    if (mode == SPSP_SCAN_FILTER) {
        CK(cudaSetDevice(c->device));
        int rc = ensure_filter(c);
        if (rc) return rc;
        if (!c->filter_ready) return fail(-3, "spsp_scan_config: filter unavailable for this m");
    }
    c->mode = mode;
    return 0;
}

extern "C" int spsp_scan_filter_info(spsp_ctx *c, int *kind, int *g, uint64_t *n_selected)
{
    if (!c) return fail(-3, "null ctx");
    CK(cudaSetDevice(c->device));
    int rc = ensure_filter(c);
    if (rc) return rc;
    if (kind) *kind = c->filter_ready ? c->fp.kind : -1;
    if (g) *g = c->filter_ready ? c->fp.g : 0;
    if (n_selected) *n_selected = c->n_selected;
    return 0;
}

static int launch_scan(spsp_ctx *c, Slot &s, const uint32_t *d_packed, uint64_t n_bases, ScanOut out)
{
    if (c->mode != SPSP_SCAN_DENSE) {
        int rc = ensure_filter(c);
        if (rc) return rc;
    }
    bool use_filter = c->mode == SPSP_SCAN_FILTER || (c->mode == SPSP_SCAN_AUTO && c->filter_ready && c->filter_profitable);
    CK(cudaMemsetAsync(out.count, 0, sizeof(unsigned long long), s.stream));
    CK(cudaEventRecord(s.ev0, s.stream));
    if (use_filter)
        CK(launch_scan_filter(d_packed, n_bases, c->m, c->thr, c->fp, c->d_table, c->d_exact, out, s.stream));
    else
        CK(launch_scan_dense(d_packed, n_bases, c->m, c->thr, out, s.stream));
    CK(cudaEventRecord(s.ev1, s.stream));
    s.timed = true;
    if (n_bases >= (uint64_t)c->m) {
        std::lock_guard<std::mutex> g(c->mu);
        c->launches++;
    }
    return 0;
}

static int ensure_hits(Slot &s, uint64_t want)
{
    if (want <= s.hits_cap) return 0;
    if (s.d_hits) CK(cudaFree(s.d_hits));
    s.d_hits = nullptr; s.hits_cap = 0;
    CK(cudaMalloc(&s.d_hits, want * sizeof(spsp_hit)));
    s.hits_cap = want;
    return 0;
}

extern "C" int spsp_scan_submit(spsp_ctx *c, int slot, const uint32_t *packed, uint64_t n_bases)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_scan_submit: bad ctx/slot");
    if (!packed && n_bases) return fail(-3, "spsp_scan_submit: null input");
    Slot &s = c->slots[slot];
    if (s.pending) return fail(-3, "spsp_scan_submit: slot busy (collect first)");
    CK(cudaSetDevice(c->device));
    uint64_t words = spsp_packed_words(n_bases);
    { int rc_ = ensure_packed(s, words); if (rc_) return rc_; }
    CK(cudaMemcpyAsync(s.d_packed, packed, words * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
    double p = (double)c->thr / 18446744073709551616.0;
    uint64_t guess = (uint64_t)((double)n_bases * p * 1.5) + 4096;
    if (guess > n_bases + 1) guess = n_bases + 1;
    int rc = ensure_hits(s, guess);
    if (rc) return rc;
    s.n_bases = n_bases;
    ScanOut out{s.d_hits, s.d_count, s.hits_cap};
    rc = launch_scan(c, s, s.d_packed, n_bases, out);
    if (rc) return rc;
    CK(cudaMemcpyAsync(s.h_count, s.d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
    s.pending = true;
    return 0;
}

extern "C" int spsp_scan_collect(spsp_ctx *c, int slot, spsp_hit *hits_out, uint64_t cap, uint64_t *n_hits)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_scan_collect: bad ctx/slot");
    Slot &s = c->slots[slot];
    if (!s.pending) return fail(-3, "spsp_scan_collect: nothing submitted");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(s.stream));
    uint64_t n = *s.h_count;
    if (n > s.hits_cap) {
        // the estimate was too small: grow and rescan (exact, just slower)
        int rc = ensure_hits(s, n + n / 8 + 1024);
        if (rc) return rc;
        ScanOut out{s.d_hits, s.d_count, s.hits_cap};
        rc = launch_scan(c, s, s.d_packed, s.n_bases, out);
        if (rc) return rc;
        CK(cudaMemcpyAsync(s.h_count, s.d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        n = *s.h_count;
    }
    if (n_hits) *n_hits = n;
    if (n > cap) return fail(-2, "spsp_scan_collect: output buffer too small");
    if (n) {
        CK(cudaMemcpyAsync(hits_out, s.d_hits, n * sizeof(spsp_hit), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
    }
    s.pending = false;
    return 0;
}

extern "C" int spsp_scan_device(spsp_ctx *c, int slot, const uint32_t *d_packed, uint64_t n_bases, spsp_hit *d_hits,
                                uint64_t cap, uint64_t *d_count)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_scan_device: bad ctx/slot");
    Slot &s = c->slots[slot];
    CK(cudaSetDevice(c->device));
    ScanOut out{d_hits, reinterpret_cast<unsigned long long *>(d_count), cap};
    return launch_scan(c, s, d_packed, n_bases, out);
}

extern "C" int spsp_sync(spsp_ctx *c, int slot)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_sync: bad ctx/slot");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->slots[slot].stream));
    return 0;
}

extern "C" int spsp_scan_last_kernel_ms(spsp_ctx *c, int slot, float *ms)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size() || !ms) return fail(-3, "spsp_scan_last_kernel_ms: bad args");
    Slot &s = c->slots[slot];
    if (!s.timed) return fail(-3, "no scan launched on this slot");
    CK(cudaEventSynchronize(s.ev1));
    CK(cudaEventElapsedTime(ms, s.ev0, s.ev1));
    return 0;
}

extern "C" int spsp_stream(spsp_ctx *c, int slot, void **stream)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size() || !stream) return fail(-3, "spsp_stream: bad args");
    *stream = (void *)c->slots[slot].stream;
    return 0;
}

// ----------------------------------------------------------------- compare

// Everything below is enqueued on slot 0's stream; nothing synchronises until a result is needed on the host.
static int finish_cmp_load(spsp_ctx *c, uint32_t n_sketches, const uint64_t *sketch_off)
{
    cudaStream_t st = c->slots[0].stream;
    c->n_sketches = n_sketches;
    c->n_elems = sketch_off[n_sketches];
    for (uint32_t i = 0; i < n_sketches; i++)
        if (sketch_off[i] > sketch_off[i + 1]) return fail(-3, "spsp_cmp_load: sketch_off not monotone");
    CK(c->b_sk_off.ensure((size_t)(n_sketches + 1) * sizeof(uint64_t)));
    // the caller's offsets may be gone before the copy runs: stage them in pinned memory (the event orders this
    // against the previous copy out of the staging buffer; in the steady state it has long completed)
    CK(cudaEventSynchronize(c->ev_stage));
    CK(c->p_skoff.ensure((size_t)(n_sketches + 1) * sizeof(uint64_t)));
    memcpy(c->p_skoff.p, sketch_off, (size_t)(n_sketches + 1) * sizeof(uint64_t));
    CK(cudaMemcpyAsync(c->b_sk_off.p, c->p_skoff.p, (size_t)(n_sketches + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(c->ev_stage, st));
    double avg = n_sketches ? (double)c->n_elems / n_sketches : 0.0;
    uint64_t chunks = (uint64_t)std::ceil(32.0 * avg / (0.7 * CMP_CAP));
    if (chunks < 1) chunks = 1;
    if (chunks > 8192) chunks = 8192;
    c->n_chunks = (uint32_t)chunks;
    CK(c->b_chunk_off.ensure((size_t)n_sketches * (chunks + 1) * sizeof(uint64_t)));
    c->cmp.sk_off = static_cast<const uint64_t *>(c->b_sk_off.p);
    c->cmp.sk_end = nullptr;
    c->cmp.chunk_off = static_cast<uint64_t *>(c->b_chunk_off.p);
    CK(launch_chunk_offsets(c->cmp, n_sketches, nullptr, c->n_chunks, c->m, st));
    c->launches++;
    return 0;
}

extern "C" int spsp_cmp_load(spsp_ctx *c, uint32_t n_sketches, const uint64_t *sketch_off, const uint32_t *minimizer,
                             const uint64_t *kmer_lo, const uint64_t *kmer_hi)
{
    if (!c || !sketch_off) return fail(-3, "spsp_cmp_load: bad args");
    if ((c->k > 32) != (kmer_hi != nullptr)) return fail(-3, "spsp_cmp_load: kmer_hi must be given iff k > 32");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->slots[0].stream;
    CK(cudaStreamSynchronize(st));            // previous run may still read the buffers
    uint64_t E = sketch_off[n_sketches];
    size_t e1 = E ? E : 1;
    CK(c->b_minim.ensure(e1 * sizeof(uint32_t)));
    CK(c->b_klo.ensure(e1 * sizeof(uint64_t)));
    if (kmer_hi) CK(c->b_khi.ensure(e1 * sizeof(uint64_t)));
    c->owns_cmp = true;
    if (E) {
        // stage through pinned memory so the three copies are truly asynchronous
        size_t bytes = E * (sizeof(uint32_t) + sizeof(uint64_t) * (kmer_hi ? 2 : 1));
        CK(c->p_stage.ensure(bytes));
        unsigned char *sp = static_cast<unsigned char *>(c->p_stage.p);
        memcpy(sp, kmer_lo, E * 8);
        CK(cudaMemcpyAsync(c->b_klo.p, sp, E * 8, cudaMemcpyHostToDevice, st));
        sp += E * 8;
        if (kmer_hi) {
            memcpy(sp, kmer_hi, E * 8);
            CK(cudaMemcpyAsync(c->b_khi.p, sp, E * 8, cudaMemcpyHostToDevice, st));
            sp += E * 8;
        }
        memcpy(sp, minimizer, E * 4);
        CK(cudaMemcpyAsync(c->b_minim.p, sp, E * 4, cudaMemcpyHostToDevice, st));
    }
    c->has_hi = kmer_hi != nullptr;
    c->cmp.minim = static_cast<const uint32_t *>(c->b_minim.p);
    c->cmp.klo = static_cast<const uint64_t *>(c->b_klo.p);
    c->cmp.khi = kmer_hi ? static_cast<const uint64_t *>(c->b_khi.p) : nullptr;
    return finish_cmp_load(c, n_sketches, sketch_off);
}

extern "C" int spsp_cmp_load_device(spsp_ctx *c, uint32_t n_sketches, const uint64_t *sketch_off_host,
                                    const uint32_t *d_minimizer, const uint64_t *d_kmer_lo, const uint64_t *d_kmer_hi)
{
    if (!c || !sketch_off_host) return fail(-3, "spsp_cmp_load_device: bad args");
    if ((c->k > 32) != (d_kmer_hi != nullptr)) return fail(-3, "spsp_cmp_load_device: kmer_hi must be given iff k > 32");
    CK(cudaSetDevice(c->device));
    c->owns_cmp = false;
    c->has_hi = d_kmer_hi != nullptr;
    c->cmp.minim = d_minimizer; c->cmp.klo = d_kmer_lo; c->cmp.khi = d_kmer_hi;
    return finish_cmp_load(c, n_sketches, sketch_off_host);
}

// Units of the tile grid in dealing order (the device twin is exchange_plan_kernel): column tile jb, row tiles in
// runs of rt; symmetric: only row tiles 0 .. jb.  f(index, jb, ib0, n_ib).
template <class F>
static void for_each_unit(uint32_t nI, uint32_t nJ, bool symmetric, uint32_t rt, F f)
{
    uint64_t idx = 0;
    for (uint32_t jb = 0; jb < nJ; jb++) {
        const uint32_t lim = symmetric ? std::min(jb + 1, nI) : nI;
        for (uint32_t ib0 = 0; ib0 < lim; ib0 += rt, idx++) f(idx, jb, ib0, std::min(rt, lim - ib0));
    }
}

static int cmp_run_impl(spsp_ctx *c, uint32_t row_begin, uint32_t row_end, uint32_t col_begin, uint32_t col_end,
                        int symmetric, uint32_t tile_rank, uint32_t tile_ranks, uint32_t *d_out, uint64_t ld)
{
    if (row_end > c->n_sketches || col_end > c->n_sketches || row_begin > row_end || col_begin > col_end)
        return fail(-3, "spsp_cmp_run: range outside the loaded sketches");
    if (ld < (uint64_t)(col_end - col_begin)) return fail(-3, "spsp_cmp_run: ld too small");
    if (tile_ranks == 0 || tile_rank >= tile_ranks) return fail(-3, "spsp_cmp_run: bad tile rank");
    if (symmetric && (row_begin != col_begin || row_end != col_end))
        return fail(-3, "spsp_cmp_run: symmetric needs identical row and column ranges");
    cudaStream_t st = c->slots[0].stream;
    const uint32_t nI = (row_end - row_begin + 31) / 32, nJ = (col_end - col_begin + 31) / 32;
    if (nI > 0xFFFFu || nJ > 0xFFFFu) return fail(-3, "spsp_cmp_run: more than 2^21 sketches per side");
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    // row tiles per unit: as many as share one table build (CMP_RT) when the grid is large; fewer when that would
    // leave SMs without a CTA (small jobs are bound by the longest CTA, not by the table builds)
    uint32_t rt = c->has_hi ? CMP_RT_HI : CMP_RT;
    std::vector<uint2> units;
    for (;; rt /= 2) {
        units.clear();
        for_each_unit(nI, nJ, symmetric != 0, rt, [&](uint64_t idx, uint32_t jb, uint32_t ib0, uint32_t n_ib) {
            if (idx % tile_ranks == tile_rank) units.push_back(make_uint2(jb | (ib0 << 16), n_ib));
        });
        if (rt == 1 || units.size() * (uint64_t)c->n_chunks >= 2ull * sms) break;
    }
    CK(cudaEventRecord(c->cev0, st));
    if (!units.empty()) {
        // pinned staging that outlives the call: no stream synchronisation for the copy
        CK(cudaEventSynchronize(c->ev_stage));
        CK(c->p_units.ensure(units.size() * sizeof(uint2)));
        memcpy(c->p_units.p, units.data(), units.size() * sizeof(uint2));
        CK(c->b_tiles.ensure(units.size() * sizeof(uint2)));
        uint2 *d_units = static_cast<uint2 *>(c->b_tiles.p);
        CK(cudaMemcpyAsync(d_units, c->p_units.p, units.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(c->ev_stage, st));
        CK(cudaEventRecord(c->cev0, st));
        uint64_t groups = (2ull * sms + units.size() - 1) / units.size();
        if (groups < 1) groups = 1;
        if (groups > c->n_chunks) groups = c->n_chunks;
        CK(launch_hashjoin(c->cmp, c->has_hi, d_units, (uint32_t)units.size(), nullptr, c->n_chunks, (uint32_t)groups, row_begin,
                           row_end, col_begin, col_end, d_out, ld, st));
        c->launches++;
    }
    CK(cudaEventRecord(c->cev1, st));
    c->cmp_timed = true;
    return 0;
}

extern "C" int spsp_cmp_run_device(spsp_ctx *c, uint32_t row_begin, uint32_t row_end, uint32_t col_begin,
                                   uint32_t col_end, int symmetric, uint32_t tile_rank, uint32_t tile_ranks,
                                   uint32_t *d_out, uint64_t ld)
{
    if (!c || !d_out) return fail(-3, "spsp_cmp_run_device: bad args");
    CK(cudaSetDevice(c->device));
    int rc = cmp_run_impl(c, row_begin, row_end, col_begin, col_end, symmetric, tile_rank, tile_ranks, d_out, ld);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->slots[0].stream));       // documented: the call returns when the counts are in d_out
    return 0;
}

extern "C" int spsp_cmp_run(spsp_ctx *c, uint32_t row_begin, uint32_t row_end, uint32_t col_begin, uint32_t col_end,
                            int symmetric, uint32_t tile_rank, uint32_t tile_ranks, uint32_t *out, uint64_t ld)
{
    if (!c || !out) return fail(-3, "spsp_cmp_run: bad args");
    CK(cudaSetDevice(c->device));
    uint64_t rows = row_end - row_begin, cols = col_end - col_begin;
    if (!rows || !cols) return 0;
    cudaStream_t st = c->slots[0].stream;
    CK(c->b_out.ensure(rows * cols * sizeof(uint32_t)));
    CK(c->p_out.ensure(rows * cols * sizeof(uint32_t)));
    uint32_t *d_out = static_cast<uint32_t *>(c->b_out.p);
    CK(cudaMemsetAsync(d_out, 0, rows * cols * sizeof(uint32_t), st));
    int rc = cmp_run_impl(c, row_begin, row_end, col_begin, col_end, symmetric, tile_rank, tile_ranks, d_out, cols);
    if (rc) return rc;
    const uint32_t *tmp = static_cast<const uint32_t *>(c->p_out.p);
    CK(cudaMemcpyAsync(c->p_out.p, d_out, rows * cols * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));                       // the only synchronisation of load + run
    for (uint64_t r = 0; r < rows; r++)
        for (uint64_t q = 0; q < cols; q++) out[r * ld + q] += tmp[r * cols + q];
    return 0;
}

extern "C" int spsp_cmp_last_kernel_ms(spsp_ctx *c, float *ms)
{
    if (!c || !ms) return fail(-3, "spsp_cmp_last_kernel_ms: bad args");
    if (!c->cmp_timed) return fail(-3, "no compare launched");
    CK(cudaEventSynchronize(c->cev1));
    CK(cudaEventElapsedTime(ms, c->cev0, c->cev1));
    return 0;
}

// ------------------------------------------------------------ whole batch

static int batch_impl(spsp_ctx *c, Slot &s, const uint32_t *d_packed, uint64_t n_bases, const uint64_t *rec_begin,
                      const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec, uint32_t n_inputs,
                      unsigned abundance, spsp_batch_result *res)
{
    if (!res) return fail(-3, "spsp_sketch_batch: null result");
    cudaStream_t st = s.stream;
    // Record tables may already live on the device (a caller that keeps its inputs resident: millions of short
    // records cost more to re-validate and re-upload than to scan); they are then used in place, unchecked.
    const bool rec_on_device = n_rec && is_device_ptr(rec_begin);
    const uint64_t *d_rec_begin = rec_begin, *d_rec_end = rec_end;
    const uint32_t *d_rec_input = rec_input;
    if (rec_on_device) {
        if (!is_device_ptr(rec_end) || !is_device_ptr(rec_input))
            return fail(-3, "spsp_sketch_batch: record arrays must all be host or all be device memory");
    } else {
        { int rc_ = check_records(rec_begin, rec_end, rec_input, n_rec, n_bases, n_inputs, "spsp_sketch_batch"); if (rc_) return rc_; }
        const size_t nr = n_rec ? n_rec : 1;
        CK(s.b_rec_begin.ensure(nr * 8)); CK(s.b_rec_end.ensure(nr * 8)); CK(s.b_rec_input.ensure(nr * 4));
        if (n_rec) {
            CK(cudaMemcpyAsync(s.b_rec_begin.p, rec_begin, n_rec * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(s.b_rec_end.p, rec_end, n_rec * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(s.b_rec_input.p, rec_input, n_rec * 4, cudaMemcpyHostToDevice, st));
        }
        d_rec_begin = static_cast<const uint64_t *>(s.b_rec_begin.p);
        d_rec_end = static_cast<const uint64_t *>(s.b_rec_end.p);
        d_rec_input = static_cast<const uint32_t *>(s.b_rec_input.p);
    }
    // scan + post-pass are enqueued back to back: the hit count stays on the device, the host synchronises once,
    // when the sketch bytes have arrived.  A capacity that turns out too small (hits, pieces, bytes) makes the
    // pass report a retry; the buffers remember the larger size.
    const double p = (double)c->thr / 18446744073709551616.0;
    uint64_t guess = (uint64_t)((double)n_bases * p * 1.5) + 4096;
    if (guess > n_bases + 1) guess = n_bases + 1;
    int rc = ensure_hits(s, guess);
    if (rc) return rc;
    if (!s.pp) s.pp = postpass_buffers_create();
    uint64_t n_hits = 0;
    float scan_ms = 0, post_ms = 0;
    uint32_t launched = 0;
    bool rescan = true;
    for (int attempt = 0;; attempt++) {
        if (attempt == 6) return fail(-1, "spsp_sketch_batch: capacities did not converge");
        if (rescan) {
            ScanOut so{s.d_hits, s.d_count, s.hits_cap};
            rc = launch_scan(c, s, d_packed, n_bases, so);
            if (rc) return rc;
        } else {
            CK(cudaEventRecord(s.ev1, st));
        }
        PostpassIn in{};
        in.d_packed = d_packed; in.n_bases = n_bases; in.d_hits = s.d_hits; in.d_hit_count = s.d_count; in.hits_cap = s.hits_cap;
        in.d_rec_begin = d_rec_begin; in.d_rec_end = d_rec_end; in.d_rec_input = d_rec_input;
        in.n_rec = n_rec; in.n_inputs = n_inputs; in.k = c->k; in.m = c->m; in.abundance = abundance;
        cudaError_t e = postpass_run(s.pp, in, &s.last_batch, st);
        if (e != cudaSuccess) {
            s.has_batch = false;
            if (e == cudaErrorInvalidValue)
                return fail(-4, "device post-pass: batch too large (>= 2^32 bases, or >= 2^31 k-mer entries expected at this "
                                "sampling rate); cut the job into smaller batches");
            return fail(-1, std::string("device post-pass: ") + cudaGetErrorString(e));
        }
        launched += s.last_batch.kernels_launched;
        n_hits = s.last_batch.n_hits;
        if (s.last_batch.retry == PP_RETRY_HITS) {
            rc = ensure_hits(s, n_hits + n_hits / 8 + 1024);
            if (rc) return rc;
            rescan = true;
            continue;
        }
        if (s.last_batch.retry == PP_RETRY_POSTPASS) { rescan = false; continue; }
        break;
    }
    CK(cudaEventRecord(s.ev2, st));
    CK(cudaEventSynchronize(s.ev2));
    if (rescan) CK(cudaEventElapsedTime(&scan_ms, s.ev0, s.ev1));
    CK(cudaEventElapsedTime(&post_ms, s.ev1, s.ev2));
    {
        std::lock_guard<std::mutex> g(c->mu);
        c->launches += launched;
    }
    s.has_batch = true;
    s.last_batch_inputs = n_inputs;
    res->body = s.last_batch.h_body; res->body_off = s.last_batch.h_body_off; res->selected = s.last_batch.h_selected;
    res->elem_off = s.last_batch.h_elem_off; res->n_hits = n_hits; res->n_elems = s.last_batch.n_elems;
    res->scan_ms = scan_ms; res->post_ms = post_ms;
    return 0;
}

extern "C" int spsp_sketch_batch_device(spsp_ctx *c, int slot, const uint32_t *d_packed, uint64_t n_bases,
                                        const uint64_t *rec_begin, const uint64_t *rec_end, const uint32_t *rec_input,
                                        uint64_t n_rec, uint32_t n_inputs, unsigned abundance, spsp_batch_result *res)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_sketch_batch_device: bad ctx/slot");
    CK(cudaSetDevice(c->device));
    return batch_impl(c, c->slots[slot], d_packed, n_bases, rec_begin, rec_end, rec_input, n_rec, n_inputs, abundance, res);
}

extern "C" int spsp_sketch_batch(spsp_ctx *c, int slot, const uint32_t *packed, uint64_t n_bases, const uint64_t *rec_begin,
                                 const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec, uint32_t n_inputs,
                                 unsigned abundance, spsp_batch_result *res)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_sketch_batch: bad ctx/slot");
    if (!packed && n_bases) return fail(-3, "spsp_sketch_batch: null input");
    Slot &s = c->slots[slot];
    if (s.pending) return fail(-3, "spsp_sketch_batch: slot busy (collect first)");
    CK(cudaSetDevice(c->device));
    uint64_t words = spsp_packed_words(n_bases);
    { int rc_ = ensure_packed(s, words); if (rc_) return rc_; }
    CK(cudaMemcpyAsync(s.d_packed, packed, words * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
    return batch_impl(c, s, s.d_packed, n_bases, rec_begin, rec_end, rec_input, n_rec, n_inputs, abundance, res);
}

static int ensure_packed(Slot &s, uint64_t words)
{
    if (words <= s.d_packed_words) return 0;
    if (s.d_packed) CK(cudaFree(s.d_packed));
    s.d_packed = nullptr; s.d_packed_words = 0;
    uint64_t cap = words + words / 4;
    CK(cudaMalloc(&s.d_packed, cap * sizeof(uint32_t)));
    s.d_packed_words = cap;
    return 0;
}

extern "C" int spsp_batch_reserve(spsp_ctx *c, int slot, uint64_t total_words)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_batch_reserve: bad ctx/slot");
    CK(cudaSetDevice(c->device));
    Slot &s = c->slots[slot];
    CK(cudaStreamSynchronize(s.stream));          // nothing may still read the old buffer
    for (int u = 0; u < 2; u++) CK(cudaStreamSynchronize(s.up_stream[u]));
    return ensure_packed(s, total_words);
}

extern "C" int spsp_batch_upload(spsp_ctx *c, int slot, uint64_t word_off, const uint32_t *host_words, uint64_t n_words)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_batch_upload: bad ctx/slot");
    Slot &s = c->slots[slot];
    if (word_off + n_words > s.d_packed_words) return fail(-3, "spsp_batch_upload: outside the reserved buffer");
    if (!n_words) return 0;
    if (!host_words) return fail(-3, "spsp_batch_upload: null input");
    CK(cudaSetDevice(c->device));
    // consecutive 256 KB pieces alternate between the two copy streams (no shared state: callable from any thread)
    cudaStream_t up = s.up_stream[(word_off >> 16) & 1];
    CK(cudaMemcpyAsync(s.d_packed + word_off, host_words, n_words * sizeof(uint32_t), cudaMemcpyHostToDevice, up));
    return 0;
}

extern "C" int spsp_sketch_batch_staged(spsp_ctx *c, int slot, uint64_t n_bases, const uint64_t *rec_begin,
                                        const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec,
                                        uint32_t n_inputs, unsigned abundance, spsp_batch_result *res)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_sketch_batch_staged: bad ctx/slot");
    Slot &s = c->slots[slot];
    if (spsp_packed_words(n_bases) > s.d_packed_words) return fail(-3, "spsp_sketch_batch_staged: reserve the buffer first");
    CK(cudaSetDevice(c->device));
    for (int u = 0; u < 2; u++) {                 // the scan waits for every upload queued so far
        CK(cudaEventRecord(s.up_event[u], s.up_stream[u]));
        CK(cudaStreamWaitEvent(s.stream, s.up_event[u], 0));
    }
    if (s.n_trec) {
        // part of the batch was ingested on the device (spsp_batch_text_pack): its records are there already
        const uint64_t nt = s.n_trec;
        s.n_trec = 0;
        const uint64_t *tb = static_cast<const uint64_t *>(s.t_rec_begin.p), *te = static_cast<const uint64_t *>(s.t_rec_end.p);
        const uint32_t *ti = static_cast<const uint32_t *>(s.t_rec_input.p);
        if (!n_rec) {
            s.last_rb = tb; s.last_re = te; s.last_ri = ti; s.last_nrec = nt;
            return batch_impl(c, s, s.d_packed, n_bases, tb, te, ti, nt, n_inputs, abundance, res);
        }
        if (is_device_ptr(rec_begin)) return fail(-3, "spsp_sketch_batch_staged: host record arrays expected next to ingested text");
        { int rc_ = check_records(rec_begin, rec_end, rec_input, n_rec, n_bases, n_inputs, "spsp_sketch_batch_staged"); if (rc_) return rc_; }
        CK(s.b_rec_begin.ensure(n_rec * 8)); CK(s.b_rec_end.ensure(n_rec * 8)); CK(s.b_rec_input.ensure(n_rec * 4));
        CK(cudaMemcpyAsync(s.b_rec_begin.p, rec_begin, n_rec * 8, cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.b_rec_end.p, rec_end, n_rec * 8, cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.b_rec_input.p, rec_input, n_rec * 4, cudaMemcpyHostToDevice, s.stream));
        const uint64_t n_all = n_rec + nt;
        CK(s.m_rec_begin.ensure(n_all * 8)); CK(s.m_rec_end.ensure(n_all * 8)); CK(s.m_rec_input.ensure(n_all * 4));
        CK(launch_record_merge(static_cast<const uint64_t *>(s.b_rec_begin.p), static_cast<const uint64_t *>(s.b_rec_end.p),
                               static_cast<const uint32_t *>(s.b_rec_input.p), n_rec, tb, te, ti, nt,
                               static_cast<uint64_t *>(s.m_rec_begin.p), static_cast<uint64_t *>(s.m_rec_end.p),
                               static_cast<uint32_t *>(s.m_rec_input.p), s.stream));
        { std::lock_guard<std::mutex> g(c->mu); c->launches += 1; }
        s.last_rb = static_cast<const uint64_t *>(s.m_rec_begin.p); s.last_re = static_cast<const uint64_t *>(s.m_rec_end.p);
        s.last_ri = static_cast<const uint32_t *>(s.m_rec_input.p); s.last_nrec = n_all;
        return batch_impl(c, s, s.d_packed, n_bases, static_cast<const uint64_t *>(s.m_rec_begin.p),
                          static_cast<const uint64_t *>(s.m_rec_end.p), static_cast<const uint32_t *>(s.m_rec_input.p), n_all,
                          n_inputs, abundance, res);
    }
    const int rc = batch_impl(c, s, s.d_packed, n_bases, rec_begin, rec_end, rec_input, n_rec, n_inputs, abundance, res);
    if (is_device_ptr(rec_begin)) { s.last_rb = rec_begin; s.last_re = rec_end; s.last_ri = rec_input; }
    else {
        s.last_rb = static_cast<const uint64_t *>(s.b_rec_begin.p); s.last_re = static_cast<const uint64_t *>(s.b_rec_end.p);
        s.last_ri = static_cast<const uint32_t *>(s.b_rec_input.p);
    }
    s.last_nrec = n_rec;
    return rc;
}

extern "C" int spsp_batch_record_kmers(spsp_ctx *c, int slot, uint32_t n_inputs, uint64_t *kmers_out)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size() || !kmers_out) return fail(-3, "spsp_batch_record_kmers: bad args");
    Slot &s = c->slots[slot];
    CK(cudaSetDevice(c->device));
    const size_t bytes = (size_t)(n_inputs ? n_inputs : 1) * 8;
    CK(s.b_kmers.ensure(bytes));
    CK(cudaMemsetAsync(s.b_kmers.p, 0, bytes, s.stream));
    CK(launch_record_kmers(s.last_rb, s.last_re, s.last_ri, s.last_nrec, (uint32_t)c->k, static_cast<unsigned long long *>(s.b_kmers.p),
                           s.stream));
    CK(cudaMemcpyAsync(kmers_out, s.b_kmers.p, (size_t)n_inputs * 8, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    { std::lock_guard<std::mutex> g(c->mu); c->launches += 1; }
    return 0;
}

// ------------------------------------------------------------ device-side ingest
//
// Raw FASTA text goes to the device as it is (pinned host memory -> async copies on the slot's two copy streams);
// three kernels (csrc/device/ingest.cu) do what getLineFasta + clean_dna + the 2-bit packer do on the host and
// leave the cleaned regions in the slot's staged batch buffer plus a record table.  A batch may mix inputs packed
// on the host (spsp_batch_upload) with inputs ingested here; spsp_sketch_batch_staged merges the record tables.

extern "C" int spsp_batch_text_reserve(spsp_ctx *c, int slot, uint64_t text_bytes)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_batch_text_reserve: bad ctx/slot");
    CK(cudaSetDevice(c->device));
    Slot &s = c->slots[slot];
    if (text_bytes + 64 <= s.b_text.cap) return 0;
    CK(cudaStreamSynchronize(s.stream));          // nothing may still read the old buffer
    for (int u = 0; u < 2; u++) CK(cudaStreamSynchronize(s.raw_stream[u]));
    CK(s.b_text.ensure(text_bytes + 64));
    return 0;
}

extern "C" int spsp_batch_text_upload(spsp_ctx *c, int slot, int lane, uint64_t byte_off, const uint8_t *host_text,
                                      uint64_t n_bytes)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size() || lane > 1) return fail(-3, "spsp_batch_text_upload: bad ctx/slot/lane");
    Slot &s = c->slots[slot];
    if (byte_off + n_bytes + 64 > s.b_text.cap) return fail(-3, "spsp_batch_text_upload: outside the reserved buffer");
    if (!n_bytes) return 0;
    if (!host_text) return fail(-3, "spsp_batch_text_upload: null input");
    CK(cudaSetDevice(c->device));
    // lane < 0: alternate between the two copy streams (callable from any thread)
    cudaStream_t up = s.raw_stream[lane >= 0 ? lane : (int)(__atomic_fetch_add(&s.text_rr, 1u, __ATOMIC_RELAXED) & 1u)];
    CK(cudaMemcpyAsync(static_cast<uint8_t *>(s.b_text.p) + byte_off, host_text, n_bytes, cudaMemcpyHostToDevice, up));
    return 0;
}

extern "C" int spsp_batch_upload_idle(spsp_ctx *c, int slot, int lane, int *idle)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size() || lane < 0 || lane > 1 || !idle)
        return fail(-3, "spsp_batch_upload_idle: bad ctx/slot/lane");
    CK(cudaSetDevice(c->device));
    const cudaError_t e = cudaStreamQuery(c->slots[slot].raw_stream[lane]);
    if (e != cudaSuccess && e != cudaErrorNotReady) return fail(-1, std::string("cudaStreamQuery: ") + cudaGetErrorString(e));
    *idle = e == cudaSuccess ? 1 : 0;
    return 0;
}

extern "C" int spsp_batch_upload_wait(spsp_ctx *c, int slot, int lane)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size() || lane > 1) return fail(-3, "spsp_batch_upload_wait: bad ctx/slot/lane");
    Slot &s = c->slots[slot];
    CK(cudaSetDevice(c->device));
    for (int u = 0; u < 2; u++)
        if (lane < 0 || lane == u) CK(cudaStreamSynchronize(s.raw_stream[u]));
    return 0;
}

extern "C" int spsp_batch_text_pack(spsp_ctx *c, int slot, uint32_t n_text, const uint64_t *text_off, const uint64_t *text_len,
                                    const uint64_t *word_off, const uint32_t *input_index, uint64_t *n_bases_out,
                                    uint64_t *n_rec_out)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_batch_text_pack: bad ctx/slot");
    Slot &s = c->slots[slot];
    s.n_trec = 0;
    if (!n_text) return 0;
    if (!text_off || !text_len || !word_off || !input_index) return fail(-3, "spsp_batch_text_pack: null arrays");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = s.stream;
    // ---- the inputs, validated: aligned text, ascending disjoint regions inside the staged buffer
    const size_t in_bytes = (size_t)n_text * sizeof(IngestInput), tot_bytes = (size_t)n_text * sizeof(IngestTotals);
    CK(s.p_ing.ensure(in_bytes + tot_bytes + 16));
    CK(cudaStreamSynchronize(st));                // p_ing / the ingest scratch of the previous call
    IngestInput *h_in = static_cast<IngestInput *>(s.p_ing.p);
    IngestTotals *h_tot = reinterpret_cast<IngestTotals *>(static_cast<uint8_t *>(s.p_ing.p) + in_bytes);
    uint64_t *h_grand = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(s.p_ing.p) + in_bytes + tot_bytes);
    uint64_t n_tiles = 0;
    for (uint32_t i = 0; i < n_text; i++) {
        if (text_off[i] & 15) return fail(-3, "spsp_batch_text_pack: text offsets must be multiples of 16");
        if (text_off[i] + text_len[i] + 64 > s.b_text.cap) return fail(-3, "spsp_batch_text_pack: text outside the reserved buffer");
        if (word_off[i] + spsp_packed_words(text_len[i]) > s.d_packed_words || (word_off[i] & 3))
            return fail(-3, "spsp_batch_text_pack: region outside the reserved batch buffer");
        if (i && (word_off[i] < word_off[i - 1] + spsp_packed_words(text_len[i - 1]) || input_index[i] <= input_index[i - 1]))
            return fail(-3, "spsp_batch_text_pack: inputs must be ascending with disjoint regions");
        h_in[i] = IngestInput{text_off[i], text_len[i], word_off[i], n_tiles, input_index[i], 0};
        n_tiles += (text_len[i] + ING_TILE - 1) / ING_TILE;
    }
    CK(s.b_ing_in.ensure(in_bytes)); CK(s.b_ing_tot.ensure(tot_bytes)); CK(s.b_ing_grand.ensure(16));
    CK(s.b_ing_tiles.ensure((n_tiles + 1) * sizeof(IngestTile))); CK(s.b_ing_carry.ensure((n_tiles + 1) * sizeof(IngestCarry)));
    for (int u = 0; u < 2; u++) {                 // the kernels wait for every text upload queued so far
        CK(cudaEventRecord(s.raw_event[u], s.raw_stream[u]));
        CK(cudaStreamWaitEvent(st, s.raw_event[u], 0));
    }
    const IngestInput *d_in = static_cast<const IngestInput *>(s.b_ing_in.p);
    IngestTotals *d_tot = static_cast<IngestTotals *>(s.b_ing_tot.p);
    const uint8_t *d_text = static_cast<const uint8_t *>(s.b_text.p);
    CK(cudaMemcpyAsync(s.b_ing_in.p, h_in, in_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(s.ing0, st));
    CK(launch_ingest_summary(d_text, d_in, n_text, n_tiles, static_cast<IngestTile *>(s.b_ing_tiles.p), st));
    CK(launch_ingest_carry(d_in, n_text, static_cast<const IngestTile *>(s.b_ing_tiles.p), static_cast<IngestCarry *>(s.b_ing_carry.p),
                           d_tot, static_cast<uint64_t *>(s.b_ing_grand.p), st));
    CK(cudaMemcpyAsync(h_tot, d_tot, tot_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_grand, s.b_ing_grand.p, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));                // the record table is sized from the count
    const uint64_t n_rec = h_grand[1];
    CK(s.t_rec_begin.ensure((n_rec + 1) * 8)); CK(s.t_rec_end.ensure((n_rec + 1) * 8)); CK(s.t_rec_input.ensure((n_rec + 1) * 4));
    CK(launch_ingest_write(d_text, d_in, n_text, n_tiles, static_cast<const IngestCarry *>(s.b_ing_carry.p), d_tot, s.d_packed,
                           static_cast<uint64_t *>(s.t_rec_begin.p), static_cast<uint64_t *>(s.t_rec_end.p),
                           static_cast<uint32_t *>(s.t_rec_input.p), st));
    CK(cudaEventRecord(s.ing1, st));
    s.ing_timed = true;
    s.n_trec = n_rec;
    for (uint32_t i = 0; i < n_text; i++) {
        if (n_bases_out) n_bases_out[i] = h_tot[i].n_bases;
        if (n_rec_out) n_rec_out[i] = h_tot[i].n_rec;
    }
    { std::lock_guard<std::mutex> g(c->mu); c->launches += 5; }
    return 0;
}

extern "C" int spsp_batch_text_last_ms(spsp_ctx *c, int slot, float *ms)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size() || !ms) return fail(-3, "spsp_batch_text_last_ms: bad args");
    Slot &s = c->slots[slot];
    *ms = 0;
    if (!s.ing_timed) return 0;
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(s.ing1));
    CK(cudaEventElapsedTime(ms, s.ing0, s.ing1));
    return 0;
}

/* Copy of the record table spsp_batch_text_pack left on the slot (tests, diagnostics): n entries each. */
extern "C" int spsp_batch_text_records(spsp_ctx *c, int slot, uint64_t *rec_begin, uint64_t *rec_end, uint32_t *rec_input,
                                       uint64_t cap, uint64_t *n_rec)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size() || !n_rec) return fail(-3, "spsp_batch_text_records: bad args");
    Slot &s = c->slots[slot];
    *n_rec = s.n_trec;
    if (s.n_trec > cap) return fail(-2, "spsp_batch_text_records: arrays too small");
    CK(cudaSetDevice(c->device));
    if (s.n_trec) {
        if (rec_begin) CK(cudaMemcpyAsync(rec_begin, s.t_rec_begin.p, s.n_trec * 8, cudaMemcpyDeviceToHost, s.stream));
        if (rec_end) CK(cudaMemcpyAsync(rec_end, s.t_rec_end.p, s.n_trec * 8, cudaMemcpyDeviceToHost, s.stream));
        if (rec_input) CK(cudaMemcpyAsync(rec_input, s.t_rec_input.p, s.n_trec * 4, cudaMemcpyDeviceToHost, s.stream));
    }
    CK(cudaStreamSynchronize(s.stream));
    return 0;
}

/* Copy of words [word_off, word_off + n_words) of the slot's staged batch buffer (tests, diagnostics). */
extern "C" int spsp_batch_download(spsp_ctx *c, int slot, uint64_t word_off, uint32_t *host_words, uint64_t n_words)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_batch_download: bad ctx/slot");
    Slot &s = c->slots[slot];
    if (word_off + n_words > s.d_packed_words) return fail(-3, "spsp_batch_download: outside the reserved buffer");
    CK(cudaSetDevice(c->device));
    if (n_words) CK(cudaMemcpyAsync(host_words, s.d_packed + word_off, n_words * 4, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    return 0;
}

// ------------------------------------------------------------ dense totals

static int check_records(const uint64_t *rec_begin, const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec,
                         uint64_t n_bases, uint32_t n_inputs, const char *who)
{
    if (n_rec && (!rec_begin || !rec_end || !rec_input)) return fail(-3, std::string(who) + ": null record arrays");
    for (uint64_t r = 0; r < n_rec; r++) {
        if (rec_begin[r] > rec_end[r] || rec_end[r] > n_bases || rec_input[r] >= n_inputs ||
            (r && (rec_begin[r] < rec_end[r - 1] || rec_input[r] < rec_input[r - 1])))
            return fail(-3, std::string(who) + ": records must be ascending, disjoint and inside the buffer");
    }
    return 0;
}

static int dense_impl(spsp_ctx *c, Slot &s, const uint32_t *d_packed, uint64_t n_bases, const uint64_t *rec_begin,
                      const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec, uint32_t n_inputs,
                      uint64_t *total_superkmers, uint64_t *selected_kmers, float *kernel_ms)
{
    cudaStream_t st = s.stream;
    const bool on_device = n_rec && is_device_ptr(rec_begin);         // a table that is on the device already is used in place
    if (!on_device) {
        int rc = check_records(rec_begin, rec_end, rec_input, n_rec, n_bases, n_inputs, "spsp_dense_stats");
        if (rc) return rc;
        const size_t nr = n_rec ? n_rec : 1;
        CK(s.b_rec_begin.ensure(nr * 8)); CK(s.b_rec_end.ensure(nr * 8)); CK(s.b_rec_input.ensure(nr * 4));
        if (n_rec) {
            CK(cudaMemcpyAsync(s.b_rec_begin.p, rec_begin, n_rec * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(s.b_rec_end.p, rec_end, n_rec * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(s.b_rec_input.p, rec_input, n_rec * 4, cudaMemcpyHostToDevice, st));
        }
    }
    if (!s.dn) s.dn = dense_buffers_create();
    DenseIn in{};
    in.d_packed = d_packed; in.n_bases = n_bases;
    in.d_rec_begin = on_device ? rec_begin : static_cast<const uint64_t *>(s.b_rec_begin.p);
    in.d_rec_end = on_device ? rec_end : static_cast<const uint64_t *>(s.b_rec_end.p);
    in.d_rec_input = on_device ? rec_input : static_cast<const uint32_t *>(s.b_rec_input.p);
    in.n_rec = n_rec; in.n_inputs = n_inputs; in.k = c->k; in.m = c->m; in.thr = c->thr;
    uint32_t nl = 0;
    cudaError_t e = dense_stats_run(s.dn, in, total_superkmers, selected_kmers, kernel_ms, &nl, st);
    if (e != cudaSuccess) return fail(-1, std::string("dense totals: ") + cudaGetErrorString(e));
    {
        std::lock_guard<std::mutex> g(c->mu);
        c->launches += nl;
    }
    return 0;
}

extern "C" int spsp_dense_stats(spsp_ctx *c, int slot, const uint32_t *packed, uint64_t n_bases, const uint64_t *rec_begin,
                                const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec, uint32_t n_inputs,
                                uint64_t *total_superkmers, uint64_t *selected_kmers, float *kernel_ms)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_dense_stats: bad ctx/slot");
    if (!packed && n_bases) return fail(-3, "spsp_dense_stats: null input");
    Slot &s = c->slots[slot];
    CK(cudaSetDevice(c->device));
    const uint64_t words = spsp_packed_words(n_bases);
    { int rc_ = ensure_packed(s, words); if (rc_) return rc_; }
    CK(cudaMemcpyAsync(s.d_packed, packed, words * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
    return dense_impl(c, s, s.d_packed, n_bases, rec_begin, rec_end, rec_input, n_rec, n_inputs, total_superkmers,
                      selected_kmers, kernel_ms);
}

extern "C" int spsp_dense_stats_device(spsp_ctx *c, int slot, const uint32_t *d_packed, uint64_t n_bases,
                                       const uint64_t *rec_begin, const uint64_t *rec_end, const uint32_t *rec_input,
                                       uint64_t n_rec, uint32_t n_inputs, uint64_t *total_superkmers,
                                       uint64_t *selected_kmers, float *kernel_ms)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_dense_stats_device: bad ctx/slot");
    CK(cudaSetDevice(c->device));
    return dense_impl(c, c->slots[slot], d_packed, n_bases, rec_begin, rec_end, rec_input, n_rec, n_inputs,
                      total_superkmers, selected_kmers, kernel_ms);
}

extern "C" int spsp_dense_stats_staged(spsp_ctx *c, int slot, uint64_t n_bases, const uint64_t *rec_begin,
                                       const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec, uint32_t n_inputs,
                                       uint64_t *total_superkmers, uint64_t *selected_kmers, float *kernel_ms)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_dense_stats_staged: bad ctx/slot");
    Slot &s = c->slots[slot];
    if (spsp_packed_words(n_bases) > s.d_packed_words) return fail(-3, "spsp_dense_stats_staged: nothing staged on this slot");
    CK(cudaSetDevice(c->device));
    if (!rec_begin && !n_rec && s.last_nrec)       // the table of the batch just sketched, as the device saw it
        return dense_impl(c, s, s.d_packed, n_bases, s.last_rb, s.last_re, s.last_ri, s.last_nrec, n_inputs, total_superkmers,
                          selected_kmers, kernel_ms);
    return dense_impl(c, s, s.d_packed, n_bases, rec_begin, rec_end, rec_input, n_rec, n_inputs, total_superkmers,
                      selected_kmers, kernel_ms);
}

extern "C" int spsp_cmp_load_batch(spsp_ctx *c, int slot)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_cmp_load_batch: bad ctx/slot");
    Slot &s = c->slots[slot];
    if (!s.has_batch) return fail(-3, "spsp_cmp_load_batch: no batch on this slot");
    return spsp_cmp_load_device(c, s.last_batch_inputs, s.last_batch.h_elem_off, s.last_batch.d_minim, s.last_batch.d_klo,
                                s.last_batch.d_khi);
}

extern "C" int spsp_batch_elements(spsp_ctx *c, int slot, uint32_t *minimizer, uint64_t *kmer_lo, uint64_t *kmer_hi,
                                   const uint32_t **d_minimizer, const uint64_t **d_kmer_lo, const uint64_t **d_kmer_hi)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return fail(-3, "spsp_batch_elements: bad ctx/slot");
    Slot &s = c->slots[slot];
    if (!s.has_batch) return fail(-3, "spsp_batch_elements: no batch on this slot");
    CK(cudaSetDevice(c->device));
    const uint64_t n = s.last_batch.n_elems;
    if (n) {
        if (minimizer) CK(cudaMemcpyAsync(minimizer, s.last_batch.d_minim, n * 4, cudaMemcpyDeviceToHost, s.stream));
        if (kmer_lo) CK(cudaMemcpyAsync(kmer_lo, s.last_batch.d_klo, n * 8, cudaMemcpyDeviceToHost, s.stream));
        if (kmer_hi && s.last_batch.d_khi) CK(cudaMemcpyAsync(kmer_hi, s.last_batch.d_khi, n * 8, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
    }
    if (d_minimizer) *d_minimizer = s.last_batch.d_minim;
    if (d_kmer_lo) *d_kmer_lo = s.last_batch.d_klo;
    if (d_kmer_hi) *d_kmer_hi = s.last_batch.d_khi;
    return 0;
}

// ------------------------------------------------------- multi-GPU exchange
//
// One process per GPU.  NCCL is loaded at run time (dlopen of libnccl.so.2: the copy the process already has,
// e.g. PyTorch's, or the system one), so single-GPU users never touch it.

namespace {
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.h = h;
#define SPSP_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name))
        SPSP_SYM(GetUniqueId, "ncclGetUniqueId"); SPSP_SYM(CommInitRank, "ncclCommInitRank");
        SPSP_SYM(CommDestroy, "ncclCommDestroy"); SPSP_SYM(AllGather, "ncclAllGather"); SPSP_SYM(Send, "ncclSend");
        SPSP_SYM(Recv, "ncclRecv"); SPSP_SYM(GroupStart, "ncclGroupStart"); SPSP_SYM(GroupEnd, "ncclGroupEnd");
        SPSP_SYM(GetErrorString, "ncclGetErrorString");
#undef SPSP_SYM
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.Send || !api.Recv ||
            !api.GroupStart || !api.GroupEnd || !api.GetErrorString)
            api.h = nullptr;
    });
    return api.h ? &api : nullptr;
}
}  // namespace

#define NK(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) return fail(-1, std::string(#call) + ": " + nccl_api()->GetErrorString(r_)); \
    } while (0)

static void nccl_release(spsp_ctx *c)
{
    if (c->comm && nccl_api()) nccl_api()->CommDestroy(c->comm);
    c->comm = nullptr;
    c->x_hdrsz.release(); c->x_klo.release(); c->x_khi.release(); c->x_min.release();
    c->x_begin.release(); c->x_end.release(); c->x_compact.release(); c->x_units.release(); c->x_dims.release();
    c->x_tiles.release(); c->x_recv.release(); c->x_neg.release();
    c->xp_hdrsz.release(); c->xp_out.release(); c->xp_neg.release();
}

extern "C" int spsp_nccl_unique_id(uint8_t *id128)
{
    if (!id128) return fail(-3, "spsp_nccl_unique_id: null buffer");
    NcclApi *n = nccl_api();
    if (!n) return fail(-1, "spsp_nccl_unique_id: libnccl.so.2 not found");
    ncclUniqueId id;
    NK(n->GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return 0;
}

extern "C" int spsp_nccl_init(spsp_ctx *c, const uint8_t *id128, int rank, int world)
{
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return fail(-3, "spsp_nccl_init: bad args");
    if (world > CMP_MAX_WORLD) return fail(-3, "spsp_nccl_init: at most 64 ranks");
    NcclApi *n = nccl_api();
    if (!n) return fail(-1, "spsp_nccl_init: libnccl.so.2 not found");
    CK(cudaSetDevice(c->device));
    if (c->comm) { n->CommDestroy(c->comm); c->comm = nullptr; }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NK(n->CommInitRank(&c->comm, world, id, rank));
    c->rank = rank; c->world = world;
    c->x_ncap = c->x_qcap = c->x_ecap = 0;
    return 0;
}

static uint64_t slot_cap(uint64_t v) { return std::max<uint64_t>(16, (v + v / 4 + 15) & ~(uint64_t)15); }

// Units a rank can be dealt at most for the given slot capacities (same arithmetic on every rank).
static uint64_t units_cap_for(uint32_t W, uint64_t n_cap, uint64_t q_cap, bool symmetric, uint32_t rt)
{
    const uint64_t nJ = (W * n_cap + 31) / 32, nI = symmetric ? nJ : (W * q_cap + 31) / 32;
    uint64_t total = 0;
    if (symmetric) {
        const uint64_t full = nJ / rt, rem = nJ % rt;
        total = rt * full * (full + 1) / 2 + rem * (full + 1);
    } else {
        total = nJ * ((nI + rt - 1) / rt);
    }
    return (total + W - 1) / W;
}

// The compare stage's only exchange step.  Every rank brings n_local sketches (the first q_local of them queries)
// as device arrays + a host size list; the union is compared (all-vs-all when symmetric, else queries x all) and
// rank 0 receives the matrix.  ONE payload all-gather into fixed-capacity slots (the counts travel inside the
// payload; the capacities were agreed once and are re-derived by every rank from the same headers when a rank
// outgrows them), the plan and the join on the device, one grouped send/receive of the owned tiles to rank 0, one
// host synchronisation.  Every decision that changes the sequence of collectives is taken from the gathered
// headers, which every rank sees: no rank can leave the others waiting in a collective.
static int exchange_impl(spsp_ctx *c, uint64_t n_local, uint64_t q_local, const uint64_t *h_sizes, const uint32_t *d_min,
                         const uint64_t *d_klo, const uint64_t *d_khi, bool symmetric, uint32_t *inter_out, uint64_t ld,
                         uint64_t *sizes_out, uint32_t cap_rows, uint32_t cap_cols, uint32_t *n_rows, uint32_t *n_cols,
                         float *kernel_ms)
{
    NcclApi *n = nccl_api();
    cudaStream_t st = c->slots[0].stream;
    const uint32_t W = (uint32_t)c->world, R = (uint32_t)c->rank;
    const bool hi = c->k > 32;
    const uint32_t rt_max = hi ? CMP_RT_HI : CMP_RT;                 // also the tile slots per unit in the compact output
    uint64_t e_local = 0;
    for (uint64_t i = 0; i < n_local; i++) e_local += h_sizes[i];
    static const bool timing = getenv("SPSP_XCHG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();

    if (!c->x_ncap) {
        // first call on this communicator: agree on the slot capacities (one small all-gather + synchronisation)
        CK(c->x_neg.ensure(3 * W * 8)); CK(c->xp_neg.ensure(3 * W * 8));
        uint64_t *h = static_cast<uint64_t *>(c->xp_neg.p), *d = static_cast<uint64_t *>(c->x_neg.p);
        h[3 * R] = n_local; h[3 * R + 1] = q_local; h[3 * R + 2] = e_local;
        CK(cudaMemcpyAsync(d + 3 * R, h + 3 * R, 24, cudaMemcpyHostToDevice, st));
        NK(n->AllGather(d + 3 * R, d, 3, ncclUint64, c->comm, st));
        CK(cudaMemcpyAsync(h, d, 3 * W * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        uint64_t mn = 0, mq = 0, me = 0;
        for (uint32_t r = 0; r < W; r++) { mn = std::max(mn, h[3 * r]); mq = std::max(mq, h[3 * r + 1]); me = std::max(me, h[3 * r + 2]); }
        c->x_ncap = slot_cap(mn); c->x_qcap = slot_cap(mq); c->x_ecap = slot_cap(me);
    }
    for (int attempt = 0; attempt < 3; attempt++) {
        const uint64_t n_cap = c->x_ncap, q_cap = c->x_qcap, e_cap = c->x_ecap;
        const uint64_t stride = XHDR_WORDS + n_cap, N_cap = W * n_cap;
        const bool over = n_local > n_cap || q_local > q_cap || e_local > e_cap;
        // row tiles per unit: fewer than rt_max when the job is too small to give every SM of every rank a CTA
        // (derived from the agreed capacities, so every rank picks the same value)
        const double avg0 = n_local ? (double)e_local / (double)n_local : 1.0;
        const uint64_t chunks0 = std::min<uint64_t>(std::max<uint64_t>((uint64_t)std::ceil(32.0 * (double)e_cap / (double)n_cap / (0.7 * CMP_CAP)), 1), 8192);
        (void)avg0;
        uint32_t rt = rt_max;
        while (rt > 1 && units_cap_for(W, n_cap, q_cap, symmetric, rt) * chunks0 < 2ull * 148) rt /= 2;
        const uint64_t u_cap = units_cap_for(W, n_cap, q_cap, symmetric, rt);
        const uint64_t tile_words = u_cap * rt_max * 1024;                 // uint32 per rank
        // ---- pack this rank's slot
        CK(c->x_hdrsz.ensure(W * stride * 8)); CK(c->xp_hdrsz.ensure(stride * 8));
        CK(c->x_klo.ensure(W * e_cap * 8)); CK(c->x_min.ensure(W * e_cap * 4));
        if (hi) CK(c->x_khi.ensure(W * e_cap * 8));
        CK(cudaEventSynchronize(c->ev_stage));
        uint64_t *hp = static_cast<uint64_t *>(c->xp_hdrsz.p);
        uint64_t flags = (over ? 1u : 0u) | (symmetric ? 2u : 0u);
        if (R == 0 && !inter_out) flags |= 4u;
        hp[0] = n_local; hp[1] = q_local; hp[2] = e_local; hp[3] = flags;
        for (uint64_t i = 0; i < n_cap; i++) hp[XHDR_WORDS + i] = (!over && i < n_local) ? h_sizes[i] : 0;
        uint64_t *d_hs = static_cast<uint64_t *>(c->x_hdrsz.p);
        uint64_t *g_klo = static_cast<uint64_t *>(c->x_klo.p), *g_khi = static_cast<uint64_t *>(c->x_khi.p);
        uint32_t *g_min = static_cast<uint32_t *>(c->x_min.p);
        CK(cudaMemcpyAsync(d_hs + R * stride, hp, stride * 8, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(c->ev_stage, st));
        if (!over && e_local) {
            CK(cudaMemcpyAsync(g_klo + R * e_cap, d_klo, e_local * 8, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(g_min + R * e_cap, d_min, e_local * 4, cudaMemcpyDeviceToDevice, st));
            if (hi) CK(cudaMemcpyAsync(g_khi + R * e_cap, d_khi, e_local * 8, cudaMemcpyDeviceToDevice, st));
        }
        // ---- the exchange: one group
        NK(n->GroupStart());
        NK(n->AllGather(d_hs + R * stride, d_hs, stride, ncclUint64, c->comm, st));
        NK(n->AllGather(g_klo + R * e_cap, g_klo, e_cap, ncclUint64, c->comm, st));
        NK(n->AllGather(g_min + R * e_cap, g_min, e_cap, ncclUint32, c->comm, st));
        if (hi) NK(n->AllGather(g_khi + R * e_cap, g_khi, e_cap, ncclUint64, c->comm, st));
        NK(n->GroupEnd());
        const double t1 = now();
        // ---- plan + join on the device
        CK(c->x_begin.ensure(N_cap * 8)); CK(c->x_end.ensure(N_cap * 8)); CK(c->x_compact.ensure(N_cap * 8));
        CK(c->x_units.ensure(u_cap * sizeof(uint2))); CK(c->x_dims.ensure(32));
        CK(c->x_tiles.ensure(tile_words * 4));
        uint32_t *d_dims = static_cast<uint32_t *>(c->x_dims.p);
        uint2 *d_units = static_cast<uint2 *>(c->x_units.p);
        // rank 0 keeps the tiles of every rank side by side (its own in slot 0) and assembles the matrix on the device
        uint32_t *d_recv = nullptr;
        if (R == 0) {
            CK(c->x_recv.ensure((size_t)W * tile_words * 4));
            d_recv = static_cast<uint32_t *>(c->x_recv.p);
        }
        uint32_t *d_tiles = R == 0 ? d_recv : static_cast<uint32_t *>(c->x_tiles.p);
        CK(cudaMemsetAsync(d_tiles, 0, tile_words * 4, st));
        CK(launch_exchange_plan(d_hs, W, R, n_cap, e_cap, symmetric ? 1 : 0, rt, (uint32_t)u_cap,
                                static_cast<uint64_t *>(c->x_begin.p), static_cast<uint64_t *>(c->x_end.p),
                                static_cast<uint64_t *>(c->x_compact.p), d_units, d_dims, st));
        c->owns_cmp = false;
        c->has_hi = hi;
        c->n_sketches = 0;                                                 // the gathered set is not a loaded set
        c->cmp.minim = g_min; c->cmp.klo = g_klo; c->cmp.khi = hi ? g_khi : nullptr;
        c->cmp.sk_off = static_cast<const uint64_t *>(c->x_begin.p);
        c->cmp.sk_end = static_cast<const uint64_t *>(c->x_end.p);
        const double avg = n_local ? (double)e_local / (double)n_local : 1.0;
        uint64_t chunks = (uint64_t)std::ceil(32.0 * avg / (0.7 * CMP_CAP));
        chunks = std::min<uint64_t>(std::max<uint64_t>(chunks, 1), 8192);
        c->n_chunks = (uint32_t)chunks;
        CK(c->b_chunk_off.ensure((size_t)N_cap * (chunks + 1) * sizeof(uint64_t)));
        c->cmp.chunk_off = static_cast<uint64_t *>(c->b_chunk_off.p);
        CK(launch_chunk_offsets(c->cmp, (uint32_t)N_cap, d_dims, c->n_chunks, c->m, st));
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        uint64_t groups = (2ull * sms + u_cap - 1) / u_cap;
        groups = std::min<uint64_t>(std::max<uint64_t>(groups, 1), chunks);
        CK(cudaEventRecord(c->cev0, st));
        CK(launch_hashjoin(c->cmp, hi, d_units, (uint32_t)u_cap, d_dims, c->n_chunks, (uint32_t)groups, 0, 0, 0, 0, d_tiles, 0, st));
        CK(cudaEventRecord(c->cev1, st));
        c->cmp_timed = true;
        c->launches += 3;
        // ---- owned tiles to rank 0, matrix assembled there on the device; everything the host needs in one sweep,
        // ONE synchronisation
        const uint64_t rows_cap = symmetric ? N_cap : W * q_cap;
        const size_t meta_bytes = W * stride * 8 + N_cap * 8 + 32;
        const size_t mat_bytes = R == 0 ? (size_t)rows_cap * N_cap * 4 : 0;
        CK(c->xp_out.ensure(meta_bytes + mat_bytes));
        uint8_t *h_out = static_cast<uint8_t *>(c->xp_out.p);
        if (W > 1) {
            NK(n->GroupStart());
            if (R == 0) {
                for (uint32_t r = 1; r < W; r++) NK(n->Recv(d_recv + (size_t)r * tile_words, tile_words, ncclUint32, (int)r, c->comm, st));
            } else {
                NK(n->Send(d_tiles, tile_words, ncclUint32, 0, c->comm, st));
            }
            NK(n->GroupEnd());
        }
        CK(cudaMemcpyAsync(h_out, d_hs, W * stride * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_out + W * stride * 8, c->x_compact.p, N_cap * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_out + W * stride * 8 + N_cap * 8, d_dims, 32, cudaMemcpyDeviceToHost, st));
        if (R == 0) {
            CK(c->x_tiles.ensure(mat_bytes));                              // (rank 0 has no other use for it)
            uint32_t *d_mat = static_cast<uint32_t *>(c->x_tiles.p);
            CK(cudaMemsetAsync(d_mat, 0, mat_bytes, st));
            CK(launch_exchange_assemble(d_recv, tile_words, d_dims, W, symmetric ? 1 : 0, rt, rt_max, (uint32_t)rows_cap,
                                        (uint32_t)N_cap, d_mat, N_cap, st));
            c->launches += 1;
            CK(cudaMemcpyAsync(h_out + meta_bytes, d_mat, mat_bytes, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        const double t2 = now();
        // ---- every rank reads the same headers: same verdict everywhere
        const uint64_t *hh = reinterpret_cast<const uint64_t *>(h_out);
        bool any_over = false, bad_args = false, mode_mismatch = false;
        uint64_t mn = 0, mq = 0, me = 0;
        for (uint32_t r = 0; r < W; r++) {
            const uint64_t *h = hh + r * stride;
            mn = std::max(mn, h[0]); mq = std::max(mq, h[1]); me = std::max(me, h[2]);
            any_over |= (h[3] & 1u) != 0; bad_args |= (h[3] & 4u) != 0;
            mode_mismatch |= ((h[3] & 2u) != 0) != symmetric;
        }
        if (any_over) {                                                    // some rank outgrew its slot: larger slots, again
            c->x_ncap = std::max(c->x_ncap, slot_cap(mn)); c->x_qcap = std::max(c->x_qcap, slot_cap(mq));
            c->x_ecap = std::max(c->x_ecap, slot_cap(me));
            continue;
        }
        if (bad_args) return fail(-3, "spsp_cmp_exchange: rank 0 passed no output arrays");
        if (mode_mismatch) return fail(-3, "spsp_cmp_exchange: the ranks disagree on all-vs-all / query mode");
        const uint32_t *dims = reinterpret_cast<const uint32_t *>(h_out + W * stride * 8 + N_cap * 8);
        const uint32_t rows = dims[0], N = dims[1];
        if (dims[4] > dims[2]) return fail(-1, "spsp_cmp_exchange: unit capacity bound violated");
        if (n_rows) *n_rows = rows;
        if (n_cols) *n_cols = N;
        if (kernel_ms) spsp_cmp_last_kernel_ms(c, kernel_ms);
        if (N > cap_cols || rows > cap_rows) return fail(-2, "spsp_cmp_exchange: output arrays too small");
        if (sizes_out) memcpy(sizes_out, h_out + W * stride * 8, (size_t)N * 8);
        if (R == 0) {
            if (ld < N) return fail(-3, "spsp_cmp_exchange: ld too small");
            const uint32_t *mat = reinterpret_cast<const uint32_t *>(h_out + meta_bytes);
            for (uint64_t i = 0; i < rows; i++) memcpy(inter_out + i * ld, mat + i * N_cap, (size_t)N * 4);
        }
        if (timing && R == 0)
            fprintf(stderr, "[xchg] enqueue gather %.0f us | plan+join+collect+sync %.0f | scatter %.0f (N=%u rows=%u units/rank<=%llu)\n",
                    t1 - t0, t2 - t1, now() - t2, N, rows, (unsigned long long)u_cap);
        return 0;
    }
    return fail(-1, "spsp_cmp_exchange: slot capacities did not converge");
}

extern "C" int spsp_cmp_exchange(spsp_ctx *c, uint32_t n_local, uint32_t q_local, const uint64_t *sizes_local,
                                 const uint32_t *d_minimizer, const uint64_t *d_kmer_lo, const uint64_t *d_kmer_hi,
                                 int symmetric, uint32_t *inter_out, uint64_t ld, uint64_t *sizes_out, uint32_t cap_rows,
                                 uint32_t cap_cols, uint32_t *n_rows, uint32_t *n_cols, float *kernel_ms)
{
    if (!c) return fail(-3, "spsp_cmp_exchange: null context");
    if (!nccl_api() || !c->comm) return fail(-3, "spsp_cmp_exchange: call spsp_nccl_init first");
    if (n_local && !sizes_local) return fail(-3, "spsp_cmp_exchange: null size list");
    if (q_local > n_local) return fail(-3, "spsp_cmp_exchange: more queries than sketches");
    if (symmetric && q_local != n_local) return fail(-3, "spsp_cmp_exchange: all-vs-all needs q_local == n_local");
    if ((c->k > 32) && n_local && !d_kmer_hi) return fail(-3, "spsp_cmp_exchange: kmer_hi must be given when k > 32");
    CK(cudaSetDevice(c->device));
    return exchange_impl(c, n_local, q_local, sizes_local, d_minimizer, d_kmer_lo, d_kmer_hi, symmetric != 0, inter_out, ld,
                         sizes_out, cap_rows, cap_cols, n_rows, n_cols, kernel_ms);
}

extern "C" int spsp_cmp_exchange_batch(spsp_ctx *c, int slot, uint32_t *inter_out, uint64_t ld, uint64_t *sizes_out,
                                       uint32_t cap_sketches, uint32_t *n_total, float *kernel_ms)
{
    if (!c || slot < 0 || slot >= (int)c->slots.size() || !n_total) return fail(-3, "spsp_cmp_exchange_batch: bad args");
    if (!nccl_api() || !c->comm) return fail(-3, "spsp_cmp_exchange_batch: call spsp_nccl_init first");
    Slot &s = c->slots[slot];
    if (!s.has_batch) return fail(-3, "spsp_cmp_exchange_batch: no batch on this slot");
    CK(cudaSetDevice(c->device));
    if (slot != 0) CK(cudaStreamSynchronize(s.stream));
    const uint32_t n_local = s.last_batch_inputs;
    std::vector<uint64_t> sz(n_local ? n_local : 1);
    for (uint32_t i = 0; i < n_local; i++) sz[i] = s.last_batch.h_elem_off[i + 1] - s.last_batch.h_elem_off[i];
    uint32_t rows = 0, cols = 0;
    int rc = exchange_impl(c, n_local, n_local, sz.data(), s.last_batch.d_minim, s.last_batch.d_klo, s.last_batch.d_khi, true,
                           inter_out, ld, sizes_out, cap_sketches, cap_sketches, &rows, &cols, kernel_ms);
    *n_total = cols;
    return rc;
}

extern "C" int spsp_launch_count(spsp_ctx *c, uint64_t *n)
{
    if (!c || !n) return fail(-3, "spsp_launch_count: bad args");
    *n = c->launches;
    return 0;
}
