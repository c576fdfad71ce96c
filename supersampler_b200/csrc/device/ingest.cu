// Device-side FASTA ingest (sm_100a): the reference's getLineFasta + clean_dna (utils.cpp:706-718, 675-702) and the
// 2-bit packing (utils.cpp:13-16) as three streaming passes over raw text resident in HBM.
//
//   text semantics (same as the host packer, csrc/host/seqio.cpp):
//     * the first line of an input and every line that starts with '>' -- or with the byte 0xFF, which the
//       reference's `char c = peek(); c != EOF` (utils.cpp:709-713) cannot tell from the end of the file -- is a
//       header line: it starts a record and none of its bytes are sequence;
//     * on every other line the bytes ACGTacgt are bases (code (c >> 1) & 3: A0 C1 T2 G3), every other byte is
//       deleted, so its neighbours become adjacent;
//     * a record is the run of bases between two header lines (it may be empty; records shorter than k are kept
//       in the table and ignored by the post-pass, which asks for `end - begin >= k`).
//
//   pass 1  ingest_summary_kernel : per 16 KB tile, what the tile contributes whatever precedes it
//   pass 2  ingest_carry_kernel   : one warp per input turns the summaries into per-tile prefixes (bases, records,
//                                   "starts inside a header"), ingest_totals_kernel into per-input record bases
//   pass 3  ingest_write_kernel   : per tile, compacts the codes in shared memory (ballot-free: every thread's
//                                   bases form one <= 32-bit run) and stores whole words; writes the record table
//
// A thread owns 16 consecutive bytes (one 128-bit load); per-byte classes are computed 4 bytes at a time with the
// SIMD-in-a-word compare instructions and gathered into 16-bit masks; "inside a header" is a carry-style flood of
// the header-start bits over the non-line-start bits (4 doubling steps), so there is no per-byte branch.
#include "ingest.cuh"

#include "common.cuh"

namespace spsp {

namespace {

struct Masks {
    uint32_t ls;        // byte is the first byte of a line
    uint32_t hs;        // ... of a header line
    uint32_t hdr;       // byte lies in a header line, assuming the bytes before the first line start do not
    uint32_t pre;       // bytes before the thread's first line start
    uint32_t base;      // byte is one of ACGTacgt (inside the input)
    uint32_t nl;
    uint4 w;
};

__device__ __forceinline__ uint32_t gather4(uint32_t m)      // 0xFF/0x00 per byte -> 4 bits
{
    return (((m & 0x01010101u) * 0x00204081u) >> 21) & 0xFu;
}

__device__ __forceinline__ void classify_word(uint32_t w, uint32_t &nl, uint32_t &gt, uint32_t &base)
{
    nl = gather4(__vcmpeq4(w, 0x0A0A0A0Au));
    gt = gather4(__vcmpeq4(w, 0x3E3E3E3Eu) | __vcmpeq4(w, 0xFFFFFFFFu));     // '>' or 0xFF: header when first on a line
    const uint32_t lo = w | 0x20202020u;
    const uint32_t b = __vcmpeq4(lo, 0x61616161u) | __vcmpeq4(lo, 0x63636363u) | __vcmpeq4(lo, 0x67676767u) |
                       __vcmpeq4(lo, 0x74747474u);
    base = gather4(b);
}

// Masks of the thread's 16 bytes [g0, g0 + 16) of an input of `len` bytes; prev_nl: the byte before g0 is '\n'.
__device__ __forceinline__ Masks thread_masks(const uint8_t *text, uint64_t g0, uint64_t len, int lane)
{
    Masks k;
    uint32_t nl = 0, gt = 0, base = 0;
    k.w = make_uint4(0, 0, 0, 0);
    uint32_t inr = 0;
    if (g0 < len) {
        k.w = ld_stream_u4(reinterpret_cast<const uint4 *>(text + g0));
        const uint64_t left = len - g0;
        inr = left >= 16 ? 0xFFFFu : ((1u << (uint32_t)left) - 1u);
        uint32_t a, b, c;
        classify_word(k.w.x, a, b, c); nl |= a; gt |= b; base |= c;
        classify_word(k.w.y, a, b, c); nl |= a << 4; gt |= b << 4; base |= c << 4;
        classify_word(k.w.z, a, b, c); nl |= a << 8; gt |= b << 8; base |= c << 8;
        classify_word(k.w.w, a, b, c); nl |= a << 12; gt |= b << 12; base |= c << 12;
        nl &= inr; gt &= inr; base &= inr;
    }
    // is the byte before this thread's chunk a line end?  the previous lane knows; lane 0 reads it
    uint32_t prev = __shfl_up_sync(0xffffffffu, nl >> 15, 1);
    if (lane == 0) prev = (g0 > 0 && g0 <= len) ? (text[g0 - 1] == '\n') : 0u;
    if (g0 == 0) prev = 1u;
    k.nl = nl;
    k.ls = (((nl << 1) | prev) & 0xFFFFu) & inr;
    k.hs = k.ls & (gt | (g0 == 0 ? 1u : 0u));
    // flood the header-start bits upwards over bytes that do not start a line
    uint32_t g = k.hs, p = ~k.ls & 0xFFFFu;
    g |= p & (g << 1); p &= p << 1;
    g |= p & (g << 2); p &= p << 2;
    g |= p & (g << 4); p &= p << 4;
    g |= p & (g << 8);
    k.hdr = g & 0xFFFFu;
    k.pre = k.ls ? ((k.ls & (0u - k.ls)) - 1u) : 0xFFFFu;
    k.base = base;
    return k;
}

// value of the thread for the "latest line start" max-scan: 0 = no line start, else 1 + (tid << 1 | is header)
__device__ __forceinline__ uint32_t state_token(const Masks &k, uint32_t tid)
{
    if (!k.ls) return 0;
    const uint32_t top = 31 - __clz(k.ls);
    return 1u + ((tid << 1) | ((k.hs >> top) & 1u));
}

// CTA-wide exclusive scans (1024 threads = 32 warps); s_w: 33 words of shared memory; total of the CTA -> *total.
__device__ __forceinline__ uint32_t block_excl_max(uint32_t v, uint32_t *s_w, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, t);
    }
    uint32_t ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = 0;
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = s_w[lane], ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti = max(ti, u);
        }
        uint32_t te = __shfl_up_sync(0xffffffffu, ti, 1);
        if (lane == 0) te = 0;
        s_w[lane] = te;
        if (lane == 31) s_w[32] = ti;
    }
    __syncthreads();
    const uint32_t r = max(s_w[warp], ex);
    *total = s_w[32];
    __syncthreads();
    return r;
}
__device__ __forceinline__ uint32_t block_excl_add(uint32_t v, uint32_t *s_w, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t ex = inc - v;
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = s_w[lane], ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += u;
        }
        s_w[lane] = ti - t;
        if (lane == 31) s_w[32] = ti;
    }
    __syncthreads();
    const uint32_t r = s_w[warp] + ex;
    *total = s_w[32];
    __syncthreads();
    return r;
}

// input that owns `tile` (tile0 ascending; inputs without tiles share the tile0 of their successor)
__device__ __forceinline__ uint32_t find_input(const IngestInput *in, uint32_t n_in, uint64_t tile)
{
    uint32_t lo = 0, hi = n_in;                // first input with tile0 > tile
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (in[mid].tile0 <= tile) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

__global__ void __launch_bounds__(ING_THREADS) ingest_summary_kernel(const uint8_t *__restrict__ text,
                                                                      const IngestInput *__restrict__ in, uint32_t n_in,
                                                                      IngestTile *__restrict__ tiles)
{
    __shared__ uint32_t s_w[33];
    __shared__ uint32_t s_in;
    __shared__ uint32_t s_sum[3];
    const uint64_t tile = blockIdx.x;
    if (threadIdx.x == 0) { s_in = find_input(in, n_in, tile); s_sum[0] = s_sum[1] = s_sum[2] = 0; }
    __syncthreads();
    const IngestInput I = in[s_in];
    const uint64_t g0 = (tile - I.tile0) * (uint64_t)ING_TILE + (uint64_t)threadIdx.x * 16;
    const Masks k = thread_masks(text + I.text_off, g0, I.text_len, threadIdx.x & 31);
    uint32_t total;
    const uint32_t e = block_excl_max(state_token(k, threadIdx.x), s_w, &total);
    uint32_t pre = 0, post = 0;
    if (e) {
        const uint32_t h = k.hdr | (((e - 1u) & 1u) ? k.pre : 0u);
        post = __popc(k.base & ~h);
    } else {
        pre = __popc(k.base & k.pre);
        post = __popc(k.base & ~k.pre & ~k.hdr);
    }
    const uint32_t nh = __popc(k.hs);
    pre = __reduce_add_sync(0xffffffffu, pre);
    post = __reduce_add_sync(0xffffffffu, post);
    const uint32_t nhw = __reduce_add_sync(0xffffffffu, nh);
    if ((threadIdx.x & 31) == 0) {
        if (pre) atomicAdd(&s_sum[0], pre);
        if (post) atomicAdd(&s_sum[1], post);
        if (nhw) atomicAdd(&s_sum[2], nhw);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        IngestTile t;
        t.cnt_pre = s_sum[0]; t.cnt_post = s_sum[1]; t.n_hdr = s_sum[2];
        t.flags = total ? (1u | (((total - 1u) & 1u) << 1)) : 0u;
        tiles[tile] = t;
    }
}

// grid = inputs, one warp each
__global__ void __launch_bounds__(32) ingest_carry_kernel(const IngestInput *__restrict__ in, const IngestTile *__restrict__ tiles,
                                                           IngestCarry *__restrict__ carry, IngestTotals *__restrict__ tot)
{
    const IngestInput I = in[blockIdx.x];
    const int lane = threadIdx.x;
    const uint64_t nt = (I.text_len + ING_TILE - 1) / ING_TILE;
    uint64_t bases = 0;
    uint32_t recs = 0, state = 1;
    for (uint64_t t0 = 0; t0 < nt; t0 += 32) {
        const uint64_t t = t0 + lane;
        IngestTile s{0, 0, 0, 0};
        if (t < nt) s = tiles[I.tile0 + t];
        const uint32_t tok = (s.flags & 1u) ? 1u + (((uint32_t)lane << 1) | ((s.flags >> 1) & 1u)) : 0u;
        uint32_t inc = tok;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = max(inc, u);
        }
        uint32_t ex = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) ex = 0;
        const uint32_t in_hdr = ex ? ((ex - 1u) & 1u) : state;
        const uint32_t cnt = s.cnt_post + (in_hdr ? 0u : s.cnt_pre);
        uint32_t ci = cnt, hi = s.n_hdr;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, ci, o), b = __shfl_up_sync(0xffffffffu, hi, o);
            if (lane >= o) { ci += a; hi += b; }
        }
        if (t < nt) {
            IngestCarry c;
            c.base_prefix = bases + (ci - cnt); c.hdr_prefix = recs + (hi - s.n_hdr); c.in_header = in_hdr;
            carry[I.tile0 + t] = c;
        }
        bases += __shfl_sync(0xffffffffu, ci, 31);
        recs += __shfl_sync(0xffffffffu, hi, 31);
        const uint32_t last = __shfl_sync(0xffffffffu, inc, 31);
        if (last) state = (last - 1u) & 1u;
    }
    if (lane == 0) { tot[blockIdx.x].n_bases = bases; tot[blockIdx.x].n_rec = recs; }
}

// one CTA: rec_base = exclusive prefix of n_rec over the inputs; grand = {bases, records}
__global__ void __launch_bounds__(ING_THREADS) ingest_totals_kernel(IngestTotals *__restrict__ tot, uint32_t n_in,
                                                                     uint64_t *__restrict__ grand)
{
    __shared__ unsigned long long s_w[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long run_rec = 0, run_bases = 0;
    for (uint32_t i0 = 0; i0 < n_in; i0 += ING_THREADS) {
        const uint32_t i = i0 + threadIdx.x;
        const unsigned long long v = i < n_in ? tot[i].n_rec : 0ull, b = i < n_in ? tot[i].n_bases : 0ull;
        unsigned long long inc = v, bs = b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
            bs += __shfl_xor_sync(0xffffffffu, bs, o);
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long t = s_w[lane];
            unsigned long long ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += u;
            }
            s_w[lane] = ti - t;
            if (lane == 31) s_w[32] = ti;
        }
        __syncthreads();
        if (i < n_in) tot[i].rec_base = run_rec + s_w[warp] + (inc - v);
        run_rec += s_w[32];
        __syncthreads();
        // bases: plain sum (reuse the scratch)
        if (lane == 0) s_w[warp] = bs;
        __syncthreads();
        if (warp == 0) {
            unsigned long long t = s_w[lane];
#pragma unroll
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) s_w[32] = t;
        }
        __syncthreads();
        run_bases += s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) { grand[0] = run_bases; grand[1] = run_rec; }
}

__global__ void __launch_bounds__(256) ingest_zero_kernel(const IngestInput *__restrict__ in, uint32_t n_in,
                                                           uint32_t *__restrict__ packed)
{
    for (uint32_t ii = blockIdx.y; ii < n_in; ii += gridDim.y) {
        const IngestInput I = in[ii];
        const uint64_t words = 4 * ((I.text_len + 4 + 63) / 64) + 8;          // spsp_packed_words(text_len)
        uint4 *p = reinterpret_cast<uint4 *>(packed + I.word_off);             // word_off is a multiple of 4
        const uint64_t n4 = words / 4;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x)
            p[i] = make_uint4(0, 0, 0, 0);
    }
}

__global__ void __launch_bounds__(ING_THREADS) ingest_write_kernel(const uint8_t *__restrict__ text,
                                                                    const IngestInput *__restrict__ in, uint32_t n_in,
                                                                    const IngestCarry *__restrict__ carry,
                                                                    const IngestTotals *__restrict__ tot,
                                                                    uint32_t *__restrict__ packed,
                                                                    uint64_t *__restrict__ rec_begin,
                                                                    uint64_t *__restrict__ rec_end,
                                                                    uint32_t *__restrict__ rec_input)
{
    __shared__ uint32_t s_w[33];
    __shared__ uint32_t s_in;
    __shared__ uint32_t s_out[ING_THREADS + 4];          // (15 + 16384) bases = 1025 words, + the spill word
    const uint64_t tile = blockIdx.x;
    if (threadIdx.x == 0) s_in = find_input(in, n_in, tile);
    for (int i = threadIdx.x; i < ING_THREADS + 4; i += ING_THREADS) s_out[i] = 0;
    __syncthreads();
    const uint32_t ii = s_in;
    const IngestInput I = in[ii];
    const IngestCarry C = carry[tile];
    const IngestTotals T = tot[ii];
    const uint64_t g0 = (tile - I.tile0) * (uint64_t)ING_TILE + (uint64_t)threadIdx.x * 16;
    const Masks k = thread_masks(text + I.text_off, g0, I.text_len, threadIdx.x & 31);
    uint32_t total;
    const uint32_t e = block_excl_max(state_token(k, threadIdx.x), s_w, &total);
    const uint32_t in_hdr = e ? ((e - 1u) & 1u) : C.in_header;
    const uint32_t h = k.hdr | (in_hdr ? k.pre : 0u);
    const uint32_t V = k.base & ~h;
    const uint32_t cnt = __popc(V), nh = __popc(k.hs);
    const uint32_t ex = block_excl_add(cnt | (nh << 16), s_w, &total);
    const uint32_t r = ex & 0xFFFFu, hr = ex >> 16;
    const uint32_t a = (uint32_t)(C.base_prefix & 15u);
    const uint64_t region = 16ull * I.word_off;

    // ---- this thread's bases: one run of 2 * cnt bits, first base in the most significant position
    if (cnt) {
        uint32_t acc = 0;
        const uint32_t cw[4] = {(k.w.x >> 1) & 0x03030303u, (k.w.y >> 1) & 0x03030303u, (k.w.z >> 1) & 0x03030303u,
                                (k.w.w >> 1) & 0x03030303u};
#pragma unroll
        for (int j = 0; j < 16; j++)
            if ((V >> j) & 1u) acc = (acc << 2) | ((cw[j >> 2] >> (8 * (j & 3))) & 3u);
        const uint32_t P = a + r;
        const uint64_t win = (((uint64_t)acc) << (64 - 2 * cnt)) >> (2 * (P & 15u));
        const uint32_t hi = (uint32_t)(win >> 32), lo = (uint32_t)win;
        atomicOr(&s_out[P >> 4], hi);
        if (lo) atomicOr(&s_out[(P >> 4) + 1], lo);
    }
    // ---- records: every header line starts one at the current cleaned offset and ends the previous one
    if (nh) {
        uint32_t m = k.hs, q = 0;
        while (m) {
            const uint32_t j = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t pos = region + C.base_prefix + r + __popc(V & ((1u << j) - 1u));
            const uint64_t ri = T.rec_base + C.hdr_prefix + hr + q;
            rec_begin[ri] = pos;
            rec_input[ri] = I.input;
            if (ri > T.rec_base) rec_end[ri - 1] = pos;
            q++;
        }
    }
    const uint64_t nt = (I.text_len + ING_TILE - 1) / ING_TILE;
    if (threadIdx.x == 0 && tile - I.tile0 == nt - 1 && T.n_rec) rec_end[T.rec_base + T.n_rec - 1] = region + T.n_bases;
    __syncthreads();
    // ---- whole words to the region; the two words a tile may share with its neighbours are OR-ed in
    const uint32_t tb = total & 0xFFFFu;
    if (tb) {
        const uint32_t nw = (a + tb + 15) >> 4;
        uint32_t *dst = packed + I.word_off + (C.base_prefix >> 4);
        for (uint32_t i = threadIdx.x; i < nw; i += ING_THREADS) {
            const uint32_t v = s_out[i];
            const bool shared_word = (i == 0 && a) || (i == nw - 1 && ((a + tb) & 15u));
            if (shared_word) { if (v) atomicOr(dst + i, v); }
            else dst[i] = v;
        }
    }
}

__global__ void record_merge_kernel(const uint64_t *__restrict__ a_begin, const uint64_t *__restrict__ a_end,
                                    const uint32_t *__restrict__ a_input, uint64_t na, const uint64_t *__restrict__ b_begin,
                                    const uint64_t *__restrict__ b_end, const uint32_t *__restrict__ b_input, uint64_t nb,
                                    uint64_t *__restrict__ o_begin, uint64_t *__restrict__ o_end, uint32_t *__restrict__ o_input)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na + nb) return;
    const bool from_a = i < na;
    const uint64_t self = from_a ? i : i - na;
    const uint64_t key = from_a ? a_begin[self] : b_begin[self];
    const uint64_t *other = from_a ? b_begin : a_begin;
    uint64_t lo = 0, hi = from_a ? nb : na;
    // a: records of b that begin before it; b: records of a that begin at or before it (no cross ties in
    // practice: the two tables describe different inputs, whose regions are disjoint)
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        const bool before = from_a ? other[mid] < key : other[mid] <= key;
        if (before) lo = mid + 1; else hi = mid;
    }
    const uint64_t o = self + lo;
    o_begin[o] = key;
    o_end[o] = from_a ? a_end[self] : b_end[self];
    o_input[o] = from_a ? a_input[self] : b_input[self];
}

// k-mers of every input: sum over its records of at least k bases of (length - k + 1)  (read_kmer, SubSampler.cpp:343-347)
__global__ void record_kmers_kernel(const uint64_t *__restrict__ rec_begin, const uint64_t *__restrict__ rec_end,
                                    const uint32_t *__restrict__ rec_input, uint64_t n_rec, uint32_t k,
                                    unsigned long long *__restrict__ kmers)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    uint32_t in = 0xFFFFFFFFu;
    if (r < n_rec) {
        const uint64_t n = rec_end[r] - rec_begin[r];
        in = rec_input[r];
        if (n >= k) v = n - k + 1;
    }
    // records are grouped by input: add up the lanes that share the first lane's input, the others go alone
    const uint32_t in0 = __shfl_sync(0xffffffffu, in, 0);
    const bool same = in == in0;
    unsigned long long part = same ? v : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0 && in0 != 0xFFFFFFFFu && part) atomicAdd(kmers + in0, part);
    if (!same && in != 0xFFFFFFFFu && v) atomicAdd(kmers + in, v);
}

}  // namespace

cudaError_t launch_record_kmers(const uint64_t *rec_begin, const uint64_t *rec_end, const uint32_t *rec_input, uint64_t n_rec,
                                uint32_t k, unsigned long long *d_kmers, cudaStream_t st)
{
    if (!n_rec) return cudaSuccess;
    record_kmers_kernel<<<(unsigned)((n_rec + 255) / 256), 256, 0, st>>>(rec_begin, rec_end, rec_input, n_rec, k, d_kmers);
    return cudaGetLastError();
}

cudaError_t launch_ingest_summary(const uint8_t *d_text, const IngestInput *d_in, uint32_t n_in, uint64_t n_tiles,
                                  IngestTile *d_tiles, cudaStream_t st)
{
    if (!n_tiles) return cudaSuccess;
    if (n_tiles > 0x7fffffffull) return cudaErrorInvalidValue;
    ingest_summary_kernel<<<(unsigned)n_tiles, ING_THREADS, 0, st>>>(d_text, d_in, n_in, d_tiles);
    return cudaGetLastError();
}

cudaError_t launch_ingest_carry(const IngestInput *d_in, uint32_t n_in, const IngestTile *d_tiles, IngestCarry *d_carry,
                                IngestTotals *d_tot, uint64_t *d_grand, cudaStream_t st)
{
    if (!n_in) return cudaSuccess;
    ingest_carry_kernel<<<n_in, 32, 0, st>>>(d_in, d_tiles, d_carry, d_tot);
    ingest_totals_kernel<<<1, ING_THREADS, 0, st>>>(d_tot, n_in, d_grand);
    return cudaGetLastError();
}

cudaError_t launch_ingest_write(const uint8_t *d_text, const IngestInput *d_in, uint32_t n_in, uint64_t n_tiles,
                                const IngestCarry *d_carry, const IngestTotals *d_tot, uint32_t *d_packed,
                                uint64_t *d_rec_begin, uint64_t *d_rec_end, uint32_t *d_rec_input, cudaStream_t st)
{
    if (!n_in) return cudaSuccess;
    ingest_zero_kernel<<<dim3(64, n_in < 65535u ? n_in : 65535u), 256, 0, st>>>(d_in, n_in, d_packed);
    if (n_tiles)
        ingest_write_kernel<<<(unsigned)n_tiles, ING_THREADS, 0, st>>>(d_text, d_in, n_in, d_carry, d_tot, d_packed, d_rec_begin,
                                                                        d_rec_end, d_rec_input);
    return cudaGetLastError();
}

cudaError_t launch_record_merge(const uint64_t *a_begin, const uint64_t *a_end, const uint32_t *a_input, uint64_t na,
                                const uint64_t *b_begin, const uint64_t *b_end, const uint32_t *b_input, uint64_t nb,
                                uint64_t *o_begin, uint64_t *o_end, uint32_t *o_input, cudaStream_t st)
{
    const uint64_t n = na + nb;
    if (!n) return cudaSuccess;
    record_merge_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a_begin, a_end, a_input, na, b_begin, b_end, b_input, nb,
                                                                      o_begin, o_end, o_input);
    return cudaGetLastError();
}

}  // namespace spsp
