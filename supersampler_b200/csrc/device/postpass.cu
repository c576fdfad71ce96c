// Device post-pass of the sketch stage (see postpass.cuh): from the scan kernel's
// unordered hit list to sketch bytes + compare elements, in a handful of launches
// and without a host round trip until the results are complete:
//
//   hits ──S1 classify + sort by position──▶ R1/R2 sparse replay (count, write) ──▶ pieces
//        ──S2 stable sort of the PIECES by bucket (input, minimizer)──▶ B bucket-group kernel:
//        a CTA stages a run of whole buckets in shared memory and does everything the
//        reference does per bucket there (k-mer entries, first-occurrence de-dup with uint8
//        counts, neighbour table, greedy chains, byte sizes, canonical elements) and writes the
//        sketch bytes and the compare elements -- small batches at their final place, found by
//        decoupled look-back; large batches to ranges drawn from bump allocators that
//        pp_offsets_kernel + pp_gather_kernel then put in bucket order (see BucketArgs).
//
// Sorting pieces (one per super-k-mer) instead of k-mer entries (k-m+1 times as many) and
// de-duplicating inside a shared-memory table instead of a global one is what makes the pass
// scale: a bucket's pieces are adjacent after the sort, in the reference's insertion order
// (the sort is stable and pieces are produced in genome order).
//
// Everything that carries reference semantics is a kernel in this file, each citing the
// reference lines it restates (via csrc/host/postpass.cpp, which it mirrors).  The only
// library piece is cub::DeviceRadixSort for batches above 32 K keys (smaller ones use the single-CTA
// counting sort of this file).
#include "postpass.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cub/cub.cuh>

#include "common.cuh"

namespace spsp {

// ------------------------------------------------------------------ helpers

struct K128 {
    uint64_t lo, hi;
};
__device__ __forceinline__ bool k_lt(const K128 &a, const K128 &b) { return a.hi != b.hi ? a.hi < b.hi : a.lo < b.lo; }
__device__ __forceinline__ K128 k_shr(K128 a, int s)
{
    if (s == 0) return a;
    if (s >= 64) return K128{s >= 128 ? 0 : a.hi >> (s - 64), 0};
    return K128{(a.lo >> s) | (a.hi << (64 - s)), a.hi >> s};
}
__device__ __forceinline__ K128 k_shl(K128 a, int s)
{
    if (s == 0) return a;
    if (s >= 64) return K128{0, s >= 128 ? 0 : a.lo << (s - 64)};
    return K128{a.lo << s, (a.hi << s) | (a.lo >> (64 - s))};
}
__device__ __forceinline__ uint64_t rc_bits64(uint64_t x)
{
    uint64_t r = __brevll(x);
    r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
    return r ^ 0xAAAAAAAAAAAAAAAAULL;
}
// reverse complement of a k-mer held right-aligned in 128 bits (utils.cpp:397-438 rcb)
__device__ __forceinline__ K128 k_rc(K128 a, int k)
{
    K128 r{rc_bits64(a.hi), rc_bits64(a.lo)};
    return k_shr(r, 128 - 2 * k);
}
template <bool HI> __device__ __forceinline__ K128 k_rc_t(K128 a, int k)
{
    if (!HI) return K128{rc_bits64(a.lo) >> (64 - 2 * k), 0};              // k <= 32
    return k_rc(a, k);
}
// k <= 32: the k-mer fits three words
__device__ __forceinline__ uint64_t kmer_at64(const uint32_t *__restrict__ w, uint64_t pos, int k)
{
    const uint64_t i = pos >> 4;
    const int o = (int)(pos & 15);
    const uint32_t w0 = __ldg(w + i), w1 = __ldg(w + i + 1), w2 = __ldg(w + i + 2);
    const uint32_t t0 = __funnelshift_l(w1, w0, 2 * o), t1 = __funnelshift_l(w2, w1, 2 * o);
    return (((uint64_t)t0 << 32) | t1) >> (64 - 2 * k);
}
// k-mer starting at global base `pos`, right-aligned (first base most significant)
__device__ __forceinline__ K128 kmer_at(const uint32_t *__restrict__ w, uint64_t pos, int k)
{
    const uint64_t i = pos >> 4;
    const int o = (int)(pos & 15);
    uint32_t W[5];
#pragma unroll
    for (int j = 0; j < 5; j++) W[j] = __ldg(w + i + j);
    uint32_t T[4];
#pragma unroll
    for (int j = 0; j < 4; j++) T[j] = __funnelshift_l(W[j + 1], W[j], 2 * o);
    K128 top{((uint64_t)T[2] << 32) | T[3], ((uint64_t)T[0] << 32) | T[1]};
    return k_shr(top, 128 - 2 * k);
}
// last record r with rec_begin[r] <= pos, -1 if none
__device__ __forceinline__ long long find_rec(const uint64_t *__restrict__ rec_begin, uint64_t n_rec, uint64_t pos)
{
    uint64_t lo = 0, hi = n_rec;           // first index with begin > pos
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (rec_begin[mid] <= pos) lo = mid + 1; else hi = mid;
    }
    return (long long)lo - 1;
}

// counters kept on the device for the whole pass (one D2H at the end)
struct Counters {
    unsigned long long n_valid, n_pieces, n_entries, n_buckets, n_elems, body_bytes;
    unsigned long long big_top;        // bump allocator of the big-group pool (bytes)
    unsigned long long tmp_bytes, tmp_elems;   // bump allocators of the groups' temporary output ranges
    unsigned int overflow;             // OVF_* bits: a capacity was too small, the host retries
    unsigned int pad;
};
enum { OVF_PIECES = 1, OVF_BODY = 2, OVF_ELEMS = 4, OVF_BIG = 8, OVF_SORT = 16 };

constexpr uint32_t KEY_INVALID = 0xFFFFFFFFu;
constexpr int RP_THREADS = 256;      // replay kernels
// Bucket-group kernel: CTA size and group size go together (a group's phases are separated by CTA barriers and
// several of them keep only a few lanes busy -- the chain walk has one lane per bucket -- so what fills an SM is the
// number of resident groups, which shared memory bounds).  SPSP_BK_SCALE = 1, 2, 4 (compile time) divides both;
// measured on B200 (profiles/README.md): 1 is best -- a group's phases cost about the same cycles whatever its
// size (each is a chain of dependent shared-memory accesses per thread), so smaller groups only add groups:
// 64 x 5 Mbp batch 0.27 / 0.27 / 0.40 ms, 1 Gbp batch at s=100 3.7 / 4.2 / 7.6 ms for scale 1 / 2 / 4.
#ifndef SPSP_BK_SCALE
#define SPSP_BK_SCALE 1
#endif
constexpr int BK_THREADS = 256 / SPSP_BK_SCALE;      // bucket-group kernel
constexpr int BK_WARPS = BK_THREADS / 32;
constexpr int BK_ECAP = 1024 / SPSP_BK_SCALE;        // entries of a group staged in shared memory
constexpr int BK_PMAX = 128 / SPSP_BK_SCALE;         // pieces of such a group
constexpr int BK_SLOTS = 2048 / SPSP_BK_SCALE;       // its open-addressing table (load <= 0.5)
constexpr int BK_PP_ENTRIES = 600 / SPSP_BK_SCALE;   // entries a group aims at (pieces per group = this / (k - m + 1))
static_assert(BK_WARPS >= 2, "the bucket kernel needs a look-back warp and at least one writer warp");
constexpr int RC_SK = 192;           // 2-bit codes of a super-k-mer (2k-m <= 123, grows both ways from 64)
constexpr uint32_t H_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t VIS_START = 0, VIS_LEFT = 1, VIS_RIGHT = 2;

// ------------------------------------------------------- S1 / S2 sorts, small path
//
// Up to 32 K keys: ONE CTA, a counting sort on the top 13 key bits in shared memory (hit positions are spread
// over the batch and selected minimizers are hash-uniform, so bins hold a few keys), an insertion sort inside
// each bin -- every key counts the keys of its bin before it, on (key, original index), which makes the sort
// stable -- and each key is written to its final place.  HITS: loads the scan's hit records, classifies them (a hit counts only if its m-mer lies inside
// one record of at least k bases: host build_sketch_t drops hits that straddle records / records < k) and
// leaves positions and (canon << 1 | rev) sorted by position.  !HITS: sorts the pieces' bucket keys, leaving the
// permutation.
constexpr int SS_THREADS = 1024, SS_CAP = 32768, SS_BIN_BITS = 13, SS_BINS = 1 << SS_BIN_BITS;
constexpr size_t SS_SMEM = (size_t)SS_CAP * 4 + (size_t)SS_CAP * 2 + (size_t)(SS_BINS + 32) * 4;
constexpr int SS_MLP = 4;            // keys a thread handles at a time (their loads are in flight together)

template <bool HITS>
__global__ void __launch_bounds__(SS_THREADS, 1)
pp_sort_small_kernel(const spsp_hit *__restrict__ hits, const unsigned long long *__restrict__ hit_count, uint64_t hits_cap,
                     const uint64_t *__restrict__ rec_begin, const uint64_t *__restrict__ rec_end, uint64_t n_rec, int k, int m,
                     const uint32_t *__restrict__ key_in, int key_bits, uint32_t *__restrict__ key_out,
                     uint32_t *__restrict__ val_out, Counters *cnt)
{
    extern __shared__ __align__(16) uint8_t ss_smem[];
    uint32_t *s_key = reinterpret_cast<uint32_t *>(ss_smem);                 // [SS_CAP] keys in sorted-by-bin order
    uint32_t *s_off = s_key + SS_CAP;                                        // [SS_BINS + 1] bin cursors / ends
    uint16_t *s_idx = reinterpret_cast<uint16_t *>(s_off + SS_BINS + 32);    // [SS_CAP] original index
    __shared__ uint32_t s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t n64 = HITS ? (uint64_t)*hit_count : cnt->n_pieces;
    if (HITS && n64 > hits_cap) n64 = 0;              // the scan dropped hits: the host retries with a larger buffer
    if (n64 > (uint64_t)SS_CAP) {
        if (threadIdx.x == 0) atomicOr(&cnt->overflow, (unsigned)OVF_SORT);
        n64 = 0;
    }
    const uint32_t n = (uint32_t)n64;
    const int shift = key_bits > SS_BIN_BITS ? key_bits - SS_BIN_BITS : 0;
    for (uint32_t b = threadIdx.x; b <= SS_BINS; b += SS_THREADS) s_off[b] = 0;
    // One CTA means one memory latency per dependent load: every loop below works on SS_MLP keys per thread at a
    // time, with their loads issued together, and the record look-up of a hit starts in a sampled copy of
    // rec_begin in shared memory (complete when the batch has at most 256 records: whole genomes).
    __shared__ uint64_t s_coarse[256];
    uint64_t c_stride = 1;
    uint32_t c_n = 0;
    if (HITS && n_rec) {
        c_stride = (n_rec + 255) / 256;
        c_n = (uint32_t)((n_rec + c_stride - 1) / c_stride);
        for (uint32_t j = threadIdx.x; j < c_n; j += SS_THREADS) s_coarse[j] = rec_begin[(uint64_t)j * c_stride];
    }
    __syncthreads();
    // pass 1: keys (kept in key_out, unsorted, for pass 2) and the histogram
    for (uint32_t base = 0; base < n; base += SS_THREADS * SS_MLP) {
        uint32_t key[SS_MLP];
        if (HITS) {
            uint64_t pos[SS_MLP], rb[SS_MLP], re[SS_MLP];
            long long r[SS_MLP];
#pragma unroll
            for (int j = 0; j < SS_MLP; j++) {
                const uint32_t i = base + j * SS_THREADS + threadIdx.x;
                pos[j] = i < n ? hits[i].pos : ~0ULL;
            }
#pragma unroll
            for (int j = 0; j < SS_MLP; j++) {
                // last record whose begin is <= pos: coarse step in shared memory, the rest (if any) in global memory
                uint32_t lo = 0, hi = c_n;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (s_coarse[mid] <= pos[j]) lo = mid + 1; else hi = mid;
                }
                long long rr = -1;
                if (lo) {
                    uint64_t a = (uint64_t)(lo - 1) * c_stride, b = min(a + c_stride, n_rec);   // begin[a] <= pos
                    a++;
                    while (a < b) {
                        const uint64_t mid = (a + b) >> 1;
                        if (rec_begin[mid] <= pos[j]) a = mid + 1; else b = mid;
                    }
                    rr = (long long)a - 1;
                }
                r[j] = rr;
            }
#pragma unroll
            for (int j = 0; j < SS_MLP; j++) {
                const bool ok = r[j] >= 0 && pos[j] != ~0ULL;
                rb[j] = ok ? rec_begin[r[j]] : 0;
                re[j] = ok ? rec_end[r[j]] : 0;
            }
#pragma unroll
            for (int j = 0; j < SS_MLP; j++) {
                const uint32_t i = base + j * SS_THREADS + threadIdx.x;
                key[j] = KEY_INVALID;
                if (i < n) {
                    if (r[j] >= 0 && pos[j] + (uint64_t)m <= re[j] && re[j] - rb[j] >= (uint64_t)k) key[j] = (uint32_t)pos[j];
                    key_out[i] = key[j];
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < SS_MLP; j++) {
                const uint32_t i = base + j * SS_THREADS + threadIdx.x;
                key[j] = i < n ? key_in[i] : KEY_INVALID;
            }
        }
#pragma unroll
        for (int j = 0; j < SS_MLP; j++)
            if (key[j] != KEY_INVALID) atomicAdd(&s_off[min(key[j] >> shift, (uint32_t)SS_BINS - 1)], 1u);
    }
    __syncthreads();
    // exclusive scan of the histogram (8 bins per thread)
    uint32_t total;
    {
        uint32_t c[SS_BINS / SS_THREADS], sum = 0;
#pragma unroll
        for (int j = 0; j < SS_BINS / SS_THREADS; j++) {
            c[j] = s_off[threadIdx.x * (SS_BINS / SS_THREADS) + j];
            sum += c[j];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t wb = 0, tot = 0;
        for (int w = 0; w < 32; w++) {
            const uint32_t t = s_warp[w];
            if (w < warp) wb += t;
            tot += t;
        }
        total = tot;
        uint32_t run = wb + inc - sum;
#pragma unroll
        for (int j = 0; j < SS_BINS / SS_THREADS; j++) {
            s_off[threadIdx.x * (SS_BINS / SS_THREADS) + j] = run;
            run += c[j];
        }
    }
    __syncthreads();
    // pass 2: scatter into the bins (cursor = running end of the bin)
    for (uint32_t base = 0; base < n; base += SS_THREADS * SS_MLP) {
        uint32_t key[SS_MLP];
#pragma unroll
        for (int j = 0; j < SS_MLP; j++) {
            const uint32_t i = base + j * SS_THREADS + threadIdx.x;
            key[j] = i < n ? (HITS ? key_out[i] : key_in[i]) : KEY_INVALID;
        }
#pragma unroll
        for (int j = 0; j < SS_MLP; j++) {
            if (key[j] != KEY_INVALID) {
                const uint32_t d = atomicAdd(&s_off[min(key[j] >> shift, (uint32_t)SS_BINS - 1)], 1u);
                s_key[d] = key[j];
                s_idx[d] = (uint16_t)(base + j * SS_THREADS + threadIdx.x);
            }
        }
    }
    __syncthreads();
    // order inside the bins (s_off[b] is now the END of bin b): every key counts the keys of its bin that sort before
    // it -- (key, original index), which makes the sort stable -- and goes straight to its final place
    for (uint32_t base = 0; base < total; base += SS_THREADS * SS_MLP) {
        uint32_t dst[SS_MLP], kxs[SS_MLP], val[SS_MLP];
        uint16_t ixs[SS_MLP];
#pragma unroll
        for (int j = 0; j < SS_MLP; j++) {
            const uint32_t r = base + j * SS_THREADS + threadIdx.x;
            dst[j] = 0xFFFFFFFFu; kxs[j] = 0; ixs[j] = 0;
            if (r < total) {
                const uint32_t kx = s_key[r];
                const uint16_t ix = s_idx[r];
                const uint32_t b = min(kx >> shift, (uint32_t)SS_BINS - 1);
                const uint32_t lo = b ? s_off[b - 1] : 0, hi = s_off[b];
                uint32_t rank = 0;
                for (uint32_t q = lo; q < hi; q++) rank += (s_key[q] < kx || (s_key[q] == kx && s_idx[q] < ix)) ? 1u : 0u;
                dst[j] = lo + rank; kxs[j] = kx; ixs[j] = ix;
            }
        }
        if (HITS) {
            uint32_t cn[SS_MLP], rv[SS_MLP];
#pragma unroll
            for (int j = 0; j < SS_MLP; j++) {
                cn[j] = dst[j] != 0xFFFFFFFFu ? hits[ixs[j]].canon : 0;
                rv[j] = dst[j] != 0xFFFFFFFFu ? hits[ixs[j]].rev : 0;
            }
#pragma unroll
            for (int j = 0; j < SS_MLP; j++) val[j] = (cn[j] << 1) | (rv[j] & 1u);
        } else {
#pragma unroll
            for (int j = 0; j < SS_MLP; j++) val[j] = ixs[j];
        }
#pragma unroll
        for (int j = 0; j < SS_MLP; j++)
            if (dst[j] != 0xFFFFFFFFu) { key_out[dst[j]] = kxs[j]; val_out[dst[j]] = val[j]; }
    }
    if (HITS && threadIdx.x == 0) cnt->n_valid = total;
}

template <bool HITS>
static cudaError_t launch_sort_small(const spsp_hit *hits, const unsigned long long *hit_count, uint64_t hits_cap,
                                     const uint64_t *rec_begin, const uint64_t *rec_end, uint64_t n_rec, int k, int m,
                                     const uint32_t *key_in, int key_bits, uint32_t *key_out, uint32_t *val_out, Counters *cnt,
                                     cudaStream_t st)
{
    static PerDeviceOnce once;
    cudaError_t e = once.run([&] {
        return cudaFuncSetAttribute(pp_sort_small_kernel<HITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SS_SMEM);
    });
    if (e != cudaSuccess) return e;
    pp_sort_small_kernel<HITS><<<1, SS_THREADS, SS_SMEM, st>>>(hits, hit_count, hits_cap, rec_begin, rec_end, n_rec, k, m, key_in,
                                                              key_bits, key_out, val_out, cnt);
    return cudaGetLastError();
}

// Large path: classify into 32-bit sort keys for cub::DeviceRadixSort over the whole capacity.
__global__ void pp_classify_kernel(const spsp_hit *__restrict__ hits, const unsigned long long *__restrict__ hit_count,
                                   uint64_t hits_cap, const uint64_t *__restrict__ rec_begin, const uint64_t *__restrict__ rec_end,
                                   uint64_t n_rec, int k, int m, uint32_t *__restrict__ key, uint32_t *__restrict__ val,
                                   Counters *cnt)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hits_cap) return;
    uint64_t n = *hit_count;
    if (n > hits_cap) n = 0;
    uint32_t kk = KEY_INVALID, vv = 0;
    if (i < n) {
        const uint64_t pos = hits[i].pos;
        vv = (hits[i].canon << 1) | (hits[i].rev & 1u);
        if (n_rec) {
            const long long r = find_rec(rec_begin, n_rec, pos);
            if (r >= 0) {
                const uint64_t b = rec_begin[r], e = rec_end[r];
                if (pos + (uint64_t)m <= e && e - b >= (uint64_t)k) kk = (uint32_t)pos;
            }
        }
    }
    key[i] = kk;
    val[i] = vv;
    const unsigned ok = __popc(__ballot_sync(0xffffffffu, kk != KEY_INVALID));
    if ((threadIdx.x & 31) == 0 && ok) atomicAdd(&cnt->n_valid, (unsigned long long)ok);
}

// --------------------------------------------------------------- R1/R2 replay

struct Track {
    bool valid;
    uint32_t canon;
    uint64_t hash, posmin;
    bool rev;
};

// regular_minimizer_pos restricted to hits (SubSampler.cpp:81-169), window of
// k-mer c = hits [lo, hi); keeps the reference's position quirks (:88-93, :149-164).
__device__ Track pp_rescan(const uint32_t *__restrict__ key, const uint32_t *__restrict__ val, uint64_t rb, uint32_t lo,
                           uint32_t hi, uint64_t c, uint64_t d)
{
    Track t{false, 0, 0, 0, false};
    uint64_t position = 0;
    for (uint32_t i = hi; i-- > lo;) {
        const uint64_t p = key[i] - rb;
        const uint32_t cn = val[i] >> 1;
        const uint64_t h = xxh64_8((uint64_t)cn);
        const bool rv = val[i] & 1u;
        const uint64_t j = c + d - p;
        if (j == 0) {
            t.valid = true; t.canon = cn; t.hash = h; t.rev = rv;
            position = rv ? 0 : d;
        } else if (!t.valid || t.hash > h) {
            t.valid = true; t.canon = cn; t.hash = h; t.rev = rv;
            position = d - j;
        } else if (cn == t.canon && rv == t.rev) {
            if (t.rev && position > j) position = j;
            if (!t.rev && position > d - j) position = d - j;
        }
    }
    t.posmin = c + position;
    return t;
}

struct ReplayArgs {
    const uint32_t *key, *val;                   // valid hits sorted by position
    const uint64_t *rec_begin, *rec_end;
    const uint32_t *rec_input;
    uint64_t n_rec;
    int k, m, input_shift;
    uint32_t *cl_np, *cl_nk;                     // [hit] pieces / k-mers of the cluster that starts at the hit (0 otherwise)
    uint32_t *cta_np;                            // [CTA] pieces per CTA
    uint4 *pieces;                               // {first k-mer (global base), minimizer, input << 1 | rev, k-mers}
    uint32_t *pkey32;                            // sort key of every piece: bucket = input << 2m | minimizer ...
    uint64_t *pkey64;                            // ... as 64 bits when it does not fit 32
    uint64_t pieces_cap;
    unsigned long long *in_sel;                  // [input] selected k-mer occurrences (header field 3)
    Counters *cnt;
};

// Sparse replay of SubSampler.cpp:352-454 over one cluster of hits: hits closer than d + 1 = k - m + 1 apart
// (same record) can share a k-mer window and are replayed together, everything else is independent.
// One thread per hit; the thread of a cluster's first hit does the cluster.
// WRITE = false: count pieces and k-mers; WRITE = true: store the pieces at their final (genome-order) index.
template <bool WRITE>
__global__ void __launch_bounds__(RP_THREADS) pp_replay_kernel(ReplayArgs a)
{
    __shared__ uint32_t s_warp[RP_THREADS / 32];
    __shared__ uint32_t s_base;
    const uint64_t n_valid = a.cnt->n_valid;
    const uint64_t i0l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.k, m = a.m;
    const uint64_t d = (uint64_t)(k - m);
    const uint32_t *key = a.key, *val = a.val;

    bool start = false;
    uint32_t r = 0;
    uint64_t rb = 0, re = 0;
    if (i0l < n_valid) {
        const uint64_t pos = key[i0l];
        r = (uint32_t)find_rec(a.rec_begin, a.n_rec, pos);
        rb = a.rec_begin[r]; re = a.rec_end[r];
        // gap == d + 1 still couples two hits: the k-mer after the older hit's last window already sees the
        // newer one, and the rescan that fetches it applies the reference's position quirks
        start = i0l == 0 || (uint64_t)key[i0l - 1] < rb || pos - key[i0l - 1] > d + 1;
    }
    uint32_t my_np = 0;
    uint64_t out = 0;
    if (WRITE) {
        // pieces before this CTA (sum of the earlier CTAs' counts), then an exclusive scan inside the CTA
        uint32_t part = 0;
        for (uint32_t b = threadIdx.x; b < blockIdx.x; b += blockDim.x) part += a.cta_np[b];
        part = __reduce_add_sync(0xffffffffu, part);
        if (lane == 0) s_warp[warp] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < RP_THREADS / 32; w++) t += s_warp[w];
            s_base = t;
        }
        __syncthreads();
        const uint32_t base = s_base;
        __syncthreads();
        my_np = (i0l < n_valid) ? a.cl_np[i0l] : 0;
        uint32_t inc = my_np;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < warp; w++) wbase += s_warp[w];
        out = (uint64_t)base + wbase + (inc - my_np);
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x == RP_THREADS - 1) {
            const unsigned long long total = (unsigned long long)base + wbase + inc;
            a.cnt->n_pieces = total;
            if (total > a.pieces_cap) atomicOr(&a.cnt->overflow, (unsigned)OVF_PIECES);
        }
    }

    uint32_t np = 0, nk = 0;
    if (start && (!WRITE || my_np)) {
        const uint32_t i0 = (uint32_t)i0l;
        uint32_t i1 = i0 + 1;                                 // cluster = hits [i0, i1)
        while (i1 < n_valid && (uint64_t)key[i1] + (uint64_t)m <= re && key[i1] - key[i1 - 1] <= d + 1) i1++;
        const uint64_t n = re - rb, K = n - k + 1;
        const uint32_t input = a.rec_input[r];
        auto emit = [&](uint64_t first, uint64_t last, uint32_t mn, bool rev) {
            if (WRITE) {
                if (out < a.pieces_cap) {
                    a.pieces[out] = make_uint4((uint32_t)(rb + first), mn, (input << 1) | (rev ? 1u : 0u), (uint32_t)(last - first + 1));
                    const uint64_t bucket = ((uint64_t)input << a.input_shift) | mn;
                    if (a.pkey64) a.pkey64[out] = bucket; else a.pkey32[out] = (uint32_t)bucket;
                }
                out++;
            }
            np++;
            nk += (uint32_t)(last - first + 1);
        };
        const uint64_t p_first = key[i0] - rb, p_last = key[i1 - 1] - rb;
        uint32_t lo = i0, hi = i0;                      // hits with pos in [c, c+d] are [lo, hi)
        auto window = [&](uint64_t c) {
            while (hi < i1 && key[hi] - rb <= c + d) hi++;
            while (lo < hi && key[lo] - rb < c) lo++;
        };
        Track cur{false, 0, 0, 0, false};
        bool old_valid = false, old_rev = false, is_rev = false;
        uint32_t old_min = 0;
        uint64_t last = 0, c;
        if (p_first <= d) {
            // the record's first k-mer already sees a hit: initial rescan (:359-365)
            window(0);
            cur = pp_rescan(key, val, rb, lo, hi, 0, d);
            old_valid = cur.valid; old_rev = cur.rev; is_rev = cur.rev; old_min = cur.canon;
            last = 0;
            c = 1;
        } else {
            c = p_first - d;                          // the first hit enters on the right
        }
        const uint64_t c_end = p_last < K - 1 ? p_last : K - 1;   // after p_last the window holds no hit of this cluster
        for (; c <= c_end; c++) {
            window(c);
            const uint64_t p = c + d;
            bool dump = false;
            const bool ent = hi > lo && key[hi - 1] - rb == p;
            uint64_t h_ent = 0;
            if (ent) h_ent = xxh64_8((uint64_t)(val[hi - 1] >> 1));
            if (ent && (!cur.valid || h_ent < cur.hash)) {                       // :374-388
                cur.valid = true; cur.canon = val[hi - 1] >> 1; cur.hash = h_ent; cur.posmin = p;
                cur.rev = val[hi - 1] & 1u; is_rev = cur.rev;
            } else if (cur.valid && c - 1 >= cur.posmin) {                       // :391-398
                cur = pp_rescan(key, val, rb, lo, hi, c, d);
                if (cur.valid) is_rev = cur.rev;
                dump = true;
            }
            const bool changed = (old_valid != cur.valid) || (cur.valid && old_min != cur.canon);
            if (changed || dump) {                                               // :401-435
                if (old_valid) emit(last, c - 1, old_min, old_rev);
                last = c;
                old_valid = cur.valid; old_min = cur.canon; old_rev = is_rev;
            }
        }
        // leaving the cluster: either the record ends here (:441-450), or at c_end + 1 the window holds no hit
        // any more (the next hit is more than d + 1 away), the tracked minimizer is outdated (posmin <= p_last)
        // and the rescan finds a non-selected one: the piece ends at c_end in both cases.
        if (old_valid) emit(last, c_end, old_min, old_rev);
        if (!WRITE && nk) atomicAdd(a.in_sel + input, (unsigned long long)nk);   // header field 3: selected k-mer occurrences
    }
    if (!WRITE) {
        if (i0l < n_valid) { a.cl_np[i0l] = np; a.cl_nk[i0l] = nk; }
        const uint32_t wnp = __reduce_add_sync(0xffffffffu, np), wnk = __reduce_add_sync(0xffffffffu, nk);
        if (lane == 0) s_warp[warp] = wnp;
        if (lane == 0 && wnk) atomicAdd(&a.cnt->n_entries, (unsigned long long)wnk);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < RP_THREADS / 32; w++) t += s_warp[w];
            a.cta_np[blockIdx.x] = t;
        }
    }
}

// ------------------------------------------------------------ B bucket groups
//
// After S2 the pieces of one bucket (input, minimizer) are adjacent, in the reference's insertion order.
// Group g owns the buckets whose first piece has a sorted index in [g * pp, (g + 1) * pp).  Its CTA stages
// those pieces' k-mers and restates, per bucket:
//   handle_superkmer (SubSampler.cpp:243-302): oriented k-mer keys left to right, leftmost minimizer
//     position, `minimizer_map[min][kmer]` as an open-addressing table keyed (bucket, k-mer) whose slot keeps
//     the FIRST entry of its key (the dense map iterates in insertion order) and the number of entries (the
//     uint8 count, SubSampler.h:24);
//   find_first_kmer / reconstruct_superkmer / find_next (:512-620): all eight neighbour look-ups of every
//     unique k-mer up front (probe order A,T,C,G, :568), then one thread per bucket follows indices and
//     `seen` flags and records the visit order;
//   the writer loop (:459-504): minimizer text, u32 size, packed maximal super-k-mers, the others as
//     "prefix\nsuffix\n" text, "\n\n";
//   the comparator's view (Comparator.cpp:97-264): distinct canonical k-mers of the bucket.
// Groups that do not fit the shared-memory capacities (a bucket with thousands of occurrences: low-complexity
// sequence, dense sampling) run the same code on arrays carved from a global pool.

struct BucketArgs {
    const uint32_t *packed;
    const uint4 *pieces;
    const uint32_t *order;                 // sorted piece -> piece index
    const uint32_t *skey32;                // sorted bucket keys ...
    const uint64_t *skey64;                // ... or their 64-bit form
    uint32_t pp;                           // sorted pieces per group
    uint64_t pieces_cap;
    int k, m, input_shift;
    unsigned abundance;
    // A group writes its bytes and elements to ranges it draws from two bump allocators (any order) and leaves
    // {where, how much}; pp_offsets_kernel turns the sizes into final offsets (groups are in bucket order) and
    // pp_gather_kernel moves every group's output to its final place.  (Round 1/2a obtained the offsets inside
    // the group kernel by decoupled look-back: every group then waited, holding its SM slot, for all running
    // predecessors to reach the same point -- 40 k of its 134 k cycles.)
    unsigned long long *g_tb, *g_nb, *g_te, *g_ne;   // [groups] temp byte offset / bytes, temp element offset / elements
    // Small batches (the one-CTA sort path, <= 32 K hits) keep the decoupled look-back: there the two extra launches
    // of the two-pass form cost more than the waiting (64 x 5 Mbp batches: value 1 291 vs 1 417 Gbp/s on one box);
    // then lb_* are the look-back states and tmp_* point at the final arrays.  Large batches: 3.70 -> 3.50 ms per
    // 1 Gbp at s = 100, 2.3 -> 1.0 ms per 1 Gbp read set with the two-pass form.
    unsigned long long *lb_bytes, *lb_elems;
    uint8_t *tmp_body;
    uint64_t body_cap;
    uint32_t *tmp_min;
    uint64_t *tmp_klo, *tmp_khi;
    uint64_t elems_cap;
    unsigned long long *in_first_byte, *in_first_elem;   // [input] (group + 1) << 32 | offset inside the group of the input's first bucket, 0 = none
    uint8_t *big_pool;
    uint64_t big_cap;
    Counters *cnt;
    long long *dbg;                        // SPSP_PP_DEBUG: [group][16] phase time stamps (clock64) + sizes
};
#define PP_PHASE(i) do { if (a.dbg && threadIdx.x == 0) a.dbg[(size_t)grp * 16 + (i)] = clock64(); } while (0)

// Arrays of one group: shared memory (idx_t = uint16_t) or the global pool (uint32_t).
template <class IDX>
struct GroupMem {
    uint4 *pc;                 // [np] piece records in sorted order
    uint32_t *p_eoff;          // [np + 1] first entry of every piece
    IDX *p_bl;                 // [np] bucket (group-local) of every piece
    uint64_t *e_lo, *e_hi;     // [E] oriented k-mer of every entry (e_hi only when k > 32)
    uint8_t *e_pm;             // [E] leftmost minimizer position, 255 = none
    IDX *e_bl;                 // [E] bucket
    uint32_t *e_slot;          // [E] slot of the entry's key; later: unique id / scan scratch
    uint32_t *h_first;         // [S] smallest entry with the slot's key, H_EMPTY if free
    uint32_t *h_count;         // [S] entries with the slot's key
    IDX *h_uniq;               // [S] unique id of the slot's key
    uint64_t *u_lo, *u_hi;     // [E] unique k-mers grouped by bucket, insertion order
    uint8_t *u_cnt, *u_pm, *seen;   // [E], seen has E + 1 (sentinel "no neighbour")
    IDX *u_bl;
    IDX *adj;                  // [E * 8] neighbour unique id: left A,T,C,G then right A,T,C,G; U = none
    IDX *visit;                // [E] unique id | role << (bits - 2); all-ones = end of the bucket's list
    uint32_t *b_us;            // [NB + 1] first unique of every bucket
    uint32_t *b_hp;            // [NB] first piece
    uint32_t *b_bytes, *b_nmax, *b_eloff;   // [NB + 1]
    uint32_t S;                // slots (power of two)
};

template <class IDX> struct IdxTraits;
template <> struct IdxTraits<uint16_t> {
    static constexpr uint32_t ROLE_SHIFT = 14, END = 0xFFFFu, MASK = 0x3FFFu;
};
template <> struct IdxTraits<uint32_t> {
    static constexpr uint32_t ROLE_SHIFT = 30, END = 0xFFFFFFFFu, MASK = 0x3FFFFFFFu;
};

__device__ __forceinline__ size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Carves the arrays out of `base`; returns the bytes used.
template <class IDX>
__device__ __host__ inline size_t group_carve(GroupMem<IDX> &g, uint8_t *base, uint32_t np, uint32_t E, uint32_t S, bool hi128)
{
    size_t o = 0;
    auto take = [&](size_t bytes) { uint8_t *p = base ? base + o : nullptr; o = (o + bytes + 15) & ~(size_t)15; return p; };
    g.pc = reinterpret_cast<uint4 *>(take((size_t)np * 16));
    g.p_eoff = reinterpret_cast<uint32_t *>(take((size_t)(np + 1) * 4));
    g.p_bl = reinterpret_cast<IDX *>(take((size_t)np * sizeof(IDX)));
    g.e_lo = reinterpret_cast<uint64_t *>(take((size_t)E * 8));
    g.e_hi = hi128 ? reinterpret_cast<uint64_t *>(take((size_t)E * 8)) : nullptr;
    g.e_pm = take(E);
    g.e_bl = reinterpret_cast<IDX *>(take((size_t)E * sizeof(IDX)));
    g.e_slot = reinterpret_cast<uint32_t *>(take((size_t)E * 4));
    g.h_first = reinterpret_cast<uint32_t *>(take((size_t)S * 4));
    g.h_count = reinterpret_cast<uint32_t *>(take((size_t)S * 4));
    g.h_uniq = reinterpret_cast<IDX *>(take((size_t)S * sizeof(IDX)));
    g.u_lo = reinterpret_cast<uint64_t *>(take((size_t)E * 8));
    g.u_hi = hi128 ? reinterpret_cast<uint64_t *>(take((size_t)E * 8)) : nullptr;
    g.u_cnt = take(E);
    g.u_pm = take(E);
    g.seen = take((size_t)E + 1);
    g.u_bl = reinterpret_cast<IDX *>(take((size_t)E * sizeof(IDX)));
    g.adj = reinterpret_cast<IDX *>(take((size_t)E * 8 * sizeof(IDX)));
    g.visit = reinterpret_cast<IDX *>(take((size_t)E * sizeof(IDX)));
    g.b_us = reinterpret_cast<uint32_t *>(take((size_t)(np + 1) * 4));
    g.b_hp = reinterpret_cast<uint32_t *>(take((size_t)np * 4));
    g.b_bytes = reinterpret_cast<uint32_t *>(take((size_t)(np + 1) * 4));
    g.b_nmax = reinterpret_cast<uint32_t *>(take((size_t)(np + 1) * 4));
    g.b_eloff = reinterpret_cast<uint32_t *>(take((size_t)(np + 1) * 4));
    g.S = S;
    return o;
}

size_t bucket_smem_bytes(bool hi128)
{
    GroupMem<uint16_t> g;
    return group_carve<uint16_t>(g, nullptr, BK_PMAX, BK_ECAP, BK_SLOTS, hi128) + 16;
}

// In-place exclusive scan of data[0, n) by the whole CTA; data[n] receives the total (so data needs n + 1
// entries).  Ends with a barrier.
__device__ void block_excl_scan(uint32_t *data, uint32_t n, uint32_t *s_warp /* [BK_WARPS + 1] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < n; b0 += BK_THREADS) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < n ? data[i] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t wb = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < BK_WARPS; w++) {
            const uint32_t t = s_warp[w];
            if (w < warp) wb += t;
            tot += t;
        }
        if (i < n) data[i] = carry + wb + inc - v;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) data[n] = carry;
    __syncthreads();
}

__device__ __forceinline__ uint32_t grp_hash(uint32_t bl, uint64_t lo, uint64_t hi)
{
    uint64_t x = lo ^ (hi * 0xC2B2AE3D27D4EB4FULL) ^ ((uint64_t)bl * 0x9E3779B97F4A7C15ULL);
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ULL;
    x ^= x >> 29;
    return (uint32_t)x;
}
// Slot of (bucket, key) in the group's table, or ~0.
template <class IDX>
__device__ __forceinline__ uint32_t grp_find(const GroupMem<IDX> &g, uint32_t bl, uint64_t lo, uint64_t hi)
{
    uint32_t s = grp_hash(bl, lo, hi) & (g.S - 1);
    for (;;) {
        const uint32_t e = g.h_first[s];
        if (e == H_EMPTY) return ~0u;
        if (g.e_lo[e] == lo && (uint32_t)g.e_bl[e] == bl && (!g.e_hi || g.e_hi[e] == hi)) return s;
        s = (s + 1) & (g.S - 1);
    }
}
template <bool HI>
__device__ __forceinline__ K128 kmer_neighbour(const K128 &cur, bool left, int t, int k)
{
    const uint64_t o = (0x3120u >> (4 * t)) & 3u;                          // A, T, C, G (SubSampler.cpp:568)
    if (!HI) {                                                             // k <= 32: one word
        if (left) return K128{(cur.lo >> 2) | (o << (2 * k - 2)), 0};
        return K128{((cur.lo << 2) | o) & (~0ULL >> (64 - 2 * k)), 0};
    }
    K128 nx;
    if (left) {
        nx = k_shr(cur, 2);
        K128 top = k_shl(K128{o, 0}, 2 * k - 2);
        nx.lo |= top.lo; nx.hi |= top.hi;
    } else {
        const K128 kmask = k_shr(K128{~0ULL, ~0ULL}, 128 - 2 * k);
        nx = k_shl(cur, 2);
        nx.lo |= o;
        nx.lo &= kmask.lo; nx.hi &= kmask.hi;
    }
    return nx;
}

// Decoupled look-back over the groups' aggregates: returns the sum of the aggregates of groups < g and
// leaves this group's inclusive prefix for its successors.  state = flag << 62 | value (flag 1 = aggregate,
// 2 = inclusive prefix); called by one full warp.
__device__ unsigned long long lookback(unsigned long long *state, uint32_t g, unsigned long long agg)
{
    const int lane = threadIdx.x & 31;
    volatile unsigned long long *vs = state;
    if (lane == 0) vs[g] = ((g == 0 ? 2ULL : 1ULL) << 62) | agg;
    unsigned long long excl = 0;
    if (g == 0) return 0;
    long long j = (long long)g - 1;
    for (;;) {
        const long long idx = j - lane;
        unsigned long long v = 2ULL << 62;                       // before group 0: inclusive prefix 0
        if (idx >= 0) {
            do { v = vs[idx]; } while ((v >> 62) == 0);
        }
        const unsigned inc_mask = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const int first_inc = inc_mask ? __ffs(inc_mask) - 1 : 32;
        unsigned long long part = (lane <= first_inc) ? (v & ((1ULL << 62) - 1)) : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        excl += part;
        if (inc_mask) break;
        j -= 32;
    }
    if (lane == 0) vs[g] = (2ULL << 62) | (excl + agg);
    return excl;
}

// Leftmost occurrence of the minimizer in codes sk[0, len), -1 if none (one thread).
__device__ int find_minimizer_codes(const uint8_t *sk, int len, int m, uint32_t minimizer)
{
    const uint32_t mmask = (1u << (2 * m)) - 1u;
    uint32_t w = 0;
    for (int t = 0; t < len; t++) {
        w = ((w << 2) | sk[t]) & mmask;
        if (t >= m - 1 && w == minimizer) return t - (m - 1);
    }
    return -1;
}

template <class IDX>
struct GroupCtx {
    uint32_t np, E, U, NB;
    uint32_t first;            // sorted index of the group's first piece
};

// Everything a group does once its piece records are in g.pc[0, np).  Called by the whole CTA.
template <class IDX, bool HI>
__device__ void process_group(const BucketArgs &a, GroupMem<IDX> &g, uint32_t grp, uint32_t first, uint32_t np,
                              uint32_t *s_warp, uint8_t *s_sk /* [BK_WARPS][RC_SK] */, uint32_t *s_misc /* [8] */)
{
    using T = IdxTraits<IDX>;
    if (sizeof(IDX) == 2) {
        // the 16-bit instantiation works on shared memory only: lets the compiler use LDS/STS/ATOMS instead of
        // generic accesses
#define PP_SH(p) __builtin_assume(__isShared(p))
        PP_SH(g.pc); PP_SH(g.p_eoff); PP_SH(g.p_bl); PP_SH(g.e_lo); PP_SH(g.e_pm); PP_SH(g.e_bl); PP_SH(g.e_slot);
        PP_SH(g.h_first); PP_SH(g.h_count); PP_SH(g.h_uniq); PP_SH(g.u_lo); PP_SH(g.u_cnt); PP_SH(g.u_pm); PP_SH(g.seen);
        PP_SH(g.u_bl); PP_SH(g.adj); PP_SH(g.visit); PP_SH(g.b_us); PP_SH(g.b_hp); PP_SH(g.b_bytes); PP_SH(g.b_nmax);
        PP_SH(g.b_eloff);
        if (HI) { PP_SH(g.e_hi); PP_SH(g.u_hi); }
#undef PP_SH
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.k, m = a.m, d = k - m, full = 2 * k - m;
    const uint32_t mmask = (1u << (2 * m)) - 1u;
    auto bucket_key = [&](uint32_t sorted_idx) -> uint64_t {
        return a.skey64 ? a.skey64[sorted_idx] : (uint64_t)a.skey32[sorted_idx];
    };
    PP_PHASE(0);
    // ---- pieces: entry offsets, bucket heads
    for (uint32_t p = threadIdx.x; p < np; p += BK_THREADS) {
        g.p_eoff[p] = g.pc[p].w;
        g.b_us[p] = (p == 0 || bucket_key(first + p) != bucket_key(first + p - 1)) ? 1u : 0u;   // head flags
    }
    __syncthreads();
    block_excl_scan(g.p_eoff, np, s_warp);
    block_excl_scan(g.b_us, np, s_warp);               // b_us[p] = heads before p; a head p starts bucket b_us[p]
    const uint32_t E = g.p_eoff[np], NB = g.b_us[np];
    for (uint32_t p = threadIdx.x; p < np; p += BK_THREADS) {
        const bool head = (p + 1 <= np) && (g.b_us[p + 1] != g.b_us[p]);
        const uint32_t b = g.b_us[p] - (head ? 0 : 1);
        g.p_bl[p] = (IDX)b;
        if (head) g.b_hp[b] = p;
    }
    for (uint32_t s = threadIdx.x; s < g.S; s += BK_THREADS) { g.h_first[s] = H_EMPTY; g.h_count[s] = 0; }
    __syncthreads();
    PP_PHASE(1);
    // ---- entries (handle_superkmer, :243-302): entry e is the (e - first entry of its piece)-th k-mer of the
    // oriented piece; key as oriented, pos_min = leftmost occurrence of the minimizer in it
    for (uint32_t e = threadIdx.x; e < E; e += BK_THREADS) {
        uint32_t lo = 0, hi = np;                        // piece = last one whose first entry is <= e
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (g.p_eoff[mid] <= e) lo = mid + 1; else hi = mid;
        }
        const uint32_t p = lo - 1;
        const uint4 pc = g.pc[p];
        const uint32_t j = e - g.p_eoff[p], nk = pc.w, mn = pc.y;
        const bool rev = pc.z & 1u;
        const uint64_t pos = (uint64_t)pc.x + (rev ? (nk - 1 - j) : j);
        K128 key = HI ? kmer_at(a.packed, pos, k) : K128{kmer_at64(a.packed, pos, k), 0};
        if (rev) key = k_rc_t<HI>(key, k);
        unsigned pm = 255;
        if (!HI) {
            for (int q = 0; q <= d; q++)
                if (((uint32_t)(key.lo >> (2 * (d - q))) & mmask) == mn) { pm = (unsigned)q; break; }
        } else {
            for (int q = 0; q <= d; q++)
                if (((uint32_t)k_shr(key, 2 * (d - q)).lo & mmask) == mn) { pm = (unsigned)q; break; }
        }
        g.e_lo[e] = key.lo;
        if (HI) g.e_hi[e] = key.hi;
        g.e_pm[e] = (uint8_t)pm;
        g.e_bl[e] = g.p_bl[p];
    }
    __syncthreads();
    PP_PHASE(2);
    // ---- minimizer_map[min][kmer]: first occurrence + count per (bucket, k-mer)
    for (uint32_t e = threadIdx.x; e < E; e += BK_THREADS) {
        const uint64_t lo = g.e_lo[e], hi = HI ? g.e_hi[e] : 0;
        const uint32_t bl = g.e_bl[e];
        uint32_t s = grp_hash(bl, lo, hi) & (g.S - 1);
        for (;;) {
            uint32_t o = *reinterpret_cast<volatile uint32_t *>(g.h_first + s);
            if (o == H_EMPTY) {
                o = atomicCAS(g.h_first + s, H_EMPTY, e);
                if (o == H_EMPTY) break;                           // claimed a free slot
            }
            // the slot belongs to the key of entry o (entries are immutable, so that key can be read)
            if (g.e_lo[o] == lo && (uint32_t)g.e_bl[o] == bl && (!HI || g.e_hi[o] == hi)) {
                atomicMin(g.h_first + s, e);
                break;
            }
            s = (s + 1) & (g.S - 1);
        }
        atomicAdd(g.h_count + s, 1u);
        g.e_slot[e] = s;
    }
    __syncthreads();
    PP_PHASE(3);
    // ---- unique k-mers in insertion order: entries that are the first of their key, compacted in entry order
    // (entries are ordered by bucket, so the unique list is grouped by bucket as well)
    // e_slot[e] := is-first flag, then (unique id << 1 | flag) by a scan in place (the slot is looked up again from
    // the key afterwards)
    for (uint32_t e = threadIdx.x; e < E; e += BK_THREADS) {
        const uint32_t s = g.e_slot[e];
        g.e_slot[e] = (g.h_first[s] == e) ? 1u : 0u;
    }
    __syncthreads();
    {
        // exclusive scan of e_slot[0, E) with the total in s_misc[0]
        const int ln = lane, wp = warp;
        uint32_t carry = 0;
        for (uint32_t b0 = 0; b0 < E; b0 += BK_THREADS) {
            const uint32_t i = b0 + threadIdx.x;
            const uint32_t v = i < E ? g.e_slot[i] : 0;
            uint32_t inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (ln >= o) inc += t;
            }
            if (ln == 31) s_warp[wp] = inc;
            __syncthreads();
            uint32_t wb = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < BK_WARPS; w++) {
                const uint32_t t = s_warp[w];
                if (w < wp) wb += t;
                tot += t;
            }
            if (i < E) g.e_slot[i] = ((carry + wb + inc - v) << 1) | v;     // unique id << 1 | is-first
            carry += tot;
            __syncthreads();
        }
        if (threadIdx.x == 0) s_misc[0] = carry;
        __syncthreads();
    }
    const uint32_t U = s_misc[0];
    for (uint32_t e = threadIdx.x; e < E; e += BK_THREADS) {
        const uint32_t x = g.e_slot[e];
        if (x & 1u) {
            const uint32_t u = x >> 1;
            const uint64_t lo = g.e_lo[e], hi = HI ? g.e_hi[e] : 0;
            const uint32_t s = grp_find(g, g.e_bl[e], lo, hi);
            g.u_lo[u] = lo;
            if (HI) g.u_hi[u] = hi;
            g.u_pm[u] = g.e_pm[e];
            g.u_cnt[u] = (uint8_t)(g.h_count[s] & 0xFFu);            // uint8 counter of the reference wraps at 256
            g.u_bl[u] = g.e_bl[e];
            g.h_uniq[s] = (IDX)u;
            g.seen[u] = 0;
        }
    }
    for (uint32_t b = threadIdx.x; b < NB; b += BK_THREADS) g.b_us[b] = g.e_slot[g.p_eoff[g.b_hp[b]]] >> 1;
    if (threadIdx.x == 0) { g.b_us[NB] = U; g.seen[U] = 1; }         // seen[U]: the "no neighbour" sentinel
    __syncthreads();
    PP_PHASE(4);
    // ---- neighbour table (find_next, :566-602): for every unique k-mer its 4 left and 4 right neighbours that
    // are in the bucket with count >= abundance, in probe order
    for (uint32_t idx = threadIdx.x; idx < U * 8; idx += BK_THREADS) {
        const uint32_t u = idx >> 3, t = idx & 3;
        const bool left = !(idx & 4);
        const K128 cand = kmer_neighbour<HI>(K128{g.u_lo[u], HI ? g.u_hi[u] : 0}, left, (int)t, k);
        const uint32_t s = grp_find(g, g.u_bl[u], cand.lo, cand.hi);
        uint32_t w = U;
        if (s != ~0u) {
            w = g.h_uniq[s];
            if (g.u_cnt[w] < a.abundance) w = U;
        }
        g.adj[idx] = (IDX)w;
    }
    __syncthreads();
    PP_PHASE(5);
    // ---- chains (find_first_kmer + reconstruct_superkmer, :512-620) and byte sizes (:459-504), one thread per bucket
    // One lane per bucket.  A chain is one long dependent sequence and lanes of one warp that follow different chains
    // are issued one after the other, so the walkers of a group sit in different warps (measured per group: 12
    // walkers in one warp 57 k cycles, spread over the 8 warps 33 k; a flat state-machine loop instead of the nested
    // loops: 42 k).
    for (uint32_t b = (uint32_t)lane * BK_WARPS + warp; b < NB; b += BK_THREADS) {
        const uint32_t us = g.b_us[b], ue = g.b_us[b + 1];
        uint32_t nv = 0, n_max = 0, text_len = 0;
        const uint32_t minimizer = (uint32_t)(bucket_key(first + g.b_hp[b]) & (((uint64_t)1 << a.input_shift) - 1));
        for (uint32_t start = us; start < ue; start++) {
            // find_first_kmer (:604-620): the k-mers are stored in insertion order
            if (g.seen[start] || g.u_cnt[start] < a.abundance) continue;
            g.seen[start] = 1;
            const uint32_t v0 = nv;
            g.visit[us + nv++] = (IDX)((start - us) | (VIS_START << T::ROLE_SHIFT));
            const uint32_t pms = g.u_pm[start];
            // n_left = d - pos_min as uint64 in the reference: a k-mer without the minimizer text (pos 255)
            // extends to the left for as long as it can
            uint32_t n_left = pms == 255u ? 0x7fffffffu : (uint32_t)d - pms, n_right = pms;
            uint32_t cur = start, ext = 0, n_l = 0;          // ext = k-mers added to the start k-mer
            while (ext != (uint32_t)d) {                     // |sk| != 2k-m
                const bool left = n_left != 0;
                if (!left && n_right == 0) break;
                // find_next (:566-602): first neighbour in probe order that is in the bucket and unseen
                // the four candidates of one side are one aligned 8-byte (16-byte) word: one load, not four
                const IDX *ad = g.adj + (size_t)cur * 8 + (left ? 0 : 4);
                uint32_t c0, c1, c2, c3;
                if (sizeof(IDX) == 2) {
                    const uint2 v = *reinterpret_cast<const uint2 *>(ad);
                    c0 = v.x & 0xFFFFu; c1 = v.x >> 16; c2 = v.y & 0xFFFFu; c3 = v.y >> 16;
                } else {
                    const uint4 v = *reinterpret_cast<const uint4 *>(ad);
                    c0 = v.x; c1 = v.y; c2 = v.z; c3 = v.w;
                }
                const uint32_t s0 = g.seen[c0], s1 = g.seen[c1], s2 = g.seen[c2], s3 = g.seen[c3];
                const uint32_t found = !s0 ? c0 : !s1 ? c1 : !s2 ? c2 : !s3 ? c3 : U;
                const bool ok = found != U;
                if (ok) {
                    g.seen[found] = 1;
                    g.visit[us + nv++] = (IDX)((found - us) | ((left ? VIS_LEFT : VIS_RIGHT) << T::ROLE_SHIFT));
                    ext++;
                    n_l += left ? 1u : 0u;
                }
                if (left) {
                    n_left = ok ? n_left - 1 : 0;
                    cur = n_left == 0 ? start : found;
                } else {
                    if (!ok) break;
                    n_right--;
                    cur = found;
                }
            }
            const int len = k + (int)ext;
            if (len == full) {
                n_max++;                                                         // :479-485
            } else if (pms != 255u) {
                text_len += (uint32_t)(len - m) + 2;                             // the minimizer text is in the start k-mer
            } else {
                // rare: the start k-mer does not contain the minimizer text; look for it in the whole super-k-mer
                uint8_t sk[RC_SK];
                const int lo0 = 64 - (int)n_l;
                const K128 skey{g.u_lo[start], HI ? g.u_hi[start] : 0};
                for (int i = 0; i < k; i++) sk[64 + i] = (uint8_t)(k_shr(skey, 2 * (k - 1 - i)).lo & 3);
                uint32_t li = 0, ri = 0;
                for (uint32_t t = v0 + 1; t < nv; t++) {
                    const uint32_t x = g.visit[us + t], u = us + (x & T::MASK);
                    const K128 key{g.u_lo[u], HI ? g.u_hi[u] : 0};
                    if ((x >> T::ROLE_SHIFT) == VIS_LEFT) sk[64 - 1 - (li++)] = (uint8_t)(k_shr(key, 2 * k - 2).lo & 3);
                    else sk[64 + k + (ri++)] = (uint8_t)(key.lo & 3);
                }
                const int q = find_minimizer_codes(sk + lo0, len, m, minimizer);
                text_len += (uint32_t)(q < 0 ? len : len - m) + 2;
            }
        }
        if (nv < ue - us) g.visit[us + nv] = (IDX)T::END;       // k-mers below the abundance are never visited
        const uint32_t sz = n_max ? 1 + n_max * (uint32_t)(2 * d) / 4 : 0;
        g.b_bytes[b] = (uint32_t)m + 4 + sz + text_len + 2;
        g.b_nmax[b] = n_max;
    }
    if (a.dbg) { __syncthreads(); PP_PHASE(6); }
    // ---- compare elements (Comparator.cpp:97-264): a unique oriented k-mer is in the sketch iff count >= abundance;
    // the two orientations of one k-mer collapse
    for (uint32_t u = threadIdx.x; u < U; u += BK_THREADS) {
        bool keep = g.u_cnt[u] >= a.abundance;
        if (keep) {
            const K128 key{g.u_lo[u], HI ? g.u_hi[u] : 0};
            const K128 rc = k_rc_t<HI>(key, k);
            if (k_lt(rc, key)) {
                const uint32_t s = grp_find(g, g.u_bl[u], rc.lo, rc.hi);
                if (s != ~0u && g.u_cnt[g.h_uniq[s]] >= a.abundance) keep = false;
            }
        }
        g.e_slot[u] = keep ? 1u : 0u;
    }
    __syncthreads();
    PP_PHASE(11);
    {
        const int ln = lane, wp = warp;
        uint32_t carry = 0;
        for (uint32_t b0 = 0; b0 < U; b0 += BK_THREADS) {
            const uint32_t i = b0 + threadIdx.x;
            const uint32_t v = i < U ? g.e_slot[i] : 0;
            uint32_t inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (ln >= o) inc += t;
            }
            if (ln == 31) s_warp[wp] = inc;
            __syncthreads();
            uint32_t wb = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < BK_WARPS; w++) {
                const uint32_t t = s_warp[w];
                if (w < wp) wb += t;
                tot += t;
            }
            if (i < U) g.e_slot[i] = ((carry + wb + inc - v) << 1) | v;     // element offset << 1 | keep
            carry += tot;
            __syncthreads();
        }
        if (threadIdx.x == 0) s_misc[1] = carry;
        __syncthreads();
    }
    const uint32_t NE = s_misc[1];
    block_excl_scan(g.b_bytes, NB, s_warp);              // byte offset of every bucket inside the group
    const uint32_t bytes_total = g.b_bytes[NB];
    PP_PHASE(7);
    uint8_t *sk = s_sk + warp * RC_SK;
    auto emit_buckets = [&](uint8_t *obase, uint32_t w0, uint32_t nw) {
    for (uint32_t b = w0; b < NB; b += nw) {
        const uint32_t us = g.b_us[b], nb = g.b_us[b + 1] - us;
        const uint32_t minimizer = (uint32_t)(bucket_key(first + g.b_hp[b]) & (((uint64_t)1 << a.input_shift) - 1));
        uint8_t *o = obase + g.b_bytes[b];
        const uint32_t nm = g.b_nmax[b];
        const uint32_t sz = nm ? 1 + nm * (uint32_t)(2 * d) / 4 : 0;        // strCompressor: mod byte + 4 bases/byte
        if (lane == 0) {
            for (int i = 0; i < m; i++) o[i] = "ACTG"[(minimizer >> (2 * (m - 1 - i))) & 3];
            o[m] = (uint8_t)sz; o[m + 1] = (uint8_t)(sz >> 8); o[m + 2] = (uint8_t)(sz >> 16); o[m + 3] = (uint8_t)(sz >> 24);
            if (sz) o[m + 4] = 0;                                           // 2d is a multiple of 4: mod byte 0
        }
        uint8_t *p_max = o + m + 4 + (sz ? 1 : 0);
        uint8_t *p_txt = o + m + 4 + sz;
        uint32_t j = 0;
        while (j < nb) {
            const uint32_t head = g.visit[us + j];
            if (head == T::END) break;
            // one super-k-mer: its start entry and the extensions that follow it (lefts first, then rights)
            uint32_t e = j + 1, n_l = 0;
            while (e < nb) {
                const uint32_t x = g.visit[us + e];
                if (x == T::END || (x >> T::ROLE_SHIFT) == VIS_START) break;
                n_l += (x >> T::ROLE_SHIFT) == VIS_LEFT;
                e++;
            }
            const int lo = 64 - (int)n_l, len = k + (int)(e - j - 1);
            const uint32_t su = us + (head & T::MASK);
            const K128 skey{g.u_lo[su], HI ? g.u_hi[su] : 0};
            __syncwarp();
            for (int i = lane; i < k; i += 32) sk[64 + i] = (uint8_t)(k_shr(skey, 2 * (k - 1 - i)).lo & 3);
            for (uint32_t t = j + 1 + lane; t < e; t += 32) {
                const uint32_t x = g.visit[us + t], u = us + (x & T::MASK);
                const K128 key{g.u_lo[u], HI ? g.u_hi[u] : 0};
                const uint32_t r = t - (j + 1);
                if ((x >> T::ROLE_SHIFT) == VIS_LEFT) sk[64 - 1 - r] = (uint8_t)(k_shr(key, 2 * k - 2).lo & 3);   // r-th left extension
                else sk[64 + k + (r - n_l)] = (uint8_t)(key.lo & 3);                                               // right extension
            }
            __syncwarp();
            if (len == full) {
                for (int q = lane; q < 2 * d / 4; q += 32) {
                    uint32_t acc = 0;
                    for (int c = 0; c < 4; c++) {
                        const int i = 4 * q + c;
                        acc = (acc << 2) | ((i < d) ? sk[lo + i] : sk[lo + k + (i - d)]);
                    }
                    p_max[q] = (uint8_t)acc;
                }
                p_max += 2 * d / 4;
            } else {                                                            // :486-494
                // leftmost occurrence of the minimizer in the super-k-mer (warp-parallel find())
                int best = 0x7fffffff;
                for (int t = lane; t + m <= len; t += 32) {
                    uint32_t w = 0;
                    for (int i = 0; i < m; i++) w = (w << 2) | sk[lo + t + i];
                    if (w == minimizer) { best = t; break; }
                }
                best = __reduce_min_sync(0xffffffffu, best);
                const int q = best == 0x7fffffff ? -1 : best;
                if (q < 0) {
                    for (int t = lane; t < len; t += 32) p_txt[t] = "ACTG"[sk[lo + t]];
                    if (lane == 0) { p_txt[len] = '\n'; p_txt[len + 1] = '\n'; }
                    p_txt += len + 2;
                } else {
                    for (int t = lane; t < q; t += 32) p_txt[t] = "ACTG"[sk[lo + t]];
                    uint8_t *p2 = p_txt + q + 1;
                    for (int t = q + m + lane; t < len; t += 32) p2[t - q - m] = "ACTG"[sk[lo + t]];
                    if (lane == 0) { p_txt[q] = '\n'; p2[len - q - m] = '\n'; }
                    p_txt += (len - m) + 2;
                }
            }
            j = e;
        }
        if (lane == 0) { p_txt[0] = '\n'; p_txt[1] = '\n'; }
    }
    };
    // ---- the group's bytes, assembled in shared memory when they fit (the entry keys are dead by now)
    constexpr uint32_t STAGE_CAP = BK_ECAP * 8;
    const bool staged = sizeof(IDX) == 2 && bytes_total <= STAGE_CAP;
    uint8_t *stage = reinterpret_cast<uint8_t *>(g.e_lo);
    if (a.lb_bytes) {
        // small batches: final offsets by decoupled look-back (warp 0), hidden behind the writer warps
        if (warp == 0) {
            const unsigned long long x = lookback(a.lb_bytes, grp, bytes_total);
            if (lane == 0) { *reinterpret_cast<unsigned long long *>(s_misc + 2) = x; }
            const unsigned long long y = lookback(a.lb_elems, grp, NE);
            if (lane == 0) { *reinterpret_cast<unsigned long long *>(s_misc + 4) = y; }
        } else if (staged) {
            emit_buckets(stage, (uint32_t)warp - 1, BK_WARPS - 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long bb = *reinterpret_cast<unsigned long long *>(s_misc + 2);
            const unsigned long long eb = *reinterpret_cast<unsigned long long *>(s_misc + 4);
            atomicAdd(&a.cnt->body_bytes, (unsigned long long)bytes_total);
            atomicAdd(&a.cnt->n_elems, (unsigned long long)NE);
            atomicAdd(&a.cnt->tmp_bytes, (unsigned long long)bytes_total);
            atomicAdd(&a.cnt->n_buckets, (unsigned long long)NB);
            unsigned ovf = 0;
            if (bb + bytes_total > a.body_cap) ovf |= OVF_BODY;
            if (eb + NE > a.elems_cap) ovf |= OVF_ELEMS;
            if (ovf) atomicOr(&a.cnt->overflow, ovf);
        }
    } else {
    if (staged) emit_buckets(stage, (uint32_t)warp, BK_WARPS);
    // ---- temporary output ranges
    if (threadIdx.x == 0) {
        const unsigned long long tb = atomicAdd(&a.cnt->tmp_bytes, (unsigned long long)bytes_total);
        const unsigned long long te = atomicAdd(&a.cnt->tmp_elems, (unsigned long long)NE);
        *reinterpret_cast<unsigned long long *>(s_misc + 2) = tb;
        *reinterpret_cast<unsigned long long *>(s_misc + 4) = te;
        atomicAdd(&a.cnt->n_buckets, (unsigned long long)NB);
        unsigned ovf = 0;
        if (tb + bytes_total > a.body_cap) ovf |= OVF_BODY;
        if (te + NE > a.elems_cap) ovf |= OVF_ELEMS;
        if (ovf) atomicOr(&a.cnt->overflow, ovf);
        else { a.g_tb[grp] = tb; a.g_nb[grp] = bytes_total; a.g_te[grp] = te; a.g_ne[grp] = NE; }
    }
    __syncthreads();
    }
    const unsigned long long byte_base = *reinterpret_cast<unsigned long long *>(s_misc + 2);
    const unsigned long long elem_base = *reinterpret_cast<unsigned long long *>(s_misc + 4);
    PP_PHASE(8);
    if (a.dbg && threadIdx.x == 0) { a.dbg[(size_t)grp * 16 + 12] = E; a.dbg[(size_t)grp * 16 + 13] = U; a.dbg[(size_t)grp * 16 + 14] = NB; }
    if (byte_base + bytes_total > a.body_cap || elem_base + NE > a.elems_cap) return;
    // ---- where every input's bytes / elements begin: buckets are sorted by input first
    for (uint32_t b = threadIdx.x; b < NB; b += BK_THREADS) {
        const uint32_t hp = first + g.b_hp[b];
        const uint64_t in = bucket_key(hp) >> a.input_shift;
        if (hp == 0 || (bucket_key(hp - 1) >> a.input_shift) != in) {
            if (a.lb_bytes) {
                a.in_first_byte[in] = byte_base + g.b_bytes[b] + 1;                  // + 1: 0 means "no bucket"
                a.in_first_elem[in] = elem_base + (g.e_slot[g.b_us[b]] >> 1) + 1;
            } else {
                a.in_first_byte[in] = ((unsigned long long)(grp + 1) << 32) | g.b_bytes[b];
                a.in_first_elem[in] = ((unsigned long long)(grp + 1) << 32) | (g.e_slot[g.b_us[b]] >> 1);
            }
        }
    }
    // ---- elements
    for (uint32_t u = threadIdx.x; u < U; u += BK_THREADS) {
        const uint32_t x = g.e_slot[u];
        if (x & 1u) {
            K128 key{g.u_lo[u], HI ? g.u_hi[u] : 0};
            const K128 rc = k_rc_t<HI>(key, k);
            if (k_lt(rc, key)) key = rc;
            const unsigned long long o = elem_base + (x >> 1);
            a.tmp_min[o] = (uint32_t)(bucket_key(first + g.b_hp[g.u_bl[u]]) & (((uint64_t)1 << a.input_shift) - 1));
            a.tmp_klo[o] = key.lo;
            if (HI) a.tmp_khi[o] = key.hi;
        }
    }
    if (a.dbg) { __syncthreads(); PP_PHASE(9); }
    // ---- the sketch bytes
    if (staged) {
        uint8_t *dst = a.tmp_body + byte_base;
        for (uint32_t i = threadIdx.x; i < bytes_total; i += BK_THREADS) dst[i] = stage[i];
    } else {
        emit_buckets(a.tmp_body + byte_base, (uint32_t)warp, BK_WARPS);
    }
    if (a.dbg) { __syncthreads(); PP_PHASE(10); }
}

template <bool HI>
__global__ void __launch_bounds__(BK_THREADS) pp_bucket_kernel(BucketArgs a)
{
    extern __shared__ __align__(16) uint8_t bk_smem[];
    __shared__ uint32_t s_warp[BK_WARPS + 1];
    __shared__ uint32_t s_misc[8];
    __shared__ uint32_t s_first, s_end;
    __shared__ __align__(16) uint8_t s_sk[BK_WARPS * RC_SK];
    const uint32_t grp = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long npt = a.cnt->n_pieces;
    if (npt > a.pieces_cap) npt = 0;                             // overflow: the host retries
    const uint32_t n_pieces = (uint32_t)npt;
    auto bucket_key = [&](uint32_t i) -> uint64_t { return a.skey64 ? a.skey64[i] : (uint64_t)a.skey32[i]; };
    auto is_head = [&](uint32_t i) { return i == 0 || bucket_key(i) != bucket_key(i - 1); };
    const uint64_t lo64 = (uint64_t)grp * a.pp;
    auto publish_empty = [&]() {
        if (a.lb_bytes && warp == 0) { lookback(a.lb_bytes, grp, 0); lookback(a.lb_elems, grp, 0); }
    };
    if (lo64 >= n_pieces) {                                      // behind the data: nothing looks back at these groups
        if (a.lb_bytes && threadIdx.x == 0) {
            reinterpret_cast<volatile unsigned long long *>(a.lb_bytes)[grp] = 1ULL << 62;
            reinterpret_cast<volatile unsigned long long *>(a.lb_elems)[grp] = 1ULL << 62;
        }
        return;
    }
    const uint32_t lo = (uint32_t)lo64, nominal_end = (uint32_t)min((uint64_t)n_pieces, lo64 + a.pp);
    if (threadIdx.x == 0) { s_first = 0xFFFFFFFFu; s_end = n_pieces; }
    __syncthreads();
    for (uint32_t i = lo + threadIdx.x; i < nominal_end; i += BK_THREADS)
        if (is_head(i)) atomicMin(&s_first, i);
    __syncthreads();
    const uint32_t first = s_first;
    if (first == 0xFFFFFFFFu) { publish_empty(); return; }       // no bucket starts in this group's range
    // the group ends where the first bucket of a later range starts
    for (uint32_t b0 = nominal_end; b0 < n_pieces; b0 += BK_THREADS) {
        const uint32_t i = b0 + threadIdx.x;
        const bool h = i < n_pieces && is_head(i);
        if (h) atomicMin(&s_end, i);
        if (__syncthreads_or(h)) break;
    }
    __syncthreads();
    const uint32_t end = s_end, np = end - first;
    // entries of the group
    uint32_t part = 0;
    for (uint32_t p = threadIdx.x; p < np; p += BK_THREADS) part += a.pieces[a.order[first + p]].w;
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) s_warp[warp] = part;
    __syncthreads();
    uint32_t E = 0;
    for (int w = 0; w < BK_WARPS; w++) E += s_warp[w];
    __syncthreads();
    if (np <= (uint32_t)BK_PMAX && E <= (uint32_t)BK_ECAP) {
        GroupMem<uint16_t> g;
        group_carve<uint16_t>(g, bk_smem, BK_PMAX, BK_ECAP, BK_SLOTS, HI);
        for (uint32_t p = threadIdx.x; p < np; p += BK_THREADS) g.pc[p] = a.pieces[a.order[first + p]];
        __syncthreads();
        process_group<uint16_t, HI>(a, g, grp, first, np, s_warp, s_sk, s_misc);
    } else {
        // a group too large for shared memory: the same code on arrays carved from the global pool
        uint32_t S = 1024;
        while (S < 2 * E) S <<= 1;
        GroupMem<uint32_t> g;
        const size_t need = group_carve<uint32_t>(g, nullptr, np, E, S, HI);
        if (threadIdx.x == 0) {
            const unsigned long long off = atomicAdd(&a.cnt->big_top, (unsigned long long)need);
            *reinterpret_cast<unsigned long long *>(s_misc + 6) = off;
        }
        __syncthreads();
        const unsigned long long off = *reinterpret_cast<unsigned long long *>(s_misc + 6);
        __syncthreads();
        if (off + need > a.big_cap) {
            if (threadIdx.x == 0) atomicOr(&a.cnt->overflow, (unsigned)OVF_BIG);
            publish_empty();
            return;
        }
        group_carve<uint32_t>(g, a.big_pool + off, np, E, S, HI);
        for (uint32_t p = threadIdx.x; p < np; p += BK_THREADS) g.pc[p] = a.pieces[a.order[first + p]];
        __syncthreads();
        process_group<uint32_t, HI>(a, g, grp, first, np, s_warp, s_sk, s_misc);
    }
}

// Final offsets of the groups' outputs: exclusive prefix sums of their byte and element counts, in group (= bucket)
// order, into g_bb / g_be; the totals go to the counters.  One CTA.
__global__ void __launch_bounds__(1024) pp_offsets_kernel(const unsigned long long *__restrict__ g_nb, const unsigned long long *__restrict__ g_ne,
                                                          unsigned long long *__restrict__ g_bb, unsigned long long *__restrict__ g_be,
                                                          uint64_t n_groups, Counters *cnt)
{
    __shared__ unsigned long long s_w[2][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long run_b = 0, run_e = 0;
    for (uint64_t g0 = 0; g0 < n_groups; g0 += 1024) {
        const uint64_t g = g0 + threadIdx.x;
        const unsigned long long vb = g < n_groups ? g_nb[g] : 0, ve = g < n_groups ? g_ne[g] : 0;
        unsigned long long ib = vb, ie = ve;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long tb = __shfl_up_sync(0xffffffffu, ib, o), te = __shfl_up_sync(0xffffffffu, ie, o);
            if (lane >= o) { ib += tb; ie += te; }
        }
        if (lane == 31) { s_w[0][warp] = ib; s_w[1][warp] = ie; }
        __syncthreads();
        if (warp == 0) {
            const unsigned long long tb = s_w[0][lane], te = s_w[1][lane];
            unsigned long long xb = tb, xe = te;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long ub = __shfl_up_sync(0xffffffffu, xb, o), ue = __shfl_up_sync(0xffffffffu, xe, o);
                if (lane >= o) { xb += ub; xe += ue; }
            }
            s_w[0][lane] = xb - tb; s_w[1][lane] = xe - te;
            if (lane == 31) { s_w[0][32] = xb; s_w[1][32] = xe; }
        }
        __syncthreads();
        if (g < n_groups) { g_bb[g] = run_b + s_w[0][warp] + ib - vb; g_be[g] = run_e + s_w[1][warp] + ie - ve; }
        run_b += s_w[0][32]; run_e += s_w[1][32];
        __syncthreads();
    }
    if (threadIdx.x == 0) { cnt->body_bytes = run_b; cnt->n_elems = run_e; }
}

// Moves every group's bytes and elements from its temporary range to their final place, and turns the inputs'
// (group, offset inside the group) markers into final offsets (+ 1).  grid.x covers the groups, the CTAs behind
// them the inputs.
__global__ void __launch_bounds__(256) pp_gather_kernel(const unsigned long long *__restrict__ g_tb, const unsigned long long *__restrict__ g_nb,
                                                        const unsigned long long *__restrict__ g_te, const unsigned long long *__restrict__ g_ne,
                                                        const unsigned long long *__restrict__ g_bb, const unsigned long long *__restrict__ g_be,
                                                        uint64_t n_groups, const uint8_t *__restrict__ tmp_body, uint8_t *__restrict__ body,
                                                        uint64_t body_cap, const uint32_t *__restrict__ tmp_min,
                                                        const uint64_t *__restrict__ tmp_klo, const uint64_t *__restrict__ tmp_khi,
                                                        uint32_t *__restrict__ el_min, uint64_t *__restrict__ el_klo, uint64_t *__restrict__ el_khi,
                                                        uint64_t elems_cap, unsigned long long *__restrict__ in_first_byte,
                                                        unsigned long long *__restrict__ in_first_elem, uint32_t n_inputs, const Counters *cnt)
{
    if (cnt->overflow) return;                                   // the host retries with larger buffers
    const uint64_t g = blockIdx.x;
    if (g < n_groups) {
        const unsigned long long nb = g_nb[g], ne = g_ne[g];
        if (nb) {
            const uint8_t *src = tmp_body + g_tb[g];
            uint8_t *dst = body + g_bb[g];
            if (g_bb[g] + nb <= body_cap)
                for (unsigned long long i = threadIdx.x; i < nb; i += blockDim.x) dst[i] = src[i];
        }
        if (ne && g_be[g] + ne <= elems_cap) {
            const unsigned long long so = g_te[g], d0 = g_be[g];
            for (unsigned long long i = threadIdx.x; i < ne; i += blockDim.x) {
                el_min[d0 + i] = tmp_min[so + i];
                el_klo[d0 + i] = tmp_klo[so + i];
                if (el_khi) el_khi[d0 + i] = tmp_khi[so + i];
            }
        }
        return;
    }
    const uint64_t in = (g - n_groups) * blockDim.x + threadIdx.x;
    if (in < n_inputs) {
        const unsigned long long vb = in_first_byte[in], ve = in_first_elem[in];
        if (vb) in_first_byte[in] = g_bb[(vb >> 32) - 1] + (vb & 0xFFFFFFFFull) + 1;     // + 1: 0 means "no bucket"
        if (ve) in_first_elem[in] = g_be[(ve >> 32) - 1] + (ve & 0xFFFFFFFFull) + 1;
    }
}

__global__ void pp_iota_kernel(uint32_t *__restrict__ v, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}

// ---------------------------------------------------------------- driver

static void trace_alloc(const char *what, size_t from, size_t to)
{
    static const bool on = getenv("SPSP_TRACE_ALLOC") != nullptr;
    if (on) fprintf(stderr, "[alloc] post-pass %s buffer %zu -> %zu bytes\n", what, from, to);
}
struct DBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        trace_alloc("device", cap, bytes);
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    ~DBuf() { if (p) cudaFree(p); }
};
struct HBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        trace_alloc("pinned", cap, bytes);
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 2 + 4096;              // pinned memory is slow to allocate: grow in big steps
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
    ~HBuf() { if (p) cudaFreeHost(p); }
};

// One device arena per slot (a single allocation, grow-only) + the output buffers that must survive the call.
struct PostpassBuffers {
    DBuf arena, elems, body, big, tmp_elems, tmp_body;
    HBuf h_small, h_body, h_off;
    // capacities that were found too small once (a retry raised them): kept for the next batches
    uint64_t pieces_cap_min = 0, body_cap_min = 0, big_cap_min = 0;
    uint64_t last_body_bytes = 0;
    bool force_large_sort = false;
};

PostpassBuffers *postpass_buffers_create() { return new PostpassBuffers(); }
void postpass_buffers_destroy(PostpassBuffers *b) { delete b; }

#define PP_CK(call)                                  \
    do {                                             \
        cudaError_t e_ = (call);                     \
        if (e_ != cudaSuccess) return e_;            \
    } while (0)

static inline unsigned nblk(uint64_t n, unsigned t = 256) { return (unsigned)((n + t - 1) / t ? (n + t - 1) / t : 1); }
static int bits_for(uint64_t v)
{
    int b = 0;
    while (b < 64 && (v >> b)) b++;
    return b ? b : 1;
}

struct Arena {
    uint8_t *base;
    size_t off = 0;
    template <class T> T *take(size_t n)
    {
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off = (off + n * sizeof(T) + 255) & ~(size_t)255;
        return p;
    }
};

cudaError_t postpass_run(PostpassBuffers *b, const PostpassIn &in, PostpassOut *out, cudaStream_t st)
{
    const int k = in.k, m = in.m, d = k - m;
    const bool hi128 = k > 32;
    const uint64_t H = in.hits_cap ? in.hits_cap : 1;
    if (in.n_bases >= (1ULL << 32)) return cudaErrorInvalidValue;    // positions are sorted as 32-bit keys
    // every k-mer is in at most one piece: entries <= min(hits * (d+1), bases)
    uint64_t e_cap = H * (uint64_t)(d + 1);
    if (e_cap > in.n_bases) e_cap = in.n_bases;
    if (e_cap < 1) e_cap = 1;
    if (e_cap >= (1ULL << 31)) return cudaErrorInvalidValue;
    const int input_shift = 2 * m;                                   // bucket id = input << 2m | minimizer
    const int a_bits = input_shift + bits_for(in.n_inputs ? in.n_inputs - 1 : 0);
    const bool key64 = a_bits > 31;                                  // 32-bit keys keep one value above every bucket for the padding
    const size_t nin = in.n_inputs ? in.n_inputs : 1;
    const uint32_t pp = (uint32_t)std::min(32 / SPSP_BK_SCALE, std::max(2, BK_PP_ENTRIES / (d + 1)));
    out->retry = 0;

    const uint64_t p_cap = std::max<uint64_t>(H + H / 8 + 1024, b->pieces_cap_min);
    const bool small_hits = !b->force_large_sort && H <= 32768;
    const bool small_pc = !b->force_large_sort && p_cap <= 32768 && !key64;
    const uint64_t n_groups = (p_cap + pp - 1) / pp;
    uint64_t body_cap = std::max<uint64_t>(e_cap * 3 + (1u << 20), b->body_cap_min);
    uint64_t big_cap = std::max<uint64_t>((uint64_t)32 << 20, b->big_cap_min);
    const unsigned rp_grid = nblk(H, RP_THREADS);

    // ---- one arena for everything that lives only during the call
    size_t cub_bytes = 0;
    if (!small_hits) {
        size_t t = 0;
        PP_CK(cub::DeviceRadixSort::SortPairs(nullptr, t, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                              (uint32_t *)nullptr, (int)H, 0, 32, st));
        cub_bytes = std::max(cub_bytes, t);
    }
    if (!small_pc) {
        size_t t = 0;
        if (key64)
            PP_CK(cub::DeviceRadixSort::SortPairs(nullptr, t, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr,
                                                  (uint32_t *)nullptr, (int)p_cap, 0, a_bits + 1, st));
        else
            PP_CK(cub::DeviceRadixSort::SortPairs(nullptr, t, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                                  (uint32_t *)nullptr, (int)p_cap, 0, a_bits + 1, st));
        cub_bytes = std::max(cub_bytes, t);
    }
    Arena ar{nullptr};
    Counters *cnt = nullptr;
    unsigned long long *in_sel = nullptr, *in_fb = nullptr, *in_fe = nullptr;
    unsigned long long *g_tb = nullptr, *g_nb = nullptr, *g_te = nullptr, *g_ne = nullptr, *g_bb = nullptr, *g_be = nullptr;
    uint32_t *hkey = nullptr, *hval = nullptr, *hkey2 = nullptr, *hval2 = nullptr, *cl_np = nullptr, *cl_nk = nullptr, *cta_np = nullptr;
    uint4 *pieces = nullptr;
    uint32_t *pkey32 = nullptr, *skey32 = nullptr, *pidx = nullptr, *order = nullptr;
    uint64_t *pkey64 = nullptr, *skey64 = nullptr;
    uint8_t *cubtmp = nullptr;
    size_t small_bytes = 0, zero_bytes = 0;
    for (int pass = 0; pass < 2; pass++) {
        ar.off = 0;
        // [zeroed block] counters, per-input results (copied to the host as one block), look-back states
        cnt = ar.take<Counters>(1);
        in_sel = ar.take<unsigned long long>(nin);
        in_fb = ar.take<unsigned long long>(nin);          // first byte / element of every input (+ 1)
        in_fe = ar.take<unsigned long long>(nin);
        small_bytes = ar.off;
        g_tb = ar.take<unsigned long long>(n_groups); g_nb = ar.take<unsigned long long>(n_groups);
        g_te = ar.take<unsigned long long>(n_groups); g_ne = ar.take<unsigned long long>(n_groups);
        zero_bytes = ar.off;
        g_bb = ar.take<unsigned long long>(n_groups); g_be = ar.take<unsigned long long>(n_groups);
        hkey = ar.take<uint32_t>(H); hval = ar.take<uint32_t>(H);
        if (!small_hits) { hkey2 = ar.take<uint32_t>(H); hval2 = ar.take<uint32_t>(H); }
        cl_np = ar.take<uint32_t>(H); cl_nk = ar.take<uint32_t>(H);
        cta_np = ar.take<uint32_t>(rp_grid);
        pieces = ar.take<uint4>(p_cap);
        if (key64) { pkey64 = ar.take<uint64_t>(p_cap); skey64 = ar.take<uint64_t>(p_cap); }
        else { pkey32 = ar.take<uint32_t>(p_cap); skey32 = ar.take<uint32_t>(p_cap); }
        order = ar.take<uint32_t>(p_cap);
        if (!small_pc) pidx = ar.take<uint32_t>(p_cap);
        cubtmp = ar.take<uint8_t>(cub_bytes ? cub_bytes : 16);
        if (pass == 0) {
            PP_CK(b->arena.ensure(ar.off));
            ar.base = static_cast<uint8_t *>(b->arena.p);
        }
    }
    PP_CK(cudaMemsetAsync(ar.base, 0, zero_bytes, st));
    PP_CK(b->elems.ensure(e_cap * (4 + 8 + (hi128 ? 8 : 0)) + 64));
    uint64_t *el_klo = static_cast<uint64_t *>(b->elems.p);
    uint64_t *el_khi = hi128 ? el_klo + e_cap : nullptr;
    uint32_t *el_min = reinterpret_cast<uint32_t *>(el_klo + e_cap * (hi128 ? 2 : 1));
    uint64_t *tmp_klo = nullptr, *tmp_khi = nullptr;
    uint32_t *tmp_min = nullptr;
    if (!small_hits) {                                 // the two-pass form's temporary element ranges
        PP_CK(b->tmp_elems.ensure(e_cap * (4 + 8 + (hi128 ? 8 : 0)) + 64));
        tmp_klo = static_cast<uint64_t *>(b->tmp_elems.p);
        tmp_khi = hi128 ? tmp_klo + e_cap : nullptr;
        tmp_min = reinterpret_cast<uint32_t *>(tmp_klo + e_cap * (hi128 ? 2 : 1));
    }
    PP_CK(b->body.ensure(body_cap));
    body_cap = b->body.cap;
    if (!small_hits) PP_CK(b->tmp_body.ensure(body_cap));
    PP_CK(b->big.ensure(big_cap));
    big_cap = b->big.cap;
    uint32_t launched = 0;

    // ---- S1: valid hits sorted by position
    const int pos_bits = bits_for(in.n_bases);
    const uint32_t *skey = hkey, *sval = hval;
    if (small_hits) {
        PP_CK(launch_sort_small<true>(in.d_hits, in.d_hit_count, H, in.d_rec_begin, in.d_rec_end, in.n_rec, k, m, nullptr,
                                      pos_bits, hkey, hval, cnt, st));
        launched++;
    } else {
        pp_classify_kernel<<<nblk(H), 256, 0, st>>>(in.d_hits, in.d_hit_count, H, in.d_rec_begin, in.d_rec_end, in.n_rec, k, m,
                                                    hkey, hval, cnt);
        size_t t = cub_bytes;
        PP_CK(cub::DeviceRadixSort::SortPairs(cubtmp, t, hkey, hkey2, hval, hval2, (int)H, 0, 32, st));
        skey = hkey2; sval = hval2;
        launched++;
    }
    // ---- R1/R2: replay -> pieces in genome order
    ReplayArgs ra{};
    ra.key = skey; ra.val = sval; ra.rec_begin = in.d_rec_begin; ra.rec_end = in.d_rec_end; ra.rec_input = in.d_rec_input;
    ra.n_rec = in.n_rec; ra.k = k; ra.m = m; ra.input_shift = input_shift; ra.cl_np = cl_np; ra.cl_nk = cl_nk; ra.cta_np = cta_np;
    ra.pieces = pieces; ra.pkey32 = pkey32; ra.pkey64 = pkey64; ra.pieces_cap = p_cap; ra.in_sel = in_sel; ra.cnt = cnt;
    if (!small_pc) {                                   // the device-wide sort runs over the capacity: padding keys sort last
        if (key64) PP_CK(cudaMemsetAsync(pkey64, 0xFF, p_cap * 8, st));
        else PP_CK(cudaMemsetAsync(pkey32, 0xFF, p_cap * 4, st));
    }
    pp_replay_kernel<false><<<rp_grid, RP_THREADS, 0, st>>>(ra);
    pp_replay_kernel<true><<<rp_grid, RP_THREADS, 0, st>>>(ra);
    launched += 2;
    // ---- S2: pieces sorted by bucket (stable: genome order inside a bucket = the reference's insertion order)
    if (small_pc) {
        PP_CK(launch_sort_small<false>(nullptr, nullptr, 0, nullptr, nullptr, 0, k, m, pkey32, a_bits, skey32, order, cnt, st));
        launched++;
    } else {
        pp_iota_kernel<<<nblk(p_cap), 256, 0, st>>>(pidx, p_cap);
        launched++;
        size_t t = cub_bytes;
        if (key64)
            PP_CK(cub::DeviceRadixSort::SortPairs(cubtmp, t, pkey64, skey64, pidx, order, (int)p_cap, 0, a_bits + 1, st));
        else
            PP_CK(cub::DeviceRadixSort::SortPairs(cubtmp, t, pkey32, skey32, pidx, order, (int)p_cap, 0, a_bits + 1, st));
    }
    // ---- B: buckets -> sketch bytes + elements at their final place
    BucketArgs ba{};
    ba.packed = in.d_packed; ba.pieces = pieces; ba.order = order; ba.skey32 = skey32; ba.skey64 = skey64; ba.pp = pp;
    ba.pieces_cap = p_cap; ba.k = k; ba.m = m; ba.input_shift = input_shift; ba.abundance = in.abundance;
    const bool two_pass = !small_hits;                 // see BucketArgs
    ba.g_tb = g_tb; ba.g_nb = g_nb; ba.g_te = g_te; ba.g_ne = g_ne; ba.body_cap = body_cap; ba.elems_cap = e_cap;
    if (two_pass) {
        ba.tmp_body = static_cast<uint8_t *>(b->tmp_body.p);
        ba.tmp_min = tmp_min; ba.tmp_klo = tmp_klo; ba.tmp_khi = tmp_khi;
    } else {
        ba.lb_bytes = g_tb; ba.lb_elems = g_te;        // (zeroed with the group arrays)
        ba.tmp_body = static_cast<uint8_t *>(b->body.p);
        ba.tmp_min = el_min; ba.tmp_klo = el_klo; ba.tmp_khi = el_khi;
    }
    ba.in_first_byte = in_fb; ba.in_first_elem = in_fe; ba.big_pool = static_cast<uint8_t *>(b->big.p); ba.big_cap = big_cap;
    ba.cnt = cnt;
    static const bool pp_debug = getenv("SPSP_PP_DEBUG") != nullptr;
    DBuf dbgbuf;
    if (pp_debug) {
        PP_CK(dbgbuf.ensure(n_groups * 16 * 8));
        PP_CK(cudaMemsetAsync(dbgbuf.p, 0, n_groups * 16 * 8, st));
        ba.dbg = static_cast<long long *>(dbgbuf.p);
    }
    const size_t bk_smem = bucket_smem_bytes(hi128);
    {
        static PerDeviceOnce once[2];
        PP_CK(once[hi128 ? 1 : 0].run([&] {
            return hi128 ? cudaFuncSetAttribute(pp_bucket_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_smem)
                         : cudaFuncSetAttribute(pp_bucket_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_smem);
        }));
    }
    if (hi128) pp_bucket_kernel<true><<<(unsigned)n_groups, BK_THREADS, bk_smem, st>>>(ba);
    else pp_bucket_kernel<false><<<(unsigned)n_groups, BK_THREADS, bk_smem, st>>>(ba);
    launched++;
    PP_CK(cudaGetLastError());
    // ---- final offsets (groups are in bucket order) and the move of every group's output to its final place
    if (two_pass) {
        pp_offsets_kernel<<<1, 1024, 0, st>>>(g_nb, g_ne, g_bb, g_be, n_groups, cnt);
        const uint64_t grid = n_groups + (nin + 255) / 256;
        pp_gather_kernel<<<(unsigned)grid, 256, 0, st>>>(g_tb, g_nb, g_te, g_ne, g_bb, g_be, n_groups, ba.tmp_body,
                                                        static_cast<uint8_t *>(b->body.p), body_cap, tmp_min, tmp_klo, tmp_khi, el_min,
                                                        el_klo, el_khi, e_cap, in_fb, in_fe, (uint32_t)nin, cnt);
        launched += 2;
    }
    PP_CK(cudaGetLastError());
    // ---- results to the host: ONE synchronisation in the steady state.  The counters and per-input arrays travel as
    // one block; the sketch bytes are copied speculatively, sized by the previous batch, and topped up if this batch
    // turned out larger.
    PP_CK(b->h_small.ensure(small_bytes + 8));
    PP_CK(b->h_off.ensure((nin + 1) * 8 * 2));
    uint64_t pred = b->last_body_bytes ? std::min<uint64_t>(body_cap, b->last_body_bytes + b->last_body_bytes / 4 + 4096) : 0;
    // the guess must not regrow pinned memory by a few bytes (cudaHostAlloc of 13 MB was measured at 16 ms, five
    // times the whole pass): a buffer that holds the previous batch with some room is copied as far as it goes
    if (pred > b->h_body.cap && b->h_body.cap >= b->last_body_bytes + b->last_body_bytes / 16) pred = b->h_body.cap;
    if (pred) PP_CK(b->h_body.ensure(pred));
    uint8_t *hs = b->h_small.as<uint8_t>();
    PP_CK(cudaMemcpyAsync(hs, ar.base, small_bytes, cudaMemcpyDeviceToHost, st));
    PP_CK(cudaMemcpyAsync(hs + small_bytes, in.d_hit_count, 8, cudaMemcpyDeviceToHost, st));
    if (pred) PP_CK(cudaMemcpyAsync(b->h_body.p, b->body.p, pred, cudaMemcpyDeviceToHost, st));
    PP_CK(cudaStreamSynchronize(st));
    const Counters hc = *reinterpret_cast<const Counters *>(hs);
    if (pp_debug) {
        std::vector<long long> h(n_groups * 16);
        cudaMemcpy(h.data(), dbgbuf.p, h.size() * 8, cudaMemcpyDeviceToHost);
        double ph[11] = {0}; double n = 0, e = 0, u = 0, nb = 0;
        for (uint64_t g = 0; g < n_groups; g++) {
            const long long *r = &h[g * 16];
            if (!r[0] || !r[10]) continue;
            n++; e += r[12]; u += r[13]; nb += r[14];
            for (int i = 1; i <= 10; i++) ph[i] += (double)(r[i] - r[i - 1]);
        }
        fprintf(stderr, "[pp] groups %.0f  entries/grp %.0f uniques %.0f buckets %.1f | cycles per group:", n, e / n, u / n, nb / n);
        const char *nm[11] = {"", "pieces", "entries", "insert", "unique", "adjacency", "walk", "elemflag+scan", "stage bytes + ranges", "elements out", "bytes out"};
        for (int i = 1; i <= 10; i++) fprintf(stderr, " %s %.0f", nm[i], ph[i] / n);
        double f11 = 0;
        for (uint64_t g = 0; g < n_groups; g++) {
            const long long *r = &h[g * 16];
            if (!r[0] || !r[10]) continue;
            f11 += (double)(r[11] - r[6]);
        }
        fprintf(stderr, " | element flags alone %.0f\n", f11 / n);
    }
    const uint64_t n_hits = *reinterpret_cast<const uint64_t *>(hs + small_bytes);
    out->n_hits = n_hits;
    out->kernels_launched = launched;
    static const bool trace_retry = getenv("SPSP_TRACE_ALLOC") != nullptr;
    if (trace_retry && (n_hits > H || hc.overflow))
        fprintf(stderr, "[alloc] post-pass retry: hits %llu / cap %llu, overflow bits %u (pieces %llu body %llu big %llu)\n",
                (unsigned long long)n_hits, (unsigned long long)H, hc.overflow, hc.n_pieces, hc.body_bytes, hc.big_top);
    if (n_hits > H) { out->retry = PP_RETRY_HITS; return cudaSuccess; }      // the caller grows the hit buffer and rescans
    if (hc.overflow) {
        if (hc.overflow & OVF_PIECES) b->pieces_cap_min = hc.n_pieces + hc.n_pieces / 8 + 1024;
        if (hc.overflow & OVF_BODY) b->body_cap_min = hc.tmp_bytes + hc.tmp_bytes / 8 + 4096;     // (the demand of every group)
        if (hc.overflow & OVF_BIG) b->big_cap_min = hc.big_top + hc.big_top / 8 + 4096;
        if (hc.overflow & OVF_SORT) b->force_large_sort = true;
        if (hc.overflow & OVF_ELEMS) return cudaErrorUnknown;                 // cannot happen: the bound is exact
        out->retry = PP_RETRY_POSTPASS;
        return cudaSuccess;
    }
    if (hc.body_bytes > pred) {
        PP_CK(b->h_body.ensure(hc.body_bytes));
        PP_CK(cudaMemcpyAsync(b->h_body.p, b->body.p, hc.body_bytes, cudaMemcpyDeviceToHost, st));
        PP_CK(cudaStreamSynchronize(st));
    } else if (!b->h_body.p) {
        PP_CK(b->h_body.ensure(1));
    }
    b->last_body_bytes = hc.body_bytes;
    const unsigned long long *h_sel = reinterpret_cast<const unsigned long long *>(hs + ((uint8_t *)in_sel - ar.base));
    const unsigned long long *h_fb = reinterpret_cast<const unsigned long long *>(hs + ((uint8_t *)in_fb - ar.base));
    const unsigned long long *h_fe = reinterpret_cast<const unsigned long long *>(hs + ((uint8_t *)in_fe - ar.base));
    uint64_t *off = b->h_off.as<uint64_t>();
    uint64_t *eoffh = off + (nin + 1);
    off[in.n_inputs] = hc.body_bytes;
    eoffh[in.n_inputs] = hc.n_elems;
    for (uint32_t i = in.n_inputs; i-- > 0;) {          // an input without any bucket owns an empty range
        off[i] = h_fb[i] ? h_fb[i] - 1 : off[i + 1];
        eoffh[i] = h_fe[i] ? h_fe[i] - 1 : eoffh[i + 1];
    }
    out->h_body = b->h_body.as<uint8_t>();
    out->h_body_off = off;
    out->h_selected = reinterpret_cast<const uint64_t *>(h_sel);
    out->h_elem_off = eoffh;
    out->d_minim = el_min;
    out->d_klo = el_klo;
    out->d_khi = el_khi;
    out->n_elems = hc.n_elems;
    return cudaSuccess;
}

}  // namespace spsp
