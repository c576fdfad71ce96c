// Device post-pass of the sketch stage (see postpass.cuh).  Pipeline, all on
// one stream, no host round trip until the sizes of the results are known:
//
//   hits ──K1 classify──▶ CUB sort by position ──K2 clusters──▶ K3 replay (count, write)
//        ──▶ pieces ──K4 k-mer entries──▶ K5 global hash table keyed (bucket, k-mer): first
//        occurrence, count mod 256 ──▶ unique k-mers, ONE stable CUB sort by bucket ──▶ buckets
//        in insertion order ──K6 chain walk (visit order) ──▶ measure / emit ──▶ sketch bytes
//        ──K7 canonical elements──▶ compare stage (device resident)
//
// The generic steps (radix sort, prefix sum) are CUB library calls; everything
// that carries reference semantics is a kernel in this file, each citing the
// reference lines it restates (via csrc/host/postpass.cpp, which it mirrors).
#include "postpass.cuh"

#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"

namespace spsp {

// ------------------------------------------------------------------ helpers

struct DBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 1024;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
    ~DBuf() { if (p) cudaFree(p); }
};
struct HBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 1024;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
    ~HBuf() { if (p) cudaFreeHost(p); }
};

// counters kept on the device between kernels
struct Counters {
    unsigned long long n_valid, n_clusters, n_pieces, n_entries, n_unique, n_buckets, n_elems, body_bytes;
    unsigned long long chain_next;      // work queue of pp_chain_kernel (next bucket to walk)
};

struct PostpassBuffers {
    DBuf cnt, hkey, hval, hkey2, hval2, hhash, hrec, cflag, cid, cl_first, cl_np, cl_nk, cl_poff, cl_eoff;
    DBuf pc_first, pc_nk, pc_min, pc_meta, pc_eoff;
    DBuf eA, eklo, ekhi, epm, skey, skey2, head, uid;
    DBuf uA, uklo, ukhi, upm, ucnt, bflag, bidm, bstart, uidx0, uidx2, uent, eslot, hfirst, hcount, huniq, seen;
    DBuf visit, bbytes, bnmax, boff, body, in_bytes, in_sel, in_elems, eflag, eoff, el_min, el_klo, el_khi, cubtmp;
    HBuf h_cnt, h_body, h_in, h_off;
};

PostpassBuffers *postpass_buffers_create() { return new PostpassBuffers(); }
void postpass_buffers_destroy(PostpassBuffers *b) { delete b; }

struct K128 {
    uint64_t lo, hi;
};
__device__ __forceinline__ bool k_eq(const K128 &a, const K128 &b) { return a.lo == b.lo && a.hi == b.hi; }
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
// last record r with rec_begin[r] <= pos (n_rec >= 1, rec_begin[0] <= pos assumed checked by caller)
__device__ __forceinline__ long long find_rec(const uint64_t *__restrict__ rec_begin, uint64_t n_rec, uint64_t pos)
{
    uint64_t lo = 0, hi = n_rec;           // first index with begin > pos
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (rec_begin[mid] <= pos) lo = mid + 1; else hi = mid;
    }
    return (long long)lo - 1;
}

// ------------------------------------------------------------- K1 classify

// A hit counts only if its m-mer lies inside one record of at least k bases
// (host: build_sketch_t drops hits that straddle a record / records < k).
__global__ void pp_classify_kernel(const spsp_hit *__restrict__ hits, uint64_t n_hits, const uint64_t *__restrict__ rec_begin,
                                   const uint64_t *__restrict__ rec_end, uint64_t n_rec, int k, int m,
                                   uint64_t invalid_key, uint64_t *__restrict__ key, uint32_t *__restrict__ val,
                                   Counters *cnt)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_hits) return;
    const uint64_t pos = hits[i].pos;
    bool valid = false;
    if (n_rec) {
        long long r = find_rec(rec_begin, n_rec, pos);
        if (r >= 0) {
            uint64_t b = rec_begin[r], e = rec_end[r];
            valid = pos + (uint64_t)m <= e && e - b >= (uint64_t)k;
        }
    }
    key[i] = valid ? pos : invalid_key;             // sorts behind every position
    val[i] = (hits[i].canon << 1) | (hits[i].rev & 1u);
    if (valid) atomicAdd(&cnt->n_valid, 1ULL);
}

// ------------------------------------------------------------- K2 clusters

// Hits closer than d = k-m apart (same record) can share a k-mer window and
// must be replayed together; everything else is independent.
__global__ void pp_cluster_flag_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ val, uint64_t n_hits,
                                       const uint64_t *__restrict__ rec_begin, uint64_t n_rec, int d,
                                       uint32_t *__restrict__ hrec, uint64_t *__restrict__ hhash,
                                       uint32_t *__restrict__ cflag, const Counters *cnt)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_hits) return;
    if (i >= cnt->n_valid) { cflag[i] = 0; return; }
    const uint64_t pos = key[i];
    const uint32_t r = (uint32_t)find_rec(rec_begin, n_rec, pos);
    hrec[i] = r;
    hhash[i] = xxh64_8((uint64_t)(val[i] >> 1));
    bool start = true;
    if (i > 0) {
        const uint64_t pp = key[i - 1];
        // gap == d + 1 still couples two hits: the k-mer after the older hit's last window already sees the
        // newer one, and the rescan that fetches it applies the reference's position quirks
        start = pp < rec_begin[r] || pos - pp > (uint64_t)d + 1;
    }
    cflag[i] = start ? 1u : 0u;
}

__global__ void pp_cluster_first_kernel(const uint32_t *__restrict__ cflag, const uint32_t *__restrict__ cid, uint64_t n_hits,
                                        uint32_t *__restrict__ cl_first, Counters *cnt)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_hits || i >= cnt->n_valid) return;
    if (cflag[i]) cl_first[cid[i]] = (uint32_t)i;
    if (i + 1 == cnt->n_valid) cnt->n_clusters = cid[i] + cflag[i];
}

// --------------------------------------------------------------- K3 replay

struct Track {
    bool valid;
    uint32_t canon;
    uint64_t hash, posmin;
    bool rev;
};

// regular_minimizer_pos restricted to hits (SubSampler.cpp:81-169), window of
// k-mer c = hits [lo, hi); keeps the reference's position quirks (:88-93, :149-164).
__device__ Track pp_rescan(const uint64_t *__restrict__ key, const uint32_t *__restrict__ val,
                           const uint64_t *__restrict__ hash, uint64_t rb, uint32_t lo, uint32_t hi, uint64_t c, uint64_t d)
{
    Track t{false, 0, 0, 0, false};
    uint64_t position = 0;
    for (uint32_t i = hi; i-- > lo;) {
        const uint64_t p = key[i] - rb, h = hash[i];
        const uint32_t cn = val[i] >> 1;
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

// Sparse replay of SubSampler.cpp:352-454 over one cluster of hits.
// WRITE = false: count pieces and k-mers; WRITE = true: store the pieces.
template <bool WRITE>
__global__ void pp_replay_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ val,
                                 const uint64_t *__restrict__ hash, const uint32_t *__restrict__ hrec,
                                 const uint32_t *__restrict__ cl_first, const uint64_t *__restrict__ rec_begin,
                                 const uint64_t *__restrict__ rec_end, const uint32_t *__restrict__ rec_input, int k, int m,
                                 uint32_t *__restrict__ cl_np, uint32_t *__restrict__ cl_nk,
                                 const uint32_t *__restrict__ cl_poff, uint64_t *__restrict__ pc_first,
                                 uint32_t *__restrict__ pc_nk, uint32_t *__restrict__ pc_min, uint32_t *__restrict__ pc_meta,
                                 uint64_t max_pieces, unsigned long long *__restrict__ in_sel, const Counters *cnt)
{
    const uint64_t c_id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c_id >= cnt->n_clusters) return;
    const uint32_t i0 = cl_first[c_id];
    const uint32_t i1 = (c_id + 1 < cnt->n_clusters) ? cl_first[c_id + 1] : (uint32_t)cnt->n_valid;
    const uint32_t r = hrec[i0];
    const uint64_t rb = rec_begin[r], n = rec_end[r] - rb;
    const uint64_t d = (uint64_t)(k - m), K = n - k + 1;
    const uint32_t input = rec_input[r];
    uint32_t np = 0, nk = 0;
    uint64_t out = WRITE ? cl_poff[c_id] : 0;
    auto emit = [&](uint64_t first, uint64_t last, uint32_t mn, bool rev) {
        if (WRITE) {
            if (out < max_pieces) {
                pc_first[out] = rb + first;
                pc_nk[out] = (uint32_t)(last - first + 1);
                pc_min[out] = mn;
                pc_meta[out] = (input << 1) | (rev ? 1u : 0u);
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
        cur = pp_rescan(key, val, hash, rb, lo, hi, 0, d);
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
        if (ent && (!cur.valid || hash[hi - 1] < cur.hash)) {                // :374-388
            cur.valid = true; cur.canon = val[hi - 1] >> 1; cur.hash = hash[hi - 1]; cur.posmin = p;
            cur.rev = val[hi - 1] & 1u; is_rev = cur.rev;
        } else if (cur.valid && c - 1 >= cur.posmin) {                       // :391-398
            cur = pp_rescan(key, val, hash, rb, lo, hi, c, d);
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
    if (!WRITE) {
        cl_np[c_id] = np; cl_nk[c_id] = nk;
        if (nk) atomicAdd(in_sel + input, (unsigned long long)nk);     // header field 3: selected k-mer occurrences
    }
}

// ---------------------------------------------------------- K4 k-mer entries

// handle_superkmer (SubSampler.cpp:243-302): entry t is the (t - first entry of
// its piece)-th k-mer of the oriented piece; key as oriented, pos_min = leftmost
// occurrence of the minimizer in it.
__global__ void pp_entries_kernel(const uint32_t *__restrict__ packed, const uint64_t *__restrict__ pc_first,
                                  const uint32_t *__restrict__ pc_nk, const uint32_t *__restrict__ pc_min,
                                  const uint32_t *__restrict__ pc_meta, const uint32_t *__restrict__ pc_eoff, int k, int m,
                                  uint64_t bound, int input_shift, uint64_t *__restrict__ eA, uint64_t *__restrict__ eklo,
                                  uint64_t *__restrict__ ekhi, uint8_t *__restrict__ epm, const Counters *cnt)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bound || t >= cnt->n_entries) return;
    // piece = last one whose first entry is <= t
    uint64_t lo = 0, hi = cnt->n_pieces;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if ((uint64_t)pc_eoff[mid] <= t) lo = mid + 1; else hi = mid;
    }
    const uint64_t pc = lo - 1;
    const uint32_t j = (uint32_t)(t - pc_eoff[pc]), nk = pc_nk[pc], mn = pc_min[pc], meta = pc_meta[pc];
    const bool rev = meta & 1u;
    const uint64_t pos = pc_first[pc] + (rev ? (nk - 1 - j) : j);
    K128 key = kmer_at(packed, pos, k);
    if (rev) key = k_rc(key, k);
    const int d = k - m;
    const uint32_t mmask = (1u << (2 * m)) - 1u;
    unsigned pm = 255;
    if (k <= 32) {
        for (int q = 0; q <= d; q++)
            if (((uint32_t)(key.lo >> (2 * (d - q))) & mmask) == mn) { pm = (unsigned)q; break; }
    } else {
        for (int q = 0; q <= d; q++)
            if (((uint32_t)k_shr(key, 2 * (d - q)).lo & mmask) == mn) { pm = (unsigned)q; break; }
    }
    eA[t] = ((uint64_t)(meta >> 1) << input_shift) | mn;
    eklo[t] = key.lo;
    if (ekhi) ekhi[t] = key.hi;
    epm[t] = (uint8_t)pm;
}

// ------------------------------------------------------------ K5 unique k-mers
//
// handle_superkmer's `minimizer_map[min][kmer]` (SubSampler.cpp:243-302) as ONE global open-addressing table
// keyed by (bucket, oriented k-mer): a slot remembers the smallest entry index that carries its key (the first
// occurrence: its order is the insertion order of the reference's dense map) and how many entries do (the
// uint8 count, SubSampler.h:24).  No k-mer sort is needed: first occurrences, compacted in entry order and
// stably sorted by bucket only, ARE the buckets in insertion order.
constexpr uint32_t H_EMPTY = 0xFFFFFFFFu;

struct GHash {
    uint32_t *first;             // [slots] smallest entry index with the slot's key, H_EMPTY if free
    uint32_t *count;             // [slots] entries with the slot's key
    uint32_t *uniq;              // [slots] index of the key in the unique (bucket-grouped) arrays
    uint64_t mask;               // slots - 1
    const uint64_t *eA, *eklo, *ekhi;   // entry arrays the keys live in
};
__device__ __forceinline__ uint64_t gh_hash(uint64_t A, uint64_t lo, uint64_t hi)
{
    uint64_t x = lo ^ (hi * 0xC2B2AE3D27D4EB4FULL) ^ (A * 0x9E3779B97F4A7C15ULL);
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ULL;
    x ^= x >> 29;
    return x;
}
// Slot of (A, lo, hi), or ~0 when the key is not in the table.
__device__ __forceinline__ uint64_t gh_find(const GHash &h, uint64_t A, uint64_t lo, uint64_t hi)
{
    uint64_t s = gh_hash(A, lo, hi) & h.mask;
    for (;;) {
        const uint32_t e = h.first[s];
        if (e == H_EMPTY) return ~0ULL;
        if (h.eA[e] == A && h.eklo[e] == lo && (!h.ekhi || h.ekhi[e] == hi)) return s;
        s = (s + 1) & h.mask;
    }
}

__global__ void pp_hash_insert_kernel(GHash h, uint32_t *__restrict__ eslot, const Counters *cnt)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt->n_entries) return;
    const uint64_t A = h.eA[t], lo = h.eklo[t], hi = h.ekhi ? h.ekhi[t] : 0;
    uint64_t s = gh_hash(A, lo, hi) & h.mask;
    for (;;) {
        uint32_t e = *reinterpret_cast<volatile uint32_t *>(h.first + s);
        if (e == H_EMPTY) {
            e = atomicCAS(h.first + s, H_EMPTY, (uint32_t)t);
            if (e == H_EMPTY) break;                          // claimed a free slot
        }
        // the slot belongs to the key of entry e (entries are immutable, so that key can be read)
        if (h.eA[e] == A && h.eklo[e] == lo && (!h.ekhi || h.ekhi[e] == hi)) {
            atomicMin(h.first + s, (uint32_t)t);
            break;
        }
        s = (s + 1) & h.mask;
    }
    atomicAdd(h.count + s, 1u);
    eslot[t] = (uint32_t)s;
}

__global__ void pp_first_flag_kernel(const uint32_t *__restrict__ hfirst, const uint32_t *__restrict__ eslot, uint64_t bound,
                                     uint32_t *__restrict__ head, const Counters *cnt)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bound) return;
    head[t] = (t < cnt->n_entries && hfirst[eslot[t]] == (uint32_t)t) ? 1u : 0u;
}

// first occurrences in entry order: bucket key for the sort, entry index, identity permutation
__global__ void pp_unique_list_kernel(const uint64_t *__restrict__ eA, const uint32_t *__restrict__ head,
                                      const uint32_t *__restrict__ uid, uint64_t bound, uint64_t *__restrict__ ukey,
                                      uint32_t *__restrict__ uent, uint32_t *__restrict__ uidx, Counters *cnt)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bound) return;
    uidx[t] = (uint32_t)t;
    ukey[t] = ~0ULL;                                          // padding sorts last (real keys are written below)
    if (t + 1 == bound) cnt->n_unique = uid[t] + head[t];
}
__global__ void pp_unique_list_fill_kernel(const uint64_t *__restrict__ eA, const uint32_t *__restrict__ head,
                                           const uint32_t *__restrict__ uid, uint64_t *__restrict__ ukey,
                                           uint32_t *__restrict__ uent, const Counters *cnt)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt->n_entries || !head[t]) return;
    const uint32_t u = uid[t];
    ukey[u] = eA[t];
    uent[u] = (uint32_t)t;
}

// unique k-mers grouped by bucket (stable sort of the list above): key, leftmost minimizer position, count
// mod 256, bucket heads; every slot learns where its key ended up.
__global__ void pp_unique_finish_kernel(const uint64_t *__restrict__ skey, const uint32_t *__restrict__ order,
                                        const uint32_t *__restrict__ uent, const uint32_t *__restrict__ eslot, GHash h,
                                        const uint8_t *__restrict__ epm, uint64_t bound, uint64_t *__restrict__ uA,
                                        uint64_t *__restrict__ uklo, uint64_t *__restrict__ ukhi, uint8_t *__restrict__ upm,
                                        uint8_t *__restrict__ ucnt, uint32_t *__restrict__ bflag, uint8_t *__restrict__ seen,
                                        const Counters *cnt)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= bound) return;
    seen[j] = 0;
    if (j >= cnt->n_unique) { bflag[j] = 0; return; }
    const uint32_t t = uent[order[j]], s = eslot[t];
    uA[j] = skey[j];
    uklo[j] = h.eklo[t];
    if (ukhi) ukhi[j] = h.ekhi[t];
    upm[j] = epm[t];
    ucnt[j] = (uint8_t)(h.count[s] & 0xFFu);                  // uint8 counter of the reference wraps at 256
    h.uniq[s] = (uint32_t)j;
    bflag[j] = (j == 0 || skey[j] != skey[j - 1]) ? 1u : 0u;
}

__global__ void pp_bucket_start_kernel(const uint32_t *__restrict__ bflag, const uint32_t *__restrict__ bid, uint64_t bound,
                                       uint32_t *__restrict__ bstart, Counters *cnt)
{
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= bound || u >= cnt->n_unique) return;
    if (bflag[u]) bstart[bid[u]] = (uint32_t)u;               // bid = exclusive scan of the head flags
    if (u + 1 == cnt->n_unique) cnt->n_buckets = bid[u] + bflag[u];
}

// -------------------------------------------------------- K6 reconstruction

// One bucket's unique k-mers in insertion order; indices are bucket-relative.  The arrays live in shared
// memory (small buckets, staged by the warp) or in global memory (buckets larger than RC_CAP, looked up
// through the global table).
struct BucketView {
    const uint64_t *klo, *khi;     // khi null when k <= 32
    const uint8_t *cnt;
    uint8_t *seen;
    uint32_t n, bs;                // size, first unique index of the bucket
    uint64_t A;                    // bucket key
    unsigned abundance;
    int k;
};

__device__ __forceinline__ K128 bv_key(const BucketView &v, uint32_t u)
{
    return K128{v.klo[u], v.khi ? v.khi[u] : 0};
}
// bucket-relative index of `key` in bucket v, or -1
__device__ __forceinline__ int bv_find(const GHash &h, const BucketView &v, const K128 &key)
{
    const uint64_t s = gh_find(h, v.A, key.lo, key.hi);
    return s == ~0ULL ? -1 : (int)(h.uniq[s] - v.bs);
}
__device__ __forceinline__ K128 bv_neighbour(const K128 &cur, bool left, int t, int k)
{
    const uint64_t o = (0x3120u >> (4 * t)) & 3u;                          // A, T, C, G (SubSampler.cpp:568)
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
constexpr int RC_CAP = 256;          // bucket size staged in shared memory
constexpr int RC_SLOTS = 512;        // open-addressing table over the staged bucket (load <= 0.5)
constexpr int RC_WARPS = 4;          // buckets in flight per CTA (one warp each)
constexpr int RC_SK = 192;           // 2-bit codes of a super-k-mer (2k-m <= 123, grows both ways from 64)
constexpr uint32_t VISIT_START = 0u << 30, VISIT_LEFT = 1u << 30, VISIT_RIGHT = 2u << 30, VISIT_MASK = (1u << 30) - 1u;
constexpr uint32_t VISIT_END = 0xFFFFFFFFu;
constexpr uint16_t ADJ_NONE = RC_CAP;    // sentinel neighbour: seen[RC_CAP] is always set

// Shared-memory slice of one warp of pp_chain_kernel.
struct RcSmem {
    uint64_t *klo, *khi;
    uint32_t *slot;
    uint16_t *adj;               // [u][left 0..3, right 0..3] neighbour index in probe order A,T,C,G
    uint8_t *cnt, *seen, *pm;    // seen has RC_CAP + 16 entries
};
static size_t rc_smem_bytes(bool hi128)
{
    return (size_t)RC_WARPS * (RC_CAP * (8 + (hi128 ? 8 : 0) + 16 + 3) + 16 + RC_SLOTS * 4);
}
__device__ __forceinline__ RcSmem rc_carve(uint8_t *base, int wi, bool hi128)
{
    RcSmem r;
    r.klo = reinterpret_cast<uint64_t *>(base) + (size_t)wi * RC_CAP;
    base += (size_t)RC_WARPS * RC_CAP * 8;
    r.khi = nullptr;
    if (hi128) { r.khi = reinterpret_cast<uint64_t *>(base) + (size_t)wi * RC_CAP; base += (size_t)RC_WARPS * RC_CAP * 8; }
    r.adj = reinterpret_cast<uint16_t *>(base) + (size_t)wi * RC_CAP * 8;
    base += (size_t)RC_WARPS * RC_CAP * 16;
    r.slot = reinterpret_cast<uint32_t *>(base) + (size_t)wi * RC_SLOTS;
    base += (size_t)RC_WARPS * RC_SLOTS * 4;
    r.cnt = base + (size_t)wi * RC_CAP;
    r.pm = base + (size_t)(RC_WARPS + wi) * RC_CAP;
    r.seen = base + (size_t)2 * RC_WARPS * RC_CAP + (size_t)wi * (RC_CAP + 16);
    return r;
}
__device__ __forceinline__ uint32_t rc_hash(const K128 &key)
{
    return (uint32_t)(((key.lo ^ (key.hi * 0xC2B2AE3D27D4EB4FULL)) * 0x9E3779B97F4A7C15ULL) >> (64 - 9));
}
static_assert(RC_SLOTS == 512, "rc_hash yields 9 bits");

// Index of `key` in the staged bucket, or -1.
__device__ __forceinline__ int rc_lookup(const RcSmem &sm, const K128 &key)
{
    uint32_t s = rc_hash(key);
    for (;;) {
        const uint32_t idx = sm.slot[s];
        if (idx == 0xFFFFFFFFu) return -1;
        if (sm.klo[idx] == key.lo && (!sm.khi || sm.khi[idx] == key.hi)) return (int)idx;
        s = (s + 1) & (RC_SLOTS - 1);
    }
}

// find_first_kmer + reconstruct_superkmer + find_next (SubSampler.cpp:512-620) on a
// bucket staged in shared memory.  The four neighbour look-ups of every k-mer in
// both directions are done up front by the whole warp (hash table); the greedy
// chain itself -- inherently sequential -- then only follows indices and `seen`
// flags on lane 0, which records the visit order with each k-mer's role.
__device__ __forceinline__ void rc_walk_staged(const RcSmem &sm, uint32_t nb, int k, int m, unsigned abundance,
                                               uint32_t *__restrict__ visit)
{
    const int lane = threadIdx.x & 31;
    const int d = k - m;
    const bool hi128 = sm.khi != nullptr;
    for (uint32_t idx = lane; idx < nb * 8; idx += 32) {
        const uint32_t u = idx >> 3, t = idx & 3;
        const bool left = !(idx & 4);
        const K128 cand = bv_neighbour(K128{sm.klo[u], hi128 ? sm.khi[u] : 0}, left, (int)t, k);
        const int w = rc_lookup(sm, cand);
        sm.adj[idx] = (w >= 0 && sm.cnt[w] >= abundance) ? (uint16_t)w : ADJ_NONE;
    }
    __syncwarp();
    if (lane == 0) {
        uint32_t nv = 0;
        // find_first_kmer (:604-620): the k-mers are stored in insertion order
        for (uint32_t start = 0; start < nb; start++) {
            if (sm.seen[start] || sm.cnt[start] < abundance) continue;
            sm.seen[start] = 1;
            visit[nv++] = start | VISIT_START;
            const uint32_t pms = sm.pm[start];
            // n_left = d - pos_min as uint64 in the reference: a k-mer without the minimizer text (pos 255)
            // extends to the left for as long as it can
            uint32_t n_left = pms == 255u ? 0x7fffffffu : (uint32_t)d - pms, n_right = pms;
            uint32_t cur = start, ext = 0;                   // ext = k-mers added to the start k-mer
            while (ext != (uint32_t)d) {                     // |sk| != 2k-m
                const bool left = n_left != 0;
                if (!left && n_right == 0) break;
                // find_next (:566-602): first neighbour in probe order that is in the bucket and unseen
                const uint2 pk = *reinterpret_cast<const uint2 *>(sm.adj + cur * 8 + (left ? 0 : 4));
                const uint32_t c0 = pk.x & 0xFFFFu, c1 = pk.x >> 16, c2 = pk.y & 0xFFFFu, c3 = pk.y >> 16;
                const uint32_t s0 = sm.seen[c0], s1 = sm.seen[c1], s2 = sm.seen[c2], s3 = sm.seen[c3];
                const uint32_t found = !s0 ? c0 : !s1 ? c1 : !s2 ? c2 : !s3 ? c3 : (uint32_t)ADJ_NONE;
                const bool ok = found != ADJ_NONE;
                sm.seen[found] = 1;                          // the sentinel's flag is set anyway
                if (ok) { visit[nv++] = found | (left ? VISIT_LEFT : VISIT_RIGHT); ext++; }
                if (left) {
                    n_left = ok ? n_left - 1 : 0;
                    cur = n_left == 0 ? start : found;
                } else {
                    if (!ok) break;
                    n_right--;
                    cur = found;
                }
            }
        }
        if (nv < nb) visit[nv] = VISIT_END;                  // k-mers below the abundance are never visited
    }
    __syncwarp();
}

// Same walk for a bucket too large to stage: global memory, binary search, lanes
// 0..3 probing the four neighbours at once.
__device__ __forceinline__ int rc_step_global(const GHash &h, const BucketView &v, const K128 &cur, bool left, K128 *out)
{
    const int lane = threadIdx.x & 31;
    int u = -1;
    if (lane < 4) {
        u = bv_find(h, v, bv_neighbour(cur, left, lane, v.k));
        if (u >= 0 && (v.seen[u] || v.cnt[u] < v.abundance)) u = -1;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, u >= 0);
    if (!bal) return -1;
    const int t = __ffs(bal) - 1;
    u = __shfl_sync(0xffffffffu, u, t);
    if (lane == 0) v.seen[u] = 1;
    __syncwarp();
    *out = bv_neighbour(cur, left, t, v.k);
    return u;
}
__device__ __forceinline__ void rc_walk_global(const GHash &h, const BucketView &v, const uint8_t *__restrict__ pm, int k,
                                               int m, uint32_t *__restrict__ visit)
{
    const int lane = threadIdx.x & 31;
    const int d = k - m;
    const uint32_t nb = v.n;
    uint32_t nv = 0;
    for (uint32_t start = 0; start < nb; start++) {
        if (v.seen[start] || v.cnt[start] < v.abundance) continue;
        __syncwarp();
        if (lane == 0) { v.seen[start] = 1; visit[nv] = start | VISIT_START; }
        __syncwarp();
        nv++;
        const K128 skey = bv_key(v, start);
        uint64_t n_left = (uint64_t)d - pm[start], n_right = pm[start];
        K128 cur = skey;
        uint32_t ext = 0;
        while (ext != (uint32_t)d) {
            if (n_left != 0) {
                K128 nx;
                const int u = rc_step_global(h, v, cur, true, &nx);
                n_left--;
                if (u >= 0) {
                    if (lane == 0) visit[nv] = (uint32_t)u | VISIT_LEFT;
                    nv++; ext++;
                } else {
                    n_left = 0;
                }
                cur = (n_left == 0) ? skey : nx;
            } else if (n_right != 0) {
                K128 nx;
                const int u = rc_step_global(h, v, cur, false, &nx);
                n_right--;
                if (u < 0) break;
                if (lane == 0) visit[nv] = (uint32_t)u | VISIT_RIGHT;
                nv++; ext++;
                cur = nx;
            } else {
                break;
            }
        }
    }
    if (lane == 0 && nv < nb) visit[nv] = VISIT_END;
}

// One bucket per warp: the visit order of its k-mers (start / left extension /
// right extension of each super-k-mer), everything pp_emit_kernel needs to size
// and write the bytes without another look-up.
__global__ void __launch_bounds__(RC_WARPS * 32)
pp_chain_kernel(const uint64_t *__restrict__ uA, const uint64_t *__restrict__ uklo, const uint64_t *__restrict__ ukhi,
                const uint8_t *__restrict__ ucnt, const uint8_t *__restrict__ upm, uint8_t *__restrict__ seen_g, GHash gh,
                const uint32_t *__restrict__ bstart, int k, int m, unsigned abundance, uint32_t *__restrict__ visit,
                Counters *cnt)
{
    extern __shared__ __align__(16) uint8_t rc_smem[];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const RcSmem sm = rc_carve(rc_smem, wi, ukhi != nullptr);
    const uint64_t n_buckets = cnt->n_buckets;
    // persistent warps pull buckets from a queue: the makespan is the largest bucket, not a wave of them
    for (;;) {
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(&cnt->chain_next, 1ULL);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= n_buckets) break;
        const uint32_t bs = bstart[b], be = (b + 1 < n_buckets) ? bstart[b + 1] : (uint32_t)cnt->n_unique;
        const uint32_t nb = be - bs;
        __syncwarp();
        if (nb <= (uint32_t)RC_CAP) {
            for (int i = lane; i < RC_SLOTS; i += 32) sm.slot[i] = 0xFFFFFFFFu;
            for (uint32_t i = lane; i < nb; i += 32) {
                sm.klo[i] = uklo[bs + i];
                if (ukhi) sm.khi[i] = ukhi[bs + i];
                sm.cnt[i] = ucnt[bs + i];
                sm.pm[i] = upm[bs + i];
                sm.seen[i] = 0;
            }
            if (lane == 0) sm.seen[RC_CAP] = 1;                                 // the "no neighbour" sentinel
            __syncwarp();
            for (uint32_t i = lane; i < nb; i += 32) {                          // keys of a bucket are distinct
                uint32_t s = rc_hash(K128{sm.klo[i], ukhi ? sm.khi[i] : 0});
                while (atomicCAS(&sm.slot[s], 0xFFFFFFFFu, i) != 0xFFFFFFFFu) s = (s + 1) & (RC_SLOTS - 1);
            }
            __syncwarp();
            rc_walk_staged(sm, nb, k, m, abundance, visit + bs);
        } else {
            BucketView v{uklo + bs, ukhi ? ukhi + bs : nullptr, ucnt + bs, seen_g + bs, nb, bs, uA[bs], abundance, k};
            rc_walk_global(gh, v, upm + bs, k, m, visit + bs);
        }
    }
}

// Leftmost occurrence of the minimizer in the super-k-mer sk[lo, lo+len), -1 if none (warp-parallel find()).
__device__ __forceinline__ int rc_find_minimizer(const uint8_t *sk, int lo, int len, int m, uint32_t minimizer)
{
    const int lane = threadIdx.x & 31;
    int best = 0x7fffffff;
    for (int t = lane; t + m <= len; t += 32) {
        uint32_t w = 0;
        for (int i = 0; i < m; i++) w = (w << 2) | sk[lo + t + i];
        if (w == minimizer) { best = t; break; }
    }
    best = __reduce_min_sync(0xffffffffu, best);
    return best == 0x7fffffff ? -1 : best;
}

// Writer loop (SubSampler.cpp:459-504), one bucket per warp, from the visit order:
// minimizer text, u32 size, packed maximal super-k-mers (prefix(d) + suffix(d), 4
// bases per byte), the others as "prefix\nsuffix\n" text split at the leftmost
// minimizer occurrence, "\n\n".  WRITE = false only measures (bytes per bucket and
// per input), WRITE = true emits at the offsets computed from that.
template <bool WRITE>
__global__ void __launch_bounds__(8 * 32)
pp_emit_kernel(const uint64_t *__restrict__ uA, const uint64_t *__restrict__ uklo, const uint64_t *__restrict__ ukhi,
               const uint32_t *__restrict__ visit, const uint32_t *__restrict__ bstart, int k, int m, int input_shift,
               uint32_t *__restrict__ bbytes, uint32_t *__restrict__ bnmax, const uint64_t *__restrict__ boff,
               uint8_t *__restrict__ body, unsigned long long *__restrict__ in_bytes, const Counters *cnt)
{
    __shared__ uint8_t s_sk[8][RC_SK];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    uint8_t *sk = s_sk[wi];
    const uint64_t n_buckets = cnt->n_buckets;
    const int d = k - m, full = 2 * k - m;
    for (uint64_t b = (uint64_t)blockIdx.x * 8 + wi; b < n_buckets; b += (uint64_t)gridDim.x * 8) {
        const uint32_t bs = bstart[b], be = (b + 1 < n_buckets) ? bstart[b + 1] : (uint32_t)cnt->n_unique;
        const uint32_t nb = be - bs;
        const uint32_t minimizer = (uint32_t)(uA[bs] & ((1ULL << input_shift) - 1));
        uint8_t *p_max = nullptr, *p_txt = nullptr;
        if (WRITE) {
            uint8_t *o = body + boff[b];
            const uint32_t nm = bnmax[b];
            const uint32_t sz = nm ? 1 + nm * (uint32_t)(2 * d) / 4 : 0;        // strCompressor: mod byte + 4 bases/byte
            if (lane == 0) {
                for (int i = 0; i < m; i++) o[i] = "ACTG"[(minimizer >> (2 * (m - 1 - i))) & 3];
                o[m] = (uint8_t)sz; o[m + 1] = (uint8_t)(sz >> 8); o[m + 2] = (uint8_t)(sz >> 16); o[m + 3] = (uint8_t)(sz >> 24);
                if (sz) o[m + 4] = 0;                                           // 2d is a multiple of 4: mod byte 0
            }
            p_max = o + m + 4 + (sz ? 1 : 0);
            p_txt = o + m + 4 + sz;
        }
        uint32_t n_max = 0, text_len = 0;
        uint32_t j = 0;
        while (j < nb) {
            const uint32_t head = visit[bs + j];
            if (head == VISIT_END) break;
            // one super-k-mer: its start entry and the extensions that follow it (lefts first, then rights)
            uint32_t e = j + 1, n_l = 0;
            while (e < nb) {
                const uint32_t x = visit[bs + e];
                if (x == VISIT_END || (x >> 30) == 0) break;
                n_l += (x >> 30) == 1;
                e++;
            }
            const int lo = 64 - (int)n_l, len = k + (int)(e - j - 1);
            if (!WRITE && len == full) { n_max++; j = e; continue; }           // :479-485, size known without the string
            const K128 skey{uklo[bs + (head & VISIT_MASK)], ukhi ? ukhi[bs + (head & VISIT_MASK)] : 0};
            __syncwarp();
            for (int i = lane; i < k; i += 32) sk[64 + i] = (uint8_t)(k_shr(skey, 2 * (k - 1 - i)).lo & 3);
            for (uint32_t t = j + 1 + lane; t < e; t += 32) {
                const uint32_t x = visit[bs + t], u = x & VISIT_MASK;
                const K128 key{uklo[bs + u], ukhi ? ukhi[bs + u] : 0};
                const uint32_t r = t - (j + 1);
                if ((x >> 30) == 1) sk[64 - 1 - r] = (uint8_t)(k_shr(key, 2 * k - 2).lo & 3);     // r-th left extension
                else sk[64 + k + (r - n_l)] = (uint8_t)(key.lo & 3);                                // right extension
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
                const int q = rc_find_minimizer(sk, lo, len, m, minimizer);
                if (WRITE) {
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
                text_len += (uint32_t)(q < 0 ? len : len - m) + 2;
            }
            j = e;
        }
        if (WRITE) {
            if (lane == 0) { p_txt[0] = '\n'; p_txt[1] = '\n'; }
        } else if (lane == 0) {
            const uint32_t sz = n_max ? 1 + n_max * (uint32_t)(2 * d) / 4 : 0;
            const uint32_t bytes = (uint32_t)m + 4 + sz + text_len + 2;
            bbytes[b] = bytes;
            bnmax[b] = n_max;
            atomicAdd(in_bytes + (uA[bs] >> input_shift), (unsigned long long)bytes);
        }
    }
}

// ------------------------------------------------------ K7 compare elements

// What the comparator decodes from the sketch (Comparator.cpp:97-264): the
// distinct canonical k-mers of every bucket.  A unique oriented k-mer is in the
// sketch iff count >= abundance; two orientations of one k-mer collapse.
__global__ void pp_element_flag_kernel(const uint64_t *__restrict__ uA, const uint64_t *__restrict__ uklo,
                                       const uint64_t *__restrict__ ukhi, const uint8_t *__restrict__ ucnt, GHash gh, int k,
                                       unsigned abundance, uint64_t bound, uint32_t *__restrict__ eflag,
                                       const Counters *cnt)
{
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= bound) return;
    bool keep = false;
    if (u < cnt->n_unique && ucnt[u] >= abundance) {
        keep = true;
        const K128 key{uklo[u], ukhi ? ukhi[u] : 0};
        const K128 rc = k_rc(key, k);
        if (k_lt(rc, key)) {
            // non-canonical orientation: drop it if the canonical one is in the bucket too
            const uint64_t s = gh_find(gh, uA[u], rc.lo, rc.hi);
            if (s != ~0ULL && ucnt[gh.uniq[s]] >= abundance) keep = false;
        }
    }
    eflag[u] = keep ? 1u : 0u;
}

__global__ void pp_element_write_kernel(const uint64_t *__restrict__ uA, const uint64_t *__restrict__ uklo,
                                        const uint64_t *__restrict__ ukhi, const uint32_t *__restrict__ eflag,
                                        const uint32_t *__restrict__ eoff, int k, int input_shift, uint64_t bound,
                                        uint32_t *__restrict__ el_min, uint64_t *__restrict__ el_klo,
                                        uint64_t *__restrict__ el_khi, unsigned long long *__restrict__ in_first,
                                        Counters *cnt)
{
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= bound) return;
    // unique k-mers are sorted by input: the first one of an input marks where its elements begin
    if (u < cnt->n_unique && (u == 0 || (uA[u] >> input_shift) != (uA[u - 1] >> input_shift)))
        in_first[uA[u] >> input_shift] = eoff[u];
    if (u < cnt->n_unique && eflag[u]) {
        K128 key{uklo[u], ukhi ? ukhi[u] : 0};
        const K128 rc = k_rc(key, k);
        if (k_lt(rc, key)) key = rc;
        const uint32_t o = eoff[u];
        el_min[o] = (uint32_t)(uA[u] & ((1ULL << input_shift) - 1));
        el_klo[o] = key.lo;
        if (el_khi) el_khi[o] = key.hi;
    }
    if (u + 1 == bound) cnt->n_elems = eoff[u] + eflag[u];
}

__global__ void pp_totals_kernel(const uint32_t *__restrict__ cl_np, const uint32_t *__restrict__ cl_nk,
                                 const uint32_t *__restrict__ cl_poff, const uint32_t *__restrict__ cl_eoff, Counters *cnt)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const unsigned long long nc = cnt->n_clusters;
        cnt->n_pieces = nc ? (unsigned long long)cl_poff[nc - 1] + cl_np[nc - 1] : 0;
        cnt->n_entries = nc ? (unsigned long long)cl_eoff[nc - 1] + cl_nk[nc - 1] : 0;
    }
}

__global__ void pp_body_total_kernel(const uint64_t *__restrict__ boff, const uint32_t *__restrict__ bbytes, Counters *cnt)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const unsigned long long nb = cnt->n_buckets;
        cnt->body_bytes = nb ? boff[nb - 1] + bbytes[nb - 1] : 0;
    }
}

__global__ void pp_piece_offsets_kernel(const uint32_t *__restrict__ pc_nk, uint64_t bound, uint32_t *__restrict__ tmp,
                                        const Counters *cnt)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < bound) tmp[t] = t < cnt->n_pieces ? pc_nk[t] : 0;
}

// ---------------------------------------------------------------- driver

#define PP_CK(call)                                  \
    do {                                             \
        cudaError_t e_ = (call);                     \
        if (e_ != cudaSuccess) return e_;            \
    } while (0)

static inline unsigned nblk(uint64_t n, unsigned t = 256) { return (unsigned)((n + t - 1) / t ? (n + t - 1) / t : 1); }

template <class K, class V>
static cudaError_t sort_pairs(PostpassBuffers *b, const K *kin, K *kout, const V *vin, V *vout, uint64_t n, int begin_bit,
                              int end_bit, cudaStream_t st)
{
    size_t bytes = 0;
    PP_CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, begin_bit, end_bit, st));
    PP_CK(b->cubtmp.ensure(bytes));
    return cub::DeviceRadixSort::SortPairs(b->cubtmp.p, bytes, kin, kout, vin, vout, (int)n, begin_bit, end_bit, st);
}
template <class T, class O>
static cudaError_t excl_sum(PostpassBuffers *b, const T *in, O *out, uint64_t n, cudaStream_t st)
{
    size_t bytes = 0;
    PP_CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, st));
    PP_CK(b->cubtmp.ensure(bytes));
    return cub::DeviceScan::ExclusiveSum(b->cubtmp.p, bytes, in, out, (int)n, st);
}

static int bits_for(uint64_t v)
{
    int b = 0;
    while (b < 64 && (v >> b)) b++;
    return b ? b : 1;
}

cudaError_t postpass_run(PostpassBuffers *b, const PostpassIn &in, PostpassOut *out, cudaStream_t st)
{
    const int k = in.k, m = in.m, d = k - m;
    const bool hi128 = k > 32;
    const uint64_t nh = in.n_hits;
    uint32_t launched = 0;
    // every k-mer is in at most one piece: entries <= min(hits * (d+1), bases)
    uint64_t bound = nh * (uint64_t)(d + 1);
    if (bound > in.n_bases) bound = in.n_bases;
    if (bound < 1) bound = 1;
    if (bound >= (1ULL << 31)) return cudaErrorInvalidValue;        // caller falls back to the host post-pass
    const uint64_t nhb = nh ? nh : 1;
    const int input_shift = 2 * m;                                   // bucket id = input << 2m | minimizer
    const int a_bits = input_shift + bits_for(in.n_inputs ? in.n_inputs - 1 : 0) + 1;   // +1: padding entries sort last

    PP_CK(b->cnt.ensure(sizeof(Counters)));
    Counters *cnt = b->cnt.as<Counters>();
    PP_CK(cudaMemsetAsync(cnt, 0, sizeof(Counters), st));
    const size_t nin = in.n_inputs ? in.n_inputs : 1;
    PP_CK(b->in_bytes.ensure(nin * 8)); PP_CK(b->in_sel.ensure(nin * 8)); PP_CK(b->in_elems.ensure(nin * 8));
    PP_CK(cudaMemsetAsync(b->in_bytes.p, 0, nin * 8, st));
    PP_CK(cudaMemsetAsync(b->in_sel.p, 0, nin * 8, st));
    PP_CK(cudaMemsetAsync(b->in_elems.p, 0xFF, nin * 8, st));       // first element of each input; ~0 = none

    // ---- hits: classify, sort by position, clusters
    PP_CK(b->hkey.ensure(nhb * 8)); PP_CK(b->hval.ensure(nhb * 4)); PP_CK(b->hkey2.ensure(nhb * 8)); PP_CK(b->hval2.ensure(nhb * 4));
    PP_CK(b->hhash.ensure(nhb * 8)); PP_CK(b->hrec.ensure(nhb * 4)); PP_CK(b->cflag.ensure(nhb * 4)); PP_CK(b->cid.ensure(nhb * 4));
    PP_CK(b->cl_first.ensure(nhb * 4)); PP_CK(b->cl_np.ensure(nhb * 4)); PP_CK(b->cl_nk.ensure(nhb * 4));
    PP_CK(b->cl_poff.ensure(nhb * 4)); PP_CK(b->cl_eoff.ensure(nhb * 4));
    if (nh) {
        const int pos_bits = bits_for(in.n_bases);                   // radix passes only over bits a position can have
        pp_classify_kernel<<<nblk(nh), 256, 0, st>>>(in.d_hits, nh, in.d_rec_begin, in.d_rec_end, in.n_rec, k, m,
                                                     1ULL << pos_bits, b->hkey.as<uint64_t>(), b->hval.as<uint32_t>(), cnt);
        launched++;
        PP_CK(sort_pairs(b, b->hkey.as<uint64_t>(), b->hkey2.as<uint64_t>(), b->hval.as<uint32_t>(), b->hval2.as<uint32_t>(),
                         nh, 0, pos_bits + 1, st));
        const uint64_t *key = b->hkey2.as<uint64_t>();
        const uint32_t *val = b->hval2.as<uint32_t>();
        pp_cluster_flag_kernel<<<nblk(nh), 256, 0, st>>>(key, val, nh, in.d_rec_begin, in.n_rec, d, b->hrec.as<uint32_t>(),
                                                         b->hhash.as<uint64_t>(), b->cflag.as<uint32_t>(), cnt);
        PP_CK(excl_sum(b, b->cflag.as<uint32_t>(), b->cid.as<uint32_t>(), nh, st));
        pp_cluster_first_kernel<<<nblk(nh), 256, 0, st>>>(b->cflag.as<uint32_t>(), b->cid.as<uint32_t>(), nh,
                                                          b->cl_first.as<uint32_t>(), cnt);
        launched += 2;
        // ---- replay: count, offsets, write
        PP_CK(cudaMemsetAsync(b->cl_np.p, 0, nhb * 4, st));
        PP_CK(cudaMemsetAsync(b->cl_nk.p, 0, nhb * 4, st));
        PP_CK(b->pc_first.ensure(bound * 8)); PP_CK(b->pc_nk.ensure(bound * 4)); PP_CK(b->pc_min.ensure(bound * 4));
        PP_CK(b->pc_meta.ensure(bound * 4)); PP_CK(b->pc_eoff.ensure(bound * 4));
        pp_replay_kernel<false><<<nblk(nh, 128), 128, 0, st>>>(key, val, b->hhash.as<uint64_t>(), b->hrec.as<uint32_t>(),
            b->cl_first.as<uint32_t>(), in.d_rec_begin, in.d_rec_end, in.d_rec_input, k, m, b->cl_np.as<uint32_t>(),
            b->cl_nk.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr, nullptr, bound, b->in_sel.as<unsigned long long>(), cnt);
        PP_CK(excl_sum(b, b->cl_np.as<uint32_t>(), b->cl_poff.as<uint32_t>(), nh, st));
        PP_CK(excl_sum(b, b->cl_nk.as<uint32_t>(), b->cl_eoff.as<uint32_t>(), nh, st));
        pp_totals_kernel<<<1, 32, 0, st>>>(b->cl_np.as<uint32_t>(), b->cl_nk.as<uint32_t>(), b->cl_poff.as<uint32_t>(),
                                           b->cl_eoff.as<uint32_t>(), cnt);
        pp_replay_kernel<true><<<nblk(nh, 128), 128, 0, st>>>(key, val, b->hhash.as<uint64_t>(), b->hrec.as<uint32_t>(),
            b->cl_first.as<uint32_t>(), in.d_rec_begin, in.d_rec_end, in.d_rec_input, k, m, b->cl_np.as<uint32_t>(),
            b->cl_nk.as<uint32_t>(), b->cl_poff.as<uint32_t>(), b->pc_first.as<uint64_t>(), b->pc_nk.as<uint32_t>(),
            b->pc_min.as<uint32_t>(), b->pc_meta.as<uint32_t>(), bound, nullptr, cnt);
        launched += 3;
    } else {
        PP_CK(b->pc_first.ensure(8)); PP_CK(b->pc_nk.ensure(4)); PP_CK(b->pc_min.ensure(4)); PP_CK(b->pc_meta.ensure(4));
        PP_CK(b->pc_eoff.ensure(4));
    }
    // ---- entries
    PP_CK(b->eA.ensure(bound * 8)); PP_CK(b->eklo.ensure(bound * 8)); if (hi128) PP_CK(b->ekhi.ensure(bound * 8));
    PP_CK(b->epm.ensure(bound));
    PP_CK(b->skey.ensure(bound * 8)); PP_CK(b->skey2.ensure(bound * 8)); PP_CK(b->head.ensure(bound * 4)); PP_CK(b->uid.ensure(bound * 4));
    {
        // entry offset of every piece (exclusive scan of piece sizes)
        uint32_t *tmp = b->head.as<uint32_t>();
        pp_piece_offsets_kernel<<<nblk(bound), 256, 0, st>>>(b->pc_nk.as<uint32_t>(), bound, tmp, cnt);
        PP_CK(excl_sum(b, tmp, b->pc_eoff.as<uint32_t>(), bound, st));
        launched++;
    }
    uint64_t *ekhi = hi128 ? b->ekhi.as<uint64_t>() : nullptr;
    pp_entries_kernel<<<nblk(bound), 256, 0, st>>>(in.d_packed, b->pc_first.as<uint64_t>(), b->pc_nk.as<uint32_t>(),
        b->pc_min.as<uint32_t>(), b->pc_meta.as<uint32_t>(), b->pc_eoff.as<uint32_t>(), k, m, bound, input_shift,
        b->eA.as<uint64_t>(), b->eklo.as<uint64_t>(), ekhi, b->epm.as<uint8_t>(), cnt);
    launched++;
    // ---- unique k-mers: global hash table keyed by (bucket, k-mer), then ONE stable sort by bucket
    uint64_t slots = 1024;
    while (slots < 2 * bound) slots <<= 1;
    if (slots > (1ULL << 31)) return cudaErrorInvalidValue;
    PP_CK(b->hfirst.ensure(slots * 4)); PP_CK(b->hcount.ensure(slots * 4)); PP_CK(b->huniq.ensure(slots * 4));
    PP_CK(b->eslot.ensure(bound * 4));
    PP_CK(cudaMemsetAsync(b->hfirst.p, 0xFF, slots * 4, st));
    PP_CK(cudaMemsetAsync(b->hcount.p, 0, slots * 4, st));
    GHash gh{b->hfirst.as<uint32_t>(), b->hcount.as<uint32_t>(), b->huniq.as<uint32_t>(), slots - 1, b->eA.as<uint64_t>(),
             b->eklo.as<uint64_t>(), ekhi};
    pp_hash_insert_kernel<<<nblk(bound), 256, 0, st>>>(gh, b->eslot.as<uint32_t>(), cnt);
    pp_first_flag_kernel<<<nblk(bound), 256, 0, st>>>(b->hfirst.as<uint32_t>(), b->eslot.as<uint32_t>(), bound,
                                                      b->head.as<uint32_t>(), cnt);
    PP_CK(excl_sum(b, b->head.as<uint32_t>(), b->uid.as<uint32_t>(), bound, st));
    PP_CK(b->uent.ensure(bound * 4)); PP_CK(b->uidx0.ensure(bound * 4)); PP_CK(b->uidx2.ensure(bound * 4));
    pp_unique_list_kernel<<<nblk(bound), 256, 0, st>>>(b->eA.as<uint64_t>(), b->head.as<uint32_t>(), b->uid.as<uint32_t>(), bound,
        b->skey.as<uint64_t>(), b->uent.as<uint32_t>(), b->uidx0.as<uint32_t>(), cnt);
    pp_unique_list_fill_kernel<<<nblk(bound), 256, 0, st>>>(b->eA.as<uint64_t>(), b->head.as<uint32_t>(), b->uid.as<uint32_t>(),
        b->skey.as<uint64_t>(), b->uent.as<uint32_t>(), cnt);
    // first occurrences are in entry order: a stable sort by bucket leaves every bucket in insertion order
    PP_CK(sort_pairs(b, b->skey.as<uint64_t>(), b->skey2.as<uint64_t>(), b->uidx0.as<uint32_t>(), b->uidx2.as<uint32_t>(), bound, 0,
                     a_bits, st));
    PP_CK(b->uA.ensure(bound * 8)); PP_CK(b->uklo.ensure(bound * 8)); if (hi128) PP_CK(b->ukhi.ensure(bound * 8));
    PP_CK(b->upm.ensure(bound)); PP_CK(b->ucnt.ensure(bound));
    PP_CK(b->bflag.ensure(bound * 4)); PP_CK(b->bidm.ensure(bound * 4)); PP_CK(b->bstart.ensure(bound * 4));
    PP_CK(b->seen.ensure(bound));
    uint64_t *ukhi = hi128 ? b->ukhi.as<uint64_t>() : nullptr;
    pp_unique_finish_kernel<<<nblk(bound), 256, 0, st>>>(b->skey2.as<uint64_t>(), b->uidx2.as<uint32_t>(), b->uent.as<uint32_t>(),
        b->eslot.as<uint32_t>(), gh, b->epm.as<uint8_t>(), bound, b->uA.as<uint64_t>(), b->uklo.as<uint64_t>(), ukhi,
        b->upm.as<uint8_t>(), b->ucnt.as<uint8_t>(), b->bflag.as<uint32_t>(), b->seen.as<uint8_t>(), cnt);
    PP_CK(excl_sum(b, b->bflag.as<uint32_t>(), b->bidm.as<uint32_t>(), bound, st));
    pp_bucket_start_kernel<<<nblk(bound), 256, 0, st>>>(b->bflag.as<uint32_t>(), b->bidm.as<uint32_t>(), bound,
                                                        b->bstart.as<uint32_t>(), cnt);
    launched += 6;
    // ---- reconstruction: walk every bucket's chains once (visit order + byte sizes), offsets; bytes are emitted below
    const size_t rc_smem = rc_smem_bytes(hi128);
    {
        static PerDeviceOnce once;
        PP_CK(once.run([] {
            return cudaFuncSetAttribute(pp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rc_smem_bytes(true));
        }));
    }
    const unsigned rc_grid = (unsigned)std::min<uint64_t>((bound + RC_WARPS - 1) / RC_WARPS, 148 * 8);
    PP_CK(b->bbytes.ensure(bound * 4)); PP_CK(b->bnmax.ensure(bound * 4)); PP_CK(b->boff.ensure(bound * 8));
    PP_CK(b->visit.ensure(bound * 4));
    PP_CK(cudaMemsetAsync(b->bbytes.p, 0, bound * 4, st));
    pp_chain_kernel<<<rc_grid, RC_WARPS * 32, rc_smem, st>>>(b->uA.as<uint64_t>(), b->uklo.as<uint64_t>(), ukhi,
        b->ucnt.as<uint8_t>(), b->upm.as<uint8_t>(), b->seen.as<uint8_t>(), gh, b->bstart.as<uint32_t>(), k, m, in.abundance,
        b->visit.as<uint32_t>(), cnt);
    const unsigned em_grid = (unsigned)std::min<uint64_t>((bound + 7) / 8, 148 * 8);
    pp_emit_kernel<false><<<em_grid, 256, 0, st>>>(b->uA.as<uint64_t>(), b->uklo.as<uint64_t>(), ukhi, b->visit.as<uint32_t>(),
        b->bstart.as<uint32_t>(), k, m, input_shift, b->bbytes.as<uint32_t>(), b->bnmax.as<uint32_t>(), nullptr, nullptr,
        b->in_bytes.as<unsigned long long>(), cnt);
    PP_CK(excl_sum(b, b->bbytes.as<uint32_t>(), b->boff.as<uint64_t>(), bound, st));
    pp_body_total_kernel<<<1, 32, 0, st>>>(b->boff.as<uint64_t>(), b->bbytes.as<uint32_t>(), cnt);
    launched += 3;
    // ---- compare elements (needs the bucket tables, not the bytes)
    PP_CK(b->eflag.ensure(bound * 4)); PP_CK(b->eoff.ensure(bound * 4));
    PP_CK(b->el_min.ensure(bound * 4)); PP_CK(b->el_klo.ensure(bound * 8)); if (hi128) PP_CK(b->el_khi.ensure(bound * 8));
    pp_element_flag_kernel<<<nblk(bound), 256, 0, st>>>(b->uA.as<uint64_t>(), b->uklo.as<uint64_t>(), ukhi, b->ucnt.as<uint8_t>(),
        gh, k, in.abundance, bound, b->eflag.as<uint32_t>(), cnt);
    PP_CK(excl_sum(b, b->eflag.as<uint32_t>(), b->eoff.as<uint32_t>(), bound, st));
    pp_element_write_kernel<<<nblk(bound), 256, 0, st>>>(b->uA.as<uint64_t>(), b->uklo.as<uint64_t>(), ukhi, b->eflag.as<uint32_t>(),
        b->eoff.as<uint32_t>(), k, input_shift, bound, b->el_min.as<uint32_t>(), b->el_klo.as<uint64_t>(),
        hi128 ? b->el_khi.as<uint64_t>() : nullptr, b->in_elems.as<unsigned long long>(), cnt);
    launched += 2;
    // ---- sizes to the host, then the bytes
    PP_CK(b->h_cnt.ensure(sizeof(Counters))); PP_CK(b->h_in.ensure(nin * 8 * 3)); PP_CK(b->h_off.ensure((nin + 1) * 8 * 2));
    PP_CK(cudaMemcpyAsync(b->h_cnt.p, cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    unsigned long long *h_in = b->h_in.as<unsigned long long>();
    PP_CK(cudaMemcpyAsync(h_in, b->in_bytes.p, nin * 8, cudaMemcpyDeviceToHost, st));
    PP_CK(cudaMemcpyAsync(h_in + nin, b->in_sel.p, nin * 8, cudaMemcpyDeviceToHost, st));
    PP_CK(cudaMemcpyAsync(h_in + 2 * nin, b->in_elems.p, nin * 8, cudaMemcpyDeviceToHost, st));
    PP_CK(cudaStreamSynchronize(st));
    const Counters hc = *b->h_cnt.as<Counters>();
    if (hc.n_pieces > bound || hc.n_entries > bound) return cudaErrorUnknown;    // cannot happen: bound is exact
    PP_CK(b->body.ensure(hc.body_bytes ? hc.body_bytes : 1));
    PP_CK(b->h_body.ensure(hc.body_bytes ? hc.body_bytes : 1));
    pp_emit_kernel<true><<<em_grid, 256, 0, st>>>(b->uA.as<uint64_t>(), b->uklo.as<uint64_t>(), ukhi, b->visit.as<uint32_t>(),
        b->bstart.as<uint32_t>(), k, m, input_shift, b->bbytes.as<uint32_t>(), b->bnmax.as<uint32_t>(), b->boff.as<uint64_t>(),
        b->body.as<uint8_t>(), nullptr, cnt);
    launched++;
    if (hc.body_bytes) PP_CK(cudaMemcpyAsync(b->h_body.p, b->body.p, hc.body_bytes, cudaMemcpyDeviceToHost, st));
    PP_CK(cudaStreamSynchronize(st));
    uint64_t *off = b->h_off.as<uint64_t>();
    uint64_t *eoffh = off + (nin + 1);
    off[0] = 0;
    for (uint32_t i = 0; i < in.n_inputs; i++) off[i + 1] = off[i] + h_in[i];
    eoffh[in.n_inputs] = hc.n_elems;
    for (uint32_t i = in.n_inputs; i-- > 0;)            // an input without unique k-mers owns an empty range
        eoffh[i] = h_in[2 * nin + i] == ~0ULL ? eoffh[i + 1] : h_in[2 * nin + i];
    out->h_body = b->h_body.as<uint8_t>();
    out->h_body_off = off;
    out->h_selected = reinterpret_cast<const uint64_t *>(h_in + nin);
    out->h_elem_off = eoffh;
    out->d_minim = b->el_min.as<uint32_t>();
    out->d_klo = b->el_klo.as<uint64_t>();
    out->d_khi = hi128 ? b->el_khi.as<uint64_t>() : nullptr;
    out->n_elems = hc.n_elems;
    out->kernels_launched = launched;
    return cudaGetLastError();
}

}  // namespace spsp
