// Sketch-stage kernels for sm_100a: the closed form of the reference's
// per-base loop (SubSampler.cpp:367-440).  A position p of the packed sequence
// is a *hit* iff XXH64(canonical m-mer at p) <= T; the kernels emit the sparse
// hit list, everything else (minimizer replay, super-k-mers) happens on that
// list.  Two formulations:
//
//   scan_dense_kernel   full hash at every position (~40 integer ops/base);
//                       used when hits are dense (small -s) and as cross-check.
//   scan_filter_kernel  aligned q-gram filter in shared memory + exact verify.
//                       The set of selected forward m-mers S' is tiny
//                       (4^m * T/2^64, ~200 for k31/m11/s1000) and enumerable
//                       once per (m,T).  Every m-mer contains, at the first
//                       position a = 0 mod G inside it, a q-gram (q = m-G+1)
//                       of one of G possible "phases"; a bit table over those
//                       q-grams is probed once every G bases (direct index or
//                       multiplicative hash to <= 2^20 bits) and the rare
//                       positives are queued and verified with the full hash.
//                       No false negatives by construction, so the result is
//                       identical to the dense kernel's.
//
// Input layout: 2-bit bases, 16 per little-endian u32, first base in the MSBs;
// each thread owns 64 consecutive positions (one 128-bit load + 1 halo word).
#include "common.cuh"
#include "scan.cuh"

namespace spsp {

// Warp-aggregated append of one hit (called from divergent code).
__device__ __forceinline__ void emit_hit(const ScanOut &out, uint64_t pos, uint32_t canon, uint32_t rev)
{
    unsigned mask = __activemask();
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(out.count, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    unsigned long long slot = base + __popc(mask & ((1u << lane) - 1));
    if (slot < out.cap) {
        // one 16-byte store
        uint4 v;
        v.x = (uint32_t)pos; v.y = (uint32_t)(pos >> 32); v.z = canon; v.w = rev;
        reinterpret_cast<uint4 *>(out.hits)[slot] = v;
    }
}

// Exact test of one position (used by the verify phase and the tails).
__device__ __forceinline__ void verify_pos(const uint32_t *__restrict__ packed, uint64_t p, uint64_t n_bases,
                                           int m, uint64_t thr, const ScanOut &out)
{
    if (p + (uint64_t)m > n_bases) return;
    uint64_t w = p >> 4;
    int o = (int)(p & 15);
    uint32_t t = window16(__ldg(packed + w), __ldg(packed + w + 1), o);
    uint32_t fw = t >> (32 - 2 * m);
    uint32_t rc = rc_mmer(fw, m);
    uint32_t cn = min(fw, rc);
    if (xxh64_8(cn) <= thr) emit_hit(out, p, cn, cn != fw);
}

// ------------------------------------------------------------------ dense

__global__ void __launch_bounds__(256) scan_dense_kernel(const uint32_t *__restrict__ packed, uint64_t n_bases,
                                                         int m, uint64_t thr, ScanOut out)
{
    if (n_bases < (uint64_t)m) return;
    const uint64_t n_pos = n_bases - m + 1;
    const uint64_t n_chunks = (n_pos + 63) >> 6;
    const int sh = 32 - 2 * m;
    const int rsh = 2 * (16 - m);
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks;
         c += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v = ld_stream_u4(reinterpret_cast<const uint4 *>(packed) + c);
        uint32_t W[6];
        W[0] = v.x; W[1] = v.y; W[2] = v.z; W[3] = v.w;
        W[4] = __ldg(packed + 4 * c + 4);
        W[5] = 0;
        // reverse-complement window, pre-shifted so that the rc m-mer of
        // position p sits 64-p bases into S (compile-time offsets below)
        uint32_t R[6], S[6];
#pragma unroll
        for (int j = 0; j < 5; j++) R[j] = rc_word(W[4 - j]);
        R[5] = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) S[j] = __funnelshift_l(R[j + 1], R[j], rsh);
        S[5] = 0;
        const uint64_t base = c << 6;
#pragma unroll
        for (int p = 0; p < 64; p++) {
            const int j = p >> 4, o = p & 15;
            const int s = 64 - p, js = s >> 4, os = s & 15;
            uint32_t fw = window16(W[j], W[j + 1], o) >> sh;
            uint32_t rc = window16(S[js], S[js + 1], os) >> sh;
            uint32_t cn = min(fw, rc);
            uint64_t h = xxh64_8(cn);
            if (h <= thr && base + p < n_pos) emit_hit(out, base + p, cn, cn != fw);
        }
    }
}

// -------------------------------------------------- hit-set / table build

// Enumerate the 4^m forward m-mers; for each selected one set the table bits
// of its G phase q-grams and count it.
__global__ void filter_build_kernel(int m, uint64_t thr, FilterParams fp, uint32_t *__restrict__ table,
                                    unsigned long long *__restrict__ n_selected)
{
    const uint64_t total = 1ULL << (2 * m);
    unsigned long long local = 0;
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total;
         x += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t fw = (uint32_t)x;
        uint32_t rc = rc_mmer(fw, m);
        uint32_t cn = min(fw, rc);
        if (xxh64_8(cn) > thr) continue;
        local++;
        for (int r = 0; r < fp.g; r++) {
            // q-gram that starts r bases into the m-mer
            uint32_t key = (fw >> (2 * (m - fp.q - r))) & ((1u << (2 * fp.q)) - 1u);
            uint32_t idx = filter_index(key, fp);
            atomicOr(table + (idx >> 5), 1u << (idx & 31));
        }
    }
    // block-level reduction is not worth it: selected m-mers are rare
    if (local) atomicAdd(n_selected, local);
}

// ----------------------------------------------------------------- filter

template <int G>
__global__ void __launch_bounds__(FILTER_THREADS) scan_filter_kernel(const uint32_t *__restrict__ packed,
                                                                       uint64_t n_bases, int m, uint64_t thr,
                                                                       FilterParams fp,
                                                                       const uint32_t *__restrict__ table_g,
                                                                       ScanOut out)
{
    extern __shared__ uint32_t smem[];
    uint32_t *tbl = smem;                                   // 2^bits / 32 words
    const uint32_t tbl_words = 1u << (fp.bits - 5);
    uint32_t *queue = smem + tbl_words;                     // FILTER_QUEUE entries
    __shared__ unsigned int q_count;

    for (uint32_t i = threadIdx.x; i < tbl_words; i += blockDim.x) tbl[i] = __ldg(table_g + i);
    if (threadIdx.x == 0) q_count = 0;
    __syncthreads();
    if (n_bases < (uint64_t)m) return;

    const uint64_t n_pos = n_bases - m + 1;
    const uint64_t n_chunks = (n_pos + G - 1 + 63) >> 6;    // aligned probes reach G-1 past the last m-mer start
    const uint64_t n_tiles = (n_chunks + blockDim.x - 1) / blockDim.x;
    const int qsh = 32 - 2 * fp.q;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t c = tile * blockDim.x + threadIdx.x;
        if (c < n_chunks) {
            uint4 v = ld_stream_u4(reinterpret_cast<const uint4 *>(packed) + c);
            uint32_t W[5];
            W[0] = v.x; W[1] = v.y; W[2] = v.z; W[3] = v.w;
            W[4] = __ldg(packed + 4 * c + 4);
            const uint64_t base = c << 6;
#pragma unroll
            for (int a = 0; a < 64; a += G) {
                const int j = a >> 4, o = a & 15;
                uint32_t key = window16(W[j], W[j + 1], o) >> qsh;
                uint32_t idx = filter_index(key, fp);
                uint32_t word = tbl[idx >> 5];
                if ((word >> (idx & 31)) & 1u) {
                    unsigned int slot = atomicAdd(&q_count, 1u);
                    if (slot < FILTER_QUEUE) {
                        queue[slot] = (uint32_t)(base + a - tile * ((uint64_t)blockDim.x << 6));
                    } else {
                        // queue full: verify inline (exact, just slower)
#pragma unroll 1
                        for (int r = 0; r < G; r++)
                            if (base + a >= (uint64_t)r) verify_pos(packed, base + a - r, n_bases, m, thr, out);
                    }
                }
            }
        }
        __syncthreads();
        const unsigned int nq = min(q_count, (unsigned int)FILTER_QUEUE);
        const uint64_t tile_base = tile * ((uint64_t)blockDim.x << 6);
        for (unsigned int i = threadIdx.x; i < nq * G; i += blockDim.x) {
            uint64_t a = tile_base + queue[i / G];
            unsigned int r = i % G;
            if (a >= r) verify_pos(packed, a - r, n_bases, m, thr, out);
        }
        __syncthreads();
        if (threadIdx.x == 0) q_count = 0;
        // the next iteration's first use of q_count is after the chunk loop's
        // atomics; order them behind this reset
        __syncthreads();
    }
}

// ------------------------------------------------------------- launchers

static int g_sm_count = 0;
static int sm_count()
{
    if (!g_sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

cudaError_t launch_scan_dense(const uint32_t *d_packed, uint64_t n_bases, int m, uint64_t thr, ScanOut out,
                              cudaStream_t st)
{
    if (n_bases < (uint64_t)m) return cudaSuccess;
    uint64_t n_chunks = ((n_bases - m + 1) + 63) >> 6;
    uint64_t blocks = (n_chunks + 255) / 256;
    uint64_t cap = (uint64_t)sm_count() * 8;        // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    scan_dense_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_packed, n_bases, m, thr, out);
    return cudaGetLastError();
}

cudaError_t launch_filter_build(int m, uint64_t thr, FilterParams fp, uint32_t *d_table,
                                unsigned long long *d_nsel, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(d_table, 0, (size_t)1 << (fp.bits - 3), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d_nsel, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    uint64_t total = 1ULL << (2 * m);
    uint64_t blocks = (total + 255) / 256;
    uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    filter_build_kernel<<<(unsigned)blocks, 256, 0, st>>>(m, thr, fp, d_table, d_nsel);
    return cudaGetLastError();
}

template <int G>
static cudaError_t launch_filter_g(const uint32_t *d_packed, uint64_t n_bases, int m, uint64_t thr,
                                   FilterParams fp, const uint32_t *d_table, ScanOut out, cudaStream_t st)
{
    size_t smem = ((size_t)1 << (fp.bits - 3)) + FILTER_QUEUE * sizeof(uint32_t);
    static bool attr_set[3] = {false, false, false};
    const int gi = G == 1 ? 0 : G == 2 ? 1 : 2;
    if (!attr_set[gi]) {
        cudaError_t e = cudaFuncSetAttribute(scan_filter_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(((size_t)1 << (FILTER_MAX_BITS - 3)) + FILTER_QUEUE * sizeof(uint32_t)));
        if (e != cudaSuccess) return e;
        attr_set[gi] = true;
    }
    uint64_t n_chunks = ((n_bases - m + 1) + G - 1 + 63) >> 6;
    // large tables allow one CTA per SM: use 1024 threads; small inputs use
    // smaller CTAs so that every SM gets a tile
    int threads = FILTER_THREADS;
    while (threads > 256 && (n_chunks + threads - 1) / threads < 2ull * sm_count()) threads >>= 1;
    uint64_t n_tiles = (n_chunks + threads - 1) / threads;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_filter_kernel<G>, threads, smem);
    if (per_sm < 1) per_sm = 1;
    uint64_t blocks = (uint64_t)sm_count() * per_sm;
    if (blocks > n_tiles) blocks = n_tiles;
    if (blocks < 1) blocks = 1;
    scan_filter_kernel<G><<<(unsigned)blocks, threads, smem, st>>>(d_packed, n_bases, m, thr, fp, d_table, out);
    return cudaGetLastError();
}

cudaError_t launch_scan_filter(const uint32_t *d_packed, uint64_t n_bases, int m, uint64_t thr, FilterParams fp,
                               const uint32_t *d_table, ScanOut out, cudaStream_t st)
{
    if (n_bases < (uint64_t)m) return cudaSuccess;
    switch (fp.g) {
    case 1: return launch_filter_g<1>(d_packed, n_bases, m, thr, fp, d_table, out, st);
    case 2: return launch_filter_g<2>(d_packed, n_bases, m, thr, fp, d_table, out, st);
    case 4: return launch_filter_g<4>(d_packed, n_bases, m, thr, fp, d_table, out, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace spsp
