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

// Index of a clean q-gram value in the filter table (bits wide).  Direct mode
// needs bits == 2q; hashed mode is a multiplicative hash of the 2q-bit key.
__device__ __forceinline__ uint32_t table_index(uint32_t key, const FilterParams &fp)
{
    return fp.hashed ? (key * 0x9E3779B1u) >> (32 - fp.bits) : key;
}

// Enumerate the 4^m forward m-mers (32 per thread = one word of the exact
// bitmap); for each selected one set the filter bits of its G phase q-grams.
// Filter words keep index i at bit 31-(i&31) so that the scan tests a bit with
// one left funnel shift.
__global__ void filter_build_kernel(int m, uint64_t thr, FilterParams fp, uint32_t *__restrict__ table,
                                    uint32_t *__restrict__ exact, unsigned long long *__restrict__ n_selected)
{
    const uint64_t total_words = (1ULL << (2 * m)) >> 5;
    unsigned long long local = 0;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t bitsw = 0;
#pragma unroll 4
        for (int b = 0; b < 32; b++) {
            uint32_t fw = (uint32_t)(w << 5) | b;
            uint32_t rc = rc_mmer(fw, m);
            uint32_t cn = min(fw, rc);
            if (xxh64_8(cn) > thr) continue;
            bitsw |= 1u << b;
            local++;
            for (int r = 0; r < fp.g; r++) {
                // q-gram that starts r bases into the m-mer
                uint32_t key = (fw >> (2 * (m - fp.q - r))) & ((1u << (2 * fp.q)) - 1u);
                uint32_t idx = table_index(key, fp);
                atomicOr(table + (idx >> 5), 0x80000000u >> (idx & 31));
            }
        }
        exact[w] = bitsw;
    }
    if (local) atomicAdd(n_selected, local);
}

// ----------------------------------------------------------------- filter

// Verify one candidate position against the exact m-mer bitmap.
__device__ __forceinline__ void verify_exact(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ exact,
                                             uint64_t p, uint64_t n_bases, int m, const ScanOut &out)
{
    if (p + (uint64_t)m > n_bases) return;
    uint64_t w = p >> 4;
    int o = (int)(p & 15);
    uint32_t fw = window16(__ldg(packed + w), __ldg(packed + w + 1), o) >> (32 - 2 * m);
    if ((__ldg(exact + (fw >> 5)) >> (fw & 31)) & 1u) {
        uint32_t rc = rc_mmer(fw, m);
        emit_hit(out, p, min(fw, rc), rc < fw);
    }
}

template <int G, bool HASHED>
__global__ void __launch_bounds__(FILTER_THREADS) scan_filter_kernel(const uint32_t *__restrict__ packed,
                                                                       uint64_t n_bases, int m, FilterParams fp,
                                                                       const uint32_t *__restrict__ table_g,
                                                                       const uint32_t *__restrict__ exact,
                                                                       ScanOut out)
{
    extern __shared__ uint32_t smem[];
    const uint32_t tbl_words = 1u << (fp.bits - 5);
    const uint32_t R = 1u << fp.rep_log2;
    uint32_t *tbl = smem;                                   // tbl_words * R words, copy r of word i at i*R + r
    uint32_t *queue = smem + tbl_words * R;                 // FILTER_QUEUE entries
    __shared__ unsigned int q_count;

    for (uint32_t i = threadIdx.x; i < tbl_words * R; i += blockDim.x) tbl[i] = __ldg(table_g + (i >> fp.rep_log2));
    if (threadIdx.x == 0) q_count = 0;
    __syncthreads();
    if (n_bases < (uint64_t)m) return;

    const uint64_t n_pos = n_bases - m + 1;
    const uint64_t n_chunks = (n_pos + G - 1 + 63) >> 6;    // aligned probes reach G-1 past the last m-mer start
    const uint64_t n_tiles = (n_chunks + blockDim.x - 1) / blockDim.x;
    const int ksh = 32 - 2 * fp.q;                          // window -> clean key (hashed mode)
    const int wsh = 32 - (fp.bits - 5);                     // index -> word number
    const int bsh = 32 - fp.bits;                           // index -> bit number (low 5 bits used)
    const char *tbl_lane = reinterpret_cast<const char *>(tbl + (threadIdx.x & (R - 1)));
    const uint32_t stride = 4u << fp.rep_log2;              // bytes between consecutive table words
    constexpr int NPROBE = 64 / G;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t c = tile * blockDim.x + threadIdx.x;
        if (c < n_chunks) {
            uint4 v = ld_stream_u4(reinterpret_cast<const uint4 *>(packed) + c);
            uint32_t W[5];
            W[0] = v.x; W[1] = v.y; W[2] = v.z; W[3] = v.w;
            W[4] = __ldg(packed + 4 * c + 4);
            // one result bit per probe: probe i ends up at bit NPROBE-1-i of acc (two words for G == 1)
            uint32_t acc0 = 0, acc1 = 0;
#pragma unroll
            for (int i = 0; i < NPROBE; i++) {
                const int a = i * G, j = a >> 4, o = a & 15;
                uint32_t x = window16(W[j], W[j + 1], o);
                if (HASHED) x = (x >> ksh) * 0x9E3779B1u;
                const uint32_t word = *reinterpret_cast<const uint32_t *>(tbl_lane + (x >> wsh) * stride);
                const uint32_t t = __funnelshift_l(0u, word, x >> bsh);      // wanted bit -> bit 31
                if (i < 32) acc0 = __funnelshift_l(t, acc0, 1);
                else acc1 = __funnelshift_l(t, acc1, 1);
            }
            if (acc0 | acc1) {
                const uint32_t rel = threadIdx.x << 6;       // chunk start relative to the tile
#pragma unroll 1
                for (int half = 0; half < (NPROBE > 32 ? 2 : 1); half++) {
                    uint32_t acc = half ? acc1 : acc0;
                    const int nbit = NPROBE > 32 ? 32 : NPROBE;
                    while (acc) {
                        int bit = 31 - __clz(acc);
                        acc &= ~(1u << bit);
                        uint32_t a = (uint32_t)((half * 32 + (nbit - 1 - bit)) * G);
                        unsigned int slot = atomicAdd(&q_count, 1u);
                        if (slot < FILTER_QUEUE) {
                            queue[slot] = rel + a;
                        } else {
                            // queue full: verify inline (exact, just slower)
                            uint64_t pa = (c << 6) + a;
#pragma unroll 1
                            for (int r = 0; r < G; r++)
                                if (pa >= (uint64_t)r) verify_exact(packed, exact, pa - r, n_bases, m, out);
                        }
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
            if (a >= r) verify_exact(packed, exact, a - r, n_bases, m, out);
        }
        __syncthreads();
        if (threadIdx.x == 0) q_count = 0;
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

cudaError_t launch_filter_build(int m, uint64_t thr, FilterParams fp, uint32_t *d_table, uint32_t *d_exact,
                                unsigned long long *d_nsel, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(d_table, 0, (size_t)1 << (fp.bits - 3), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d_nsel, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    uint64_t total_words = (1ULL << (2 * m)) >> 5;
    uint64_t blocks = (total_words + 255) / 256;
    uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    filter_build_kernel<<<(unsigned)blocks, 256, 0, st>>>(m, thr, fp, d_table, d_exact, d_nsel);
    return cudaGetLastError();
}

template <int G, bool HASHED>
static cudaError_t launch_filter_g(const uint32_t *d_packed, uint64_t n_bases, int m, FilterParams fp,
                                   const uint32_t *d_table, const uint32_t *d_exact, ScanOut out, cudaStream_t st)
{
    size_t smem = (((size_t)1 << (fp.bits - 3)) << fp.rep_log2) + FILTER_QUEUE * sizeof(uint32_t);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(scan_filter_kernel<G, HASHED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(FILTER_MAX_SMEM + FILTER_QUEUE * sizeof(uint32_t)));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    uint64_t n_chunks = ((n_bases - m + 1) + G - 1 + 63) >> 6;
    // small inputs use smaller CTAs so that every SM gets a tile
    int threads = FILTER_THREADS;
    while (threads > 256 && (n_chunks + threads - 1) / threads < 2ull * sm_count()) threads >>= 1;
    uint64_t n_tiles = (n_chunks + threads - 1) / threads;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_filter_kernel<G, HASHED>, threads, smem);
    if (per_sm < 1) per_sm = 1;
    uint64_t blocks = (uint64_t)sm_count() * per_sm;
    if (blocks > n_tiles) blocks = n_tiles;
    if (blocks < 1) blocks = 1;
    scan_filter_kernel<G, HASHED><<<(unsigned)blocks, threads, smem, st>>>(d_packed, n_bases, m, fp, d_table, d_exact, out);
    return cudaGetLastError();
}

cudaError_t launch_scan_filter(const uint32_t *d_packed, uint64_t n_bases, int m, uint64_t thr, FilterParams fp,
                               const uint32_t *d_table, const uint32_t *d_exact, ScanOut out, cudaStream_t st)
{
    (void)thr;
    if (n_bases < (uint64_t)m) return cudaSuccess;
#define SPSP_F(G_, H_) return launch_filter_g<G_, H_>(d_packed, n_bases, m, fp, d_table, d_exact, out, st)
    switch (fp.g * 2 + (fp.hashed ? 1 : 0)) {
    case 2: SPSP_F(1, false);
    case 3: SPSP_F(1, true);
    case 4: SPSP_F(2, false);
    case 5: SPSP_F(2, true);
    case 8: SPSP_F(4, false);
    case 9: SPSP_F(4, true);
    default: return cudaErrorInvalidValue;
    }
#undef SPSP_F
}

}  // namespace spsp
