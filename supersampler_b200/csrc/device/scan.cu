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
                if (fp.kind == 1) {
                    // byte table: bit r of byte `key` (little-endian bytes inside the u32 words)
                    atomicOr(table + (key >> 2), (1u << r) << (8 * (key & 3)));
                } else {
                    uint32_t idx = table_index(key, fp);
                    atomicOr(table + (idx >> 5), 0x80000000u >> (idx & 31));
                }
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

// One warp owns 32 consecutive 64-base chunks (2048 bases, one coalesced 512 B
// read) per iteration and never synchronises with the rest of the CTA: probe
// results are packed into per-lane bit masks, compacted into the warp's own
// shared-memory queue with a shuffle scan, and verified by the whole warp
// against the exact bitmap.  The next iteration's words are prefetched into
// registers before the probes of the current one.
template <int G, bool HASHED>
__global__ void __launch_bounds__(FILTER_THREADS, FILTER_CTAS_PER_SM)
scan_filter_kernel(const uint32_t *__restrict__ packed, uint64_t n_bases, int m, FilterParams fp,
                   const uint32_t *__restrict__ table_g, const uint32_t *__restrict__ exact, ScanOut out)
{
    extern __shared__ uint32_t smem[];
    const uint32_t tbl_words = 1u << (fp.bits - 5);
    uint32_t *tbl = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wq = smem + tbl_words + warp * FILTER_WQ;     // this warp's candidate queue

    for (uint32_t i = threadIdx.x; i < tbl_words; i += blockDim.x) tbl[i] = __ldg(table_g + i);
    __syncthreads();
    if (n_bases < (uint64_t)m) return;

    const uint64_t n_pos = n_bases - m + 1;
    const uint64_t n_chunks = (n_pos + G - 1 + 63) >> 6;    // aligned probes reach G-1 past the last m-mer start
    const uint64_t n_groups = (n_chunks + 31) >> 5;
    const uint64_t n_warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const int ksh = 32 - 2 * fp.q;                          // window -> clean key (hashed mode)
    const int wsh = 32 - (fp.bits - 5);                     // index -> word number
    const int bsh = 32 - fp.bits;                           // index -> bit number (low 5 bits used)
    constexpr int NPROBE = 64 / G;
    constexpr int NB0 = NPROBE > 32 ? 32 : NPROBE;          // probes recorded in acc0

    uint64_t g = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    uint4 nv = make_uint4(0, 0, 0, 0);
    uint32_t nw4 = 0;
    if (g < n_groups) {
        const uint64_t c = (g << 5) + lane;
        if (c < n_chunks) {
            nv = __ldg(reinterpret_cast<const uint4 *>(packed) + c);
            nw4 = __ldg(packed + 4 * c + 4);
        }
    }
    for (; g < n_groups; g += n_warps) {
        uint32_t W[5];
        W[0] = nv.x; W[1] = nv.y; W[2] = nv.z; W[3] = nv.w; W[4] = nw4;
        {
            const uint64_t gn = g + n_warps, cn = (gn << 5) + lane;
            if (gn < n_groups && cn < n_chunks) {
                nv = __ldg(reinterpret_cast<const uint4 *>(packed) + cn);
                nw4 = __ldg(packed + 4 * cn + 4);
            }
        }
        const bool live = ((g << 5) + lane) < n_chunks;
        // one result bit per probe: probe i ends up at bit NB0-1-i of acc0 (i < 32) or bit NPROBE-1-i of acc1
        uint32_t acc0 = 0, acc1 = 0;
#pragma unroll
        for (int i = 0; i < NPROBE; i++) {
            const int a = i * G, j = a >> 4, o = a & 15;
            uint32_t x = window16(W[j], W[j + 1], o);
            if (HASHED) x = (x >> ksh) * 0x9E3779B1u;
            const uint32_t word = tbl[x >> wsh];
            const uint32_t t = __funnelshift_l(0u, word, x >> bsh);          // wanted bit -> bit 31
            if (i < 32) acc0 = __funnelshift_l(t, acc0, 1);
            else acc1 = __funnelshift_l(t, acc1, 1);
        }
        if (!live) { acc0 = 0; acc1 = 0; }
        // warp-level compaction of the positives
        const uint32_t cnt = __popc(acc0) + __popc(acc1);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        const uint64_t gbase = g << 11;                      // first base of this warp's block
        if (total <= FILTER_WQ) {
            uint32_t off = incl - cnt;
            const uint32_t rel = (uint32_t)lane << 6;
            while (acc0) {
                const int bit = 31 - __clz(acc0);
                acc0 &= ~(1u << bit);
                wq[off++] = rel + (uint32_t)((NB0 - 1 - bit) * G);
            }
            while (acc1) {
                const int bit = 31 - __clz(acc1);
                acc1 &= ~(1u << bit);
                wq[off++] = rel + (uint32_t)((NPROBE - 1 - bit) * G);
            }
            __syncwarp();
            for (uint32_t i = lane; i < total * G; i += 32) {
                const uint64_t a = gbase + wq[i / G];
                const uint32_t r = i % G;
                if (a >= r) verify_exact(packed, exact, a - r, n_bases, m, out);
            }
            __syncwarp();
        } else {
            // more positives than the queue holds (dense tables): every lane verifies its own
            const uint64_t cb = gbase + ((uint64_t)lane << 6);
            while (acc0) {
                const int bit = 31 - __clz(acc0);
                acc0 &= ~(1u << bit);
                const uint64_t a = cb + (uint64_t)((NB0 - 1 - bit) * G);
#pragma unroll 1
                for (int r = 0; r < G; r++)
                    if (a >= (uint64_t)r) verify_exact(packed, exact, a - r, n_bases, m, out);
            }
            while (acc1) {
                const int bit = 31 - __clz(acc1);
                acc1 &= ~(1u << bit);
                const uint64_t a = cb + (uint64_t)((NPROBE - 1 - bit) * G);
#pragma unroll 1
                for (int r = 0; r < G; r++)
                    if (a >= (uint64_t)r) verify_exact(packed, exact, a - r, n_bases, m, out);
            }
        }
    }
}

// ------------------------------------------------- byte-table filter (2q <= 16)

// Same idea with a denser encoding for short keys: the table has one BYTE per
// q-gram and the byte is the mask of phases r (0..G-1) at which some selected
// m-mer carries that q-gram.  A probe is then a byte extract (PRMT when the key
// is the 16 bits of two whole bytes), an 8-bit shared load whose address is the
// key itself, and one multiply-add that appends the 4-bit mask to a per-lane
// accumulator (8 probes per 32-bit register).  Only the phases named by the
// mask are verified.  Each lane handles FILTER8_CH chunks per iteration before
// the warp verifies its queue, which amortises the compaction and the verify
// pass over 8192 bases.
template <bool Q8>
__global__ void __launch_bounds__(FILTER8_THREADS, FILTER8_CTAS_PER_SM)
scan_filter8_kernel(const uint32_t *__restrict__ packed, uint64_t n_bases, int m, FilterParams fp,
                    const uint32_t *__restrict__ table_g, const uint32_t *__restrict__ exact, ScanOut out)
{
    constexpr int G = 4, NPROBE = 16, CH = FILTER8_CH;
    extern __shared__ uint32_t smem[];
    const uint32_t tbl_bytes = 1u << (2 * fp.q);
    const uint8_t *tbl = reinterpret_cast<const uint8_t *>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wq = smem + (tbl_bytes >> 2) + warp * (FILTER_WQ + 1);   // [0] = count, then entries
    uint32_t *wcnt = wq;
    wq += 1;

    for (uint32_t i = threadIdx.x; i < (tbl_bytes >> 2); i += blockDim.x) smem[i] = __ldg(table_g + i);
    if (lane == 0) *wcnt = 0;
    __syncthreads();
    if (n_bases < (uint64_t)m) return;

    const uint64_t n_pos = n_bases - m + 1;
    const uint64_t n_chunks = (n_pos + G - 1 + 63) >> 6;
    const uint64_t n_groups = (n_chunks + 32 * CH - 1) / (32 * CH);   // one group = CH x 32 chunks
    const uint64_t n_warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const int ksh = 32 - 2 * fp.q;

    uint64_t g = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    uint4 nv = make_uint4(0, 0, 0, 0);
    uint32_t nw4 = 0;
    if (g < n_groups) {
        const uint64_t c = g * (32 * CH) + lane;
        if (c < n_chunks) {
            nv = __ldg(reinterpret_cast<const uint4 *>(packed) + c);
            nw4 = __ldg(packed + 4 * c + 4);
        }
    }
    for (; g < n_groups; g += n_warps) {
        const uint64_t gchunk = g * (32 * CH);               // first chunk of the group
#pragma unroll 1
        for (int kk = 0; kk < CH; kk++) {
            uint32_t W[5];
            W[0] = nv.x; W[1] = nv.y; W[2] = nv.z; W[3] = nv.w; W[4] = nw4;
            {
                // next chunk of this lane: same group (kk+1) or the first of the warp's next group
                const uint64_t cn = (kk + 1 < CH) ? gchunk + (uint64_t)(kk + 1) * 32 + lane
                                                 : (g + n_warps) * (32 * CH) + lane;
                if (cn < n_chunks) {
                    nv = __ldg(reinterpret_cast<const uint4 *>(packed) + cn);
                    nw4 = __ldg(packed + 4 * cn + 4);
                }
            }
            const bool live = (gchunk + (uint64_t)kk * 32 + lane) < n_chunks;
            // probe i leaves its 4-bit phase mask in nibble (7 - i%8) of acc[i/8]
            uint32_t acc0 = 0, acc1 = 0;
#pragma unroll
            for (int i = 0; i < NPROBE; i++) {
                const int j = i >> 2, ob = i & 3;            // word, byte offset of the probe (4 bases per byte)
                uint32_t key;
                if (Q8) {
                    if (ob == 0) key = __byte_perm(W[j], 0u, 0x4432);
                    else if (ob == 1) key = __byte_perm(W[j], 0u, 0x4421);
                    else if (ob == 2) key = __byte_perm(W[j], 0u, 0x4410);
                    else key = __funnelshift_l(W[j + 1], W[j], 24) >> 16;
                } else {
                    key = window16(W[j], W[j + 1], ob * 4) >> ksh;
                }
                const uint32_t val = tbl[key];
                if (i < 8) acc0 = acc0 * 16u + val;
                else acc1 = acc1 * 16u + val;
            }
            if (!live) { acc0 = 0; acc1 = 0; }
            if (acc0 | acc1) {
                // nonzero nibbles -> one flag bit each, then claim queue slots
                uint32_t f0 = (acc0 | (acc0 >> 1) | (acc0 >> 2) | (acc0 >> 3)) & 0x11111111u;
                uint32_t f1 = (acc1 | (acc1 >> 1) | (acc1 >> 2) | (acc1 >> 3)) & 0x11111111u;
                const uint32_t cnt = __popc(f0) + __popc(f1);
                uint32_t off = atomicAdd(wcnt, cnt);
                const uint32_t rel = ((uint32_t)kk * 32 + lane) << 6;    // chunk start relative to the group
#pragma unroll 1
                for (int half = 0; half < 2; half++) {
                    uint32_t f = half ? f1 : f0;
                    const uint32_t acc = half ? acc1 : acc0;
                    while (f) {
                        const int bit = 31 - __clz(f);       // multiple of 4: nibble index = bit / 4
                        f &= ~(1u << bit);
                        const uint32_t mask = (acc >> bit) & 15u;
                        const uint32_t a = rel + (uint32_t)((half * 8 + (7 - (bit >> 2))) * G);
                        const uint32_t ent = (a << 4) | mask;
                        if (off < FILTER_WQ) {
                            wq[off] = ent;
                        } else {
                            // queue full: verify inline
                            const uint64_t pa = (gchunk << 6) + a;
#pragma unroll 1
                            for (int r = 0; r < G; r++)
                                if (((mask >> r) & 1u) && pa >= (uint64_t)r)
                                    verify_exact(packed, exact, pa - r, n_bases, m, out);
                        }
                        off++;
                    }
                }
            }
        }
        __syncwarp();
        const uint32_t total = min(*wcnt, (uint32_t)FILTER_WQ);
        if (total) {
            const uint64_t gbase = gchunk << 6;
            // 4 (phase) lanes per entry
            for (uint32_t i = lane; i < total * G; i += 32) {
                const uint32_t ent = wq[i >> 2];
                const uint32_t r = i & 3;
                const uint64_t a = gbase + (ent >> 4);
                if (((ent >> r) & 1u) && a >= r) verify_exact(packed, exact, a - r, n_bases, m, out);
            }
            __syncwarp();
            if (lane == 0) *wcnt = 0;
            __syncwarp();
        }
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
    cudaError_t e = cudaMemsetAsync(d_table, 0, filter_table_bytes(fp), st);
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
    const int threads = FILTER_THREADS;
    size_t smem = ((size_t)1 << (fp.bits - 3)) + (size_t)(threads / 32) * FILTER_WQ * sizeof(uint32_t);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(scan_filter_kernel<G, HASHED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(FILTER_MAX_SMEM + (FILTER_THREADS / 32) * FILTER_WQ * sizeof(uint32_t)));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    uint64_t n_chunks = ((n_bases - m + 1) + G - 1 + 63) >> 6;
    uint64_t n_groups = (n_chunks + 31) >> 5;
    uint64_t want = (n_groups + (threads / 32) - 1) / (threads / 32);     // CTAs if every warp took one group
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_filter_kernel<G, HASHED>, threads, smem);
    if (per_sm < 1) per_sm = 1;
    uint64_t blocks = (uint64_t)sm_count() * per_sm;
    if (blocks > want) blocks = want;
    if (blocks < 1) blocks = 1;
    scan_filter_kernel<G, HASHED><<<(unsigned)blocks, threads, smem, st>>>(d_packed, n_bases, m, fp, d_table, d_exact, out);
    return cudaGetLastError();
}

template <bool Q8>
static cudaError_t launch_filter8(const uint32_t *d_packed, uint64_t n_bases, int m, FilterParams fp,
                                  const uint32_t *d_table, const uint32_t *d_exact, ScanOut out, cudaStream_t st)
{
    const int threads = FILTER8_THREADS;
    size_t smem = filter_table_bytes(fp) + (size_t)(threads / 32) * (FILTER_WQ + 1) * sizeof(uint32_t);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(scan_filter8_kernel<Q8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(((size_t)64 << 10) + (FILTER8_THREADS / 32) * (FILTER_WQ + 1) * sizeof(uint32_t)));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    uint64_t n_chunks = ((n_bases - m + 1) + 4 - 1 + 63) >> 6;
    uint64_t n_groups = (n_chunks + 32 * FILTER8_CH - 1) / (32 * FILTER8_CH);
    uint64_t want = (n_groups + (threads / 32) - 1) / (threads / 32);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_filter8_kernel<Q8>, threads, smem);
    if (per_sm < 1) per_sm = 1;
    uint64_t blocks = (uint64_t)sm_count() * per_sm;
    if (blocks > want) blocks = want;
    if (blocks < 1) blocks = 1;
    scan_filter8_kernel<Q8><<<(unsigned)blocks, threads, smem, st>>>(d_packed, n_bases, m, fp, d_table, d_exact, out);
    return cudaGetLastError();
}

cudaError_t launch_scan_filter(const uint32_t *d_packed, uint64_t n_bases, int m, uint64_t thr, FilterParams fp,
                               const uint32_t *d_table, const uint32_t *d_exact, ScanOut out, cudaStream_t st)
{
    (void)thr;
    if (n_bases < (uint64_t)m) return cudaSuccess;
    if (fp.kind == 1) {
        if (fp.g != 4 || 2 * fp.q > 16) return cudaErrorInvalidValue;
        return fp.q == 8 ? launch_filter8<true>(d_packed, n_bases, m, fp, d_table, d_exact, out, st)
                         : launch_filter8<false>(d_packed, n_bases, m, fp, d_table, d_exact, out, st);
    }
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
