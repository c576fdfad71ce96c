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
#include <cstdlib>
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
            if (fp.kind == 2) {
                // level 2: one bit per selected m-mer in a hashed table of 2^hbits bits;
                // level 1: one bit per aligned 2-byte key and phase (compact image, one word per row)
                const uint32_t idx = (fw * 0x9E3779B1u) >> (32 - fp.hbits);
                atomicOr(table + (idx >> 5), 1u << (idx & 31));
                uint32_t *t1 = table + (1u << (fp.hbits - 5));
                for (int r = 0; r < 4; r++) {
                    uint32_t key = (fw >> (2 * (m - 8 - r))) & 0xFFFFu;
                    atomicOr(t1 + (key >> 6), 0x80000000u >> (key & 31));
                }
                continue;
            }
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

// ----------------------------------------------------------------- filter
//
// One warp owns FILTER_CH x 32 consecutive 64-base chunks per iteration (lane l
// of pass kk reads chunk group*CH*32 + kk*32 + l: one coalesced 512 B read per
// pass, the halo word comes from the neighbour lane by shuffle) and never
// synchronises with the rest of the CTA.  Probe results are appended to two
// per-lane accumulator words; lanes with a positive store {chunk, acc0, acc1}
// into the warp's shared-memory queue at a ballot-ranked slot (no atomics, no
// scan); when the queue could overflow or the group ends, every lane takes one
// entry and verifies the positions it names against the exact bitmap.  The
// next chunk's words are requested before the current chunk is probed, and the
// table itself arrives by one TMA bulk copy.
//
//   MODE_BIT_DIRECT / MODE_BIT_HASHED: bit table, 64/G probes of one bit;
//       a set bit means "some phase may match": all G phases are verified.
//   MODE_BYTE_Q8 / MODE_BYTE: G = 4, one byte per q-gram holding the 4-bit
//       phase mask; 16 probes of one nibble; only the named phases are verified.
enum { MODE_BIT_DIRECT = 0, MODE_BIT_HASHED = 1, MODE_BYTE_Q8 = 2, MODE_BYTE = 3 };

template <int G, int MODE>
__device__ __forceinline__ void verify_entry(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ exact,
                                             uint64_t cbase, uint32_t a0, uint32_t a1, uint64_t n_bases, int m,
                                             const ScanOut &out)
{
    constexpr int NPROBE = 64 / G;
    if (MODE >= MODE_BYTE_Q8) {
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            uint32_t acc = half ? a1 : a0;
            while (acc) {
                const int nib = (31 - __clz(acc)) >> 2;              // highest non-empty nibble
                uint32_t mask = (acc >> (nib * 4)) & 15u;
                acc &= ~(15u << (nib * 4));
                const uint64_t a = cbase + (uint64_t)((half * 8 + (7 - nib)) * G);
                while (mask) {
                    const uint32_t r = __ffs(mask) - 1;
                    mask &= mask - 1;
                    if (a >= r) verify_exact(packed, exact, a - r, n_bases, m, out);
                }
            }
        }
    } else {
        constexpr int NB0 = NPROBE > 32 ? 32 : NPROBE;               // probes recorded in a0
#pragma unroll 1
        for (int half = 0; half < (NPROBE > 32 ? 2 : 1); half++) {
            uint32_t acc = half ? a1 : a0;
            while (acc) {
                const int bit = 31 - __clz(acc);
                acc &= ~(1u << bit);
                const int probe = half ? (NPROBE - 1 - bit) : (NB0 - 1 - bit);
                const uint64_t a = cbase + (uint64_t)(probe * G);
#pragma unroll 1
                for (uint32_t r = 0; r < (uint32_t)G; r++)
                    if (a >= r) verify_exact(packed, exact, a - r, n_bases, m, out);
            }
        }
    }
}

template <int G, int MODE, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS)
scan_filter_kernel(const uint32_t *__restrict__ packed, uint64_t n_bases, int m, FilterParams fp,
                   const uint32_t *__restrict__ table_g, const uint32_t *__restrict__ exact, ScanOut out)
{
    constexpr int NPROBE = 64 / G, CH = FILTER_CH;
    static_assert(MODE < MODE_BYTE_Q8 || G == 4, "byte tables use G = 4");
    extern __shared__ __align__(128) uint32_t smem[];
    const uint32_t tbl_bytes = (uint32_t)filter_table_bytes(fp);
    const uint8_t *tbl8 = reinterpret_cast<const uint8_t *>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wq = smem + (tbl_bytes >> 2) + warp * (FILTER_WQ * 3);  // entries {chunk id in group, acc0, acc1}

    __shared__ __align__(8) uint64_t tbl_bar;
    const bool use_tma = tbl_bytes >= 16;
    if (use_tma) {
        if (threadIdx.x == 0) mbar_init(&tbl_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) tma_load_1d(smem, table_g, tbl_bytes, &tbl_bar);
    } else {
        for (uint32_t i = threadIdx.x; i < (tbl_bytes >> 2); i += blockDim.x) smem[i] = __ldg(table_g + i);
        __syncthreads();
    }
    if (n_bases < (uint64_t)m) { if (use_tma) mbar_wait(&tbl_bar, 0); return; }

    const uint64_t n_pos = n_bases - m + 1;
    const uint64_t n_chunks = (n_pos + G - 1 + 63) >> 6;    // aligned probes reach G-1 past the last m-mer start
    const uint64_t n_groups = (n_chunks + 32 * CH - 1) / (32 * CH);
    const uint64_t n_warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const int ksh = 32 - 2 * fp.q;                          // window -> clean key
    const int wsh = 32 - (fp.bits - 5);                     // bit table: index -> word number
    const int bsh = 32 - fp.bits;                           // bit table: index -> bit number (low 5 bits used)
    const uint32_t lt_mask = (1u << lane) - 1u;

    uint64_t g = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    uint4 nv = make_uint4(0, 0, 0, 0);
    uint32_t nw4 = 0;
    if (g < n_groups) {
        const uint64_t c = g * (32 * CH) + lane;
        if (c < n_chunks) {
            nv = __ldg(reinterpret_cast<const uint4 *>(packed) + c);
            if (lane == 31) nw4 = __ldg(packed + 4 * c + 4);
        }
    }
    if (use_tma) mbar_wait(&tbl_bar, 0);
    for (; g < n_groups; g += n_warps) {
        const uint64_t gchunk = g * (32 * CH);               // first chunk of the group
        uint32_t qn = 0;                                     // queued entries (warp-uniform)
#pragma unroll 1
        for (int kk = 0; kk < CH; kk++) {
            uint32_t W[5];
            W[0] = nv.x; W[1] = nv.y; W[2] = nv.z; W[3] = nv.w;
            W[4] = __shfl_down_sync(0xffffffffu, nv.x, 1);   // halo = neighbour lane's first word
            if (lane == 31) W[4] = nw4;
            {
                // next chunk of this lane: same group (kk+1) or the first of the warp's next group
                const uint64_t cn = (kk + 1 < CH) ? gchunk + (uint64_t)(kk + 1) * 32 + lane
                                                 : (g + n_warps) * (32 * CH) + lane;
                if (cn < n_chunks) {
                    nv = __ldg(reinterpret_cast<const uint4 *>(packed) + cn);
                    if (lane == 31) nw4 = __ldg(packed + 4 * cn + 4);
                }
            }
            const bool live = (gchunk + (uint64_t)kk * 32 + lane) < n_chunks;
            uint32_t acc0 = 0, acc1 = 0;
            if (MODE >= MODE_BYTE_Q8) {
                // probe i leaves its 4-bit phase mask in nibble (7 - i%8) of acc[i/8]
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int j = i >> 2, ob = i & 3;        // word, byte offset of the probe (4 bases per byte)
                    uint32_t key;
                    if (MODE == MODE_BYTE_Q8) {
                        if (ob == 0) key = __byte_perm(W[j], 0u, 0x4432);
                        else if (ob == 1) key = __byte_perm(W[j], 0u, 0x4421);
                        else if (ob == 2) key = __byte_perm(W[j], 0u, 0x4410);
                        else key = __funnelshift_l(W[j + 1], W[j], 24) >> 16;
                    } else {
                        key = window16(W[j], W[j + 1], ob * 4) >> ksh;
                    }
                    const uint32_t val = tbl8[key];
                    if (i < 8) acc0 = acc0 * 16u + val;
                    else acc1 = acc1 * 16u + val;
                }
            } else {
                // probe i ends up at bit NB0-1-i of acc0 (i < 32) or bit NPROBE-1-i of acc1
#pragma unroll
                for (int i = 0; i < NPROBE; i++) {
                    const int a = i * G, j = a >> 4, o = a & 15;
                    uint32_t x = window16(W[j], W[j + 1], o);
                    if (MODE == MODE_BIT_HASHED) x = (x >> ksh) * 0x9E3779B1u;
                    const uint32_t word = smem[x >> wsh];
                    const uint32_t t = __funnelshift_l(0u, word, x >> bsh);  // wanted bit -> bit 31
                    if (i < 32) acc0 = __funnelshift_l(t, acc0, 1);
                    else acc1 = __funnelshift_l(t, acc1, 1);
                }
            }
            const bool pos = live && (acc0 | acc1) != 0;
            const uint32_t bal = __ballot_sync(0xffffffffu, pos);
            if (pos) {
                uint32_t *e = wq + (qn + __popc(bal & lt_mask)) * 3;
                e[0] = (uint32_t)kk * 32 + lane;
                e[1] = acc0;
                e[2] = acc1;
            }
            qn += __popc(bal);
            // the queue always has room for one more chunk (32 entries); verify when it could overflow
            if (qn > FILTER_WQ - 32 || kk == CH - 1) {
                __syncwarp();
                const uint64_t gbase = gchunk << 6;
                for (uint32_t ei = lane; ei < qn; ei += 32) {
                    const uint32_t *e = wq + ei * 3;
                    verify_entry<G, MODE>(packed, exact, gbase + ((uint64_t)e[0] << 6), e[1], e[2], n_bases, m, out);
                }
                __syncwarp();
                qn = 0;
            }
        }
    }
}


// ------------------------------------------------------------- row-bit filter
//
// kind 2, m == 11 (the reference's default).  An 11-mer that starts at base p
// fully contains the two aligned sequence bytes b0 = ceil(p/4) and b0+1, so
// probing the 16-bit key (byte b, byte b+1) at every byte b >= 1 finds every
// occurrence exactly once, as (b0, r = 4*b0 - p); p = 0 is checked apart.
//
// Level 1: a bit table over 15 of the 16 key bits (row = key >> 6, bit =
// key & 31), replicated so that lane l owns word [row][l]: the 32 lookups of
// a warp fall into 32 different banks, i.e. ONE shared-memory wavefront per
// probe instead of ~3.6 for random addresses in a single copy.  A set bit
// means "some phase of some selected m-mer has this key".  The kernel is then
// bound by instruction issue on the two integer pipes, so the work of a probe
// is split between them: left shifts and the address are IMADs (multipliers
// 2^n read from the constant bank, which keeps ptxas from turning them back
// into shifts) on the FMA pipe; right shifts and the two funnel shifts that
// move the wanted bit into the accumulator stay on the ALU pipe.  (Measured
// on B200, 320 Mbp: all-ALU 53.0 us, IMAD.HI for the right shifts too 54.9 us,
// this split 51.5 us; the byte-table kernel 58 us.)
// Level 2: lanes with a flagged probe park {chunk, flags, their 5 sequence
// words} in the warp's ring in shared memory; when 32 entries wait, every lane
// takes one and tests the four m-mers (phases r = 0..3) of each flagged byte
// against a hashed bit table of the selected forward m-mers (2^hbits bits,
// also in shared memory); the few survivors are settled by the exact bitmap
// over all 4^11 m-mers (global, L2-resident).  Nothing else after the
// streaming load touches global memory except the hit records.
//
// A lane's chunk c covers the probes at bytes 16c+1 .. 16c+16, all inside its
// five words (bytes 16c .. 16c+19), so the parked entry is self-contained.
__constant__ uint32_t c_pow2[3] = {1u << 8, 1u << 16, 128u};

__device__ __forceinline__ void rowbit_check(const uint32_t *l2, int l2sh, const uint32_t *__restrict__ exact, uint64_t p,
                                             uint32_t fw, uint64_t n_bases, const ScanOut &out)
{
    const uint32_t idx = (fw * 0x9E3779B1u) >> l2sh;
    if ((l2[idx >> 5] >> (idx & 31)) & 1u) {
        if (p + 11 <= n_bases && ((__ldg(exact + (fw >> 5)) >> (fw & 31)) & 1u)) {
            const uint32_t rc = rc_word(fw) >> 10;
            emit_hit(out, p, min(fw, rc), rc < fw);
        }
    }
}

// one parked entry: every flagged byte, four phases each
__device__ __forceinline__ void rowbit_verify(const uint32_t *q, uint32_t e, const uint32_t *l2, int l2sh,
                                              const uint32_t *__restrict__ exact, uint64_t n_bases, const ScanOut &out)
{
    const uint32_t c = q[e];
    uint32_t acc = q[ROWBIT_Q + e];
    while (acc) {
        const int b = 31 - __clz(acc);
        acc ^= 1u << b;
        const int i = 16 - b;                                   // byte 1..16 of the chunk's 20-byte window
        const int j = i >> 2, ob = i & 3;
        const uint32_t wa = j ? q[(1 + j) * ROWBIT_Q + e] : 0u;   // words j-1, j, j+1 of the window
        const uint32_t wb = q[(2 + j) * ROWBIT_Q + e];
        const uint32_t wc = j < 4 ? q[(3 + j) * ROWBIT_Q + e] : 0u;
        const uint64_t p0 = ((uint64_t)c << 6) + (uint64_t)(4 * i);
        // the four starts 4i-3 .. 4i sit 13+4ob .. 16+4ob bases into (wa, wb, wc)
        const uint32_t hi = ob ? wb : wa, lo = ob ? wc : wb;      // 32-base window that holds all four
        const int s0 = ob ? 4 * ob - 3 : 13;                      // first start inside (hi, lo); last start <= 16
        uint32_t fw[4], any = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            // phase r starts at 4i - r = s0 + (3 - r); clamped shift: 32 yields lo
            fw[r] = __funnelshift_lc(lo, hi, 2 * (s0 + 3 - r)) >> 10;
            const uint32_t idx = (fw[r] * 0x9E3779B1u) >> l2sh;
            any |= ((l2[idx >> 5] >> (idx & 31)) & 1u) << r;
        }
        if (any) {                                              // rare: settle with the exact bitmap
#pragma unroll
            for (int r = 0; r < 4; r++)
                if ((any >> r) & 1u) rowbit_check(l2, l2sh, exact, p0 - r, fw[r], n_bases, out);
        }
    }
}

template <int NOVERIFY>
__global__ void __launch_bounds__(1024, 1)
scan_rowbit_kernel(const uint32_t *__restrict__ packed, uint64_t n_bases, int hbits, const uint32_t *__restrict__ table_g,
                   const uint32_t *__restrict__ exact, ScanOut out)
{
    extern __shared__ __align__(128) uint32_t smem[];
    uint32_t *t1 = smem;                                        // [ROWBIT_ROWS][32]
    uint32_t *l2 = smem + ROWBIT_ROWS * 32;                     // 2^hbits bits
    uint32_t *qbase = l2 + (1u << (hbits - 5));                 // rings; first 4 KB = staging of the compact table
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *q = qbase + warp * (ROWBIT_Q * ROWBIT_EW);        // [word][slot]
    const int l2sh = 32 - hbits;

    __shared__ __align__(8) uint64_t tbl_bar;
    if (threadIdx.x == 0) mbar_init(&tbl_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) tma_load_1d(l2, table_g, (1u << (hbits - 3)) + ROWBIT_ROWS * 4u, &tbl_bar);

    if (n_bases < 11) { mbar_wait(&tbl_bar, 0); return; }
    const uint64_t n_pos = n_bases - 11 + 1;
    const uint64_t b0max = (n_pos - 1 + 3) >> 2;                // last byte that can be a first full byte
    const uint32_t n_chunks = (uint32_t)((b0max + 15) >> 4);
    const uint32_t n_pass = (n_chunks + 31) >> 5;
    const uint32_t n_warps = gridDim.x * 32u;
    const uint4 *p4 = reinterpret_cast<const uint4 *>(packed);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t lane4 = (uint32_t)lane * 4u;

    uint32_t pass = blockIdx.x * 32u + warp;
    uint4 nv = make_uint4(0, 0, 0, 0);
    uint32_t nw4 = 0;
    if (pass < n_pass) {
        const uint32_t c = min(pass * 32u + lane, n_chunks);  // chunk n_chunks (padding) supplies the last halo word
        nv = __ldg(p4 + c);
        if (lane == 31) nw4 = __ldg(packed + 4 * (size_t)c + 4);
    }
    mbar_wait(&tbl_bar, 0);
    {
        const uint32_t *stage = qbase;                          // compact table: one word per row
        for (int w = warp; w < ROWBIT_ROWS; w += 32) t1[w * 32 + lane] = stage[w];
    }
    __syncthreads();                                            // the staging area becomes ring space

    if (blockIdx.x == 0 && threadIdx.x == 0)                    // p = 0 has no first full byte >= 1
        rowbit_check(l2, l2sh, exact, 0, __ldg(packed) >> 10, n_bases, out);

    const char *t1b = reinterpret_cast<const char *>(t1);
    const uint32_t k8 = c_pow2[0], k16 = c_pow2[1], k128 = c_pow2[2];
    uint32_t head = 0, cnt = 0;                                 // ring state (warp-uniform)
    for (; pass < n_pass; pass += n_warps) {
        uint32_t W[5];
        W[0] = nv.x; W[1] = nv.y; W[2] = nv.z; W[3] = nv.w;
        W[4] = __shfl_down_sync(0xffffffffu, nv.x, 1);
        if (lane == 31) W[4] = nw4;
        const uint32_t c = pass * 32u + lane;
        {
            // next pass of this warp; past the end the padding chunk is re-read (cached, and never used):
            // no branch, so the loads stay ahead of the probes
            uint32_t cn = (pass + n_warps) * 32u + lane;
            cn = min(cn, n_chunks);
            nv = ld_stream_u4_pinned(p4 + cn);
            if (lane == 31) nw4 = ld_stream_u32(packed + 4 * (size_t)cn + 4);
        }
        uint32_t acc = 0;                                       // probe of byte i -> bit 16 - i
#pragma unroll
        for (int i = 1; i <= 16; i++) {
            const int j = i >> 2, ob = i & 3;
            // x: key in the top 16 bits; row = x >> 22; ks: any word whose low 5 bits are key bits 4..0
            uint32_t x, ks;
            if (ob == 0) { x = W[j]; ks = W[j] >> 16; }
            else if (ob == 1) { x = W[j] * k8; ks = W[j] >> 8; }
            else if (ob == 2) { x = W[j] * k16; ks = W[j]; }
            else { x = __funnelshift_l(W[j + 1], W[j], 24); ks = x >> 16; }
            const uint32_t addr = (x >> 22) * k128 + lane4;
            const uint32_t word = *reinterpret_cast<const uint32_t *>(t1b + addr);
            const uint32_t t = __funnelshift_l(0u, word, ks);   // wanted bit -> bit 31 (shift taken mod 32)
            acc = __funnelshift_l(t, acc, 1);
        }
        const bool pos = c < n_chunks && acc != 0;
        const uint32_t bal = __ballot_sync(0xffffffffu, pos);
        if (pos) {
            const uint32_t e = (head + cnt + __popc(bal & lt_mask)) & (ROWBIT_Q - 1);
            q[e] = c;
            q[ROWBIT_Q + e] = acc;
#pragma unroll
            for (int j = 0; j < 5; j++) q[(2 + j) * ROWBIT_Q + e] = W[j];
        }
        cnt += __popc(bal);
        if (cnt > ROWBIT_Q - 32) {                              // room for the next pass: drain one round of 32
            __syncwarp();
            if (!NOVERIFY) rowbit_verify(q, (head + lane) & (ROWBIT_Q - 1), l2, l2sh, exact, n_bases, out);
            __syncwarp();
            head = (head + 32) & (ROWBIT_Q - 1);
            cnt -= 32;
        }
    }
    __syncwarp();
    while (cnt) {
        const uint32_t n = cnt < 32 ? cnt : 32;
        if ((uint32_t)lane < n) rowbit_verify(q, (head + lane) & (ROWBIT_Q - 1), l2, l2sh, exact, n_bases, out);
        head = (head + n) & (ROWBIT_Q - 1);
        cnt -= n;
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

template <int G, int MODE, int THREADS, int CTAS>
static cudaError_t launch_filter_t(const uint32_t *d_packed, uint64_t n_bases, int m, FilterParams fp,
                                   const uint32_t *d_table, const uint32_t *d_exact, ScanOut out, cudaStream_t st)
{
    auto kern = scan_filter_kernel<G, MODE, THREADS, CTAS>;
    size_t smem = filter_table_bytes(fp) + (size_t)(THREADS / 32) * (FILTER_WQ * 3) * sizeof(uint32_t);
    int per_sm = 0;
    {
        static PerDeviceOnce once;                  // one per template instance = per kernel
        cudaError_t e = once.run([&] {
            return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(FILTER_MAX_SMEM + (THREADS / 32) * (FILTER_WQ * 3) * sizeof(uint32_t)));
        });
        if (e != cudaSuccess) return e;
    }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    uint64_t n_chunks = ((n_bases - m + 1) + G - 1 + 63) >> 6;
    uint64_t n_groups = (n_chunks + 32 * FILTER_CH - 1) / (32 * FILTER_CH);
    uint64_t want = (n_groups + (THREADS / 32) - 1) / (THREADS / 32);     // CTAs if every warp took one group
    uint64_t blocks = (uint64_t)sm_count() * per_sm;
    if (blocks > want) blocks = want;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, THREADS, smem, st>>>(d_packed, n_bases, m, fp, d_table, d_exact, out);
    return cudaGetLastError();
}

// Tables of up to 64 KB run as 2 CTAs x 768 threads per SM; the 128 KB bit table as 1 CTA x 1024.
template <int G, int MODE>
static cudaError_t launch_filter_m(const uint32_t *d_packed, uint64_t n_bases, int m, FilterParams fp,
                                   const uint32_t *d_table, const uint32_t *d_exact, ScanOut out, cudaStream_t st)
{
    if (filter_table_bytes(fp) <= ((size_t)64 << 10))
        return launch_filter_t<G, MODE, 768, 2>(d_packed, n_bases, m, fp, d_table, d_exact, out, st);
    return launch_filter_t<G, MODE, 1024, 1>(d_packed, n_bases, m, fp, d_table, d_exact, out, st);
}


static size_t rowbit_smem(int hbits)
{
    return (size_t)ROWBIT_ROWS * 32 * 4 + ((size_t)1 << (hbits - 3)) + (size_t)32 * ROWBIT_Q * ROWBIT_EW * 4;
}

static cudaError_t launch_rowbit(const uint32_t *d_packed, uint64_t n_bases, FilterParams fp, const uint32_t *d_table,
                                 const uint32_t *d_exact, ScanOut out, cudaStream_t st)
{
    static const bool noverify = getenv("SPSP_ROWBIT_NOVERIFY") != nullptr;   // timing experiments only: drops the hits
    auto kern = noverify ? scan_rowbit_kernel<1> : scan_rowbit_kernel<0>;
    {
        static PerDeviceOnce once[2];
        cudaError_t e = once[noverify ? 1 : 0].run([&] {
            return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rowbit_smem(ROWBIT_MAX_HBITS));
        });
        if (e != cudaSuccess) return e;
    }
    const uint64_t n_pos = n_bases - 11 + 1;
    const uint64_t n_chunks = ((((n_pos - 1 + 3) >> 2) + 15) >> 4);
    if (n_chunks >= (1ull << 32) - 64) return cudaErrorInvalidValue;
    uint64_t n_pass = (n_chunks + 31) >> 5;
    uint64_t blocks = (uint64_t)sm_count();
    const uint64_t want = (n_pass + 31) / 32;
    if (blocks > want) blocks = want;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, 1024, rowbit_smem(fp.hbits), st>>>(d_packed, n_bases, fp.hbits, d_table, d_exact, out);
    return cudaGetLastError();
}

cudaError_t launch_scan_filter(const uint32_t *d_packed, uint64_t n_bases, int m, uint64_t thr, FilterParams fp,
                               const uint32_t *d_table, const uint32_t *d_exact, ScanOut out, cudaStream_t st)
{
    (void)thr;
    if (n_bases < (uint64_t)m) return cudaSuccess;
    if (fp.kind == 2) {
        if (m != 11 || fp.hbits < ROWBIT_MIN_HBITS || fp.hbits > ROWBIT_MAX_HBITS) return cudaErrorInvalidValue;
        return launch_rowbit(d_packed, n_bases, fp, d_table, d_exact, out, st);
    }
#define SPSP_F(G_, M_) return launch_filter_m<G_, M_>(d_packed, n_bases, m, fp, d_table, d_exact, out, st)
    if (fp.kind == 1) {
        if (fp.g != 4 || 2 * fp.q > 16) return cudaErrorInvalidValue;
        if (fp.q == 8) SPSP_F(4, MODE_BYTE_Q8);
        SPSP_F(4, MODE_BYTE);
    }
    switch (fp.g * 2 + (fp.hashed ? 1 : 0)) {
    case 2: SPSP_F(1, MODE_BIT_DIRECT);
    case 3: SPSP_F(1, MODE_BIT_HASHED);
    case 4: SPSP_F(2, MODE_BIT_DIRECT);
    case 5: SPSP_F(2, MODE_BIT_HASHED);
    case 8: SPSP_F(4, MODE_BIT_DIRECT);
    case 9: SPSP_F(4, MODE_BIT_HASHED);
    default: return cudaErrorInvalidValue;
    }
#undef SPSP_F
}

}  // namespace spsp
