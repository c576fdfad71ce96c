// Compare-stage kernels for sm_100a: replaces the reference's N-way merge +
// count_intersection + compute_scores (Comparator.cpp:39-74, :177-287).
//
// Data in HBM (SoA over all elements of all sketches, sketch after sketch):
//   minim[E] u32  minimizer of the element's bucket (ascending inside a sketch)
//   klo[E]   u64  canonical k-mer, low 64 bits;  khi[E] u64 only when k > 32
//   sk_off[N+1]   element range of each sketch
// An element is one distinct (bucket, canonical k-mer) of a sketch; the result
// is out[i][j] = number of elements sketches i and j share -- exactly the
// per-bucket colour counting of the reference.
//
// The minimizer space is cut into C equal value ranges ("chunks"; selected
// minimizers are hash-uniform so equal ranges balance).  A CTA owns one *unit*:
// a tile of 32 column sketches, up to CMP_RT consecutive tiles of 32 row
// sketches, and a subset of the chunks.  Per chunk it builds, in shared memory,
// a hash table of the column sketches' elements (key -> 32-bit membership mask
// over the tile's columns) ONCE, then every warp streams the elements of its
// row sketches of every row tile through the table and turns the 32 lane masks
// into 32 per-column counts with warp ballots -- one probe serves 32 sketch
// pairs, one table build serves up to 32 x CMP_RT rows.  Column batches larger
// than the table are processed in several passes, so capacity never affects the
// result.
#include "common.cuh"
#include "compare.cuh"

namespace spsp {

__global__ void chunk_offsets_kernel(const uint32_t *__restrict__ minim, const uint64_t *__restrict__ sk_off,
                                     const uint64_t *__restrict__ sk_end, uint32_t n_sketches, const uint32_t *__restrict__ dev_dims,
                                     uint32_t n_chunks, uint64_t space, uint64_t *__restrict__ chunk_off)
{
    if (dev_dims) n_sketches = dev_dims[1];
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t total = (uint64_t)n_sketches * (n_chunks + 1);
    if (t >= total) return;
    uint32_t s = (uint32_t)(t / (n_chunks + 1)), c = (uint32_t)(t % (n_chunks + 1));
    uint64_t lo = sk_off[s], hi = sk_end ? sk_end[s] : sk_off[s + 1];
    // first element with minimizer >= boundary(c)
    uint64_t bound = space * c / n_chunks;          // space <= 2^30, c <= 2^13
    if (c == n_chunks) { chunk_off[t] = hi; return; }
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if ((uint64_t)minim[mid] < bound) lo = mid + 1; else hi = mid;
    }
    chunk_off[t] = lo;
}

__device__ __forceinline__ uint32_t elem_hash(uint64_t lo, uint64_t hi, uint32_t mn)
{
    uint64_t x = lo ^ (hi * 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)mn * 0xC2B2AE3D27D4EB4FULL);
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ULL;
    x ^= x >> 32;
    return (uint32_t)x & (CMP_SLOTS - 1);
}

// units[u] = {jb | ib0 << 16, n_ib}: column tile jb, row tiles ib0 .. ib0 + n_ib - 1.
// dev_dims (nullable, multi-GPU exchange): {rows, columns, units of this rank} computed on the device; then the
// ranges start at 0 and the counts go to compact per-unit tiles: out[(unit * CMP_RT + t) * 1024 + row * 32 + col].
template <bool HAS_HI>
__global__ void __launch_bounds__(CMP_THREADS, 1)
hashjoin_kernel(CmpData d, const uint2 *__restrict__ units, const uint32_t *__restrict__ dev_dims, uint32_t n_chunks,
                uint32_t row_begin, uint32_t row_end, uint32_t col_begin, uint32_t col_end, uint32_t *__restrict__ out, uint64_t ld)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *s_klo = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *s_khi = s_klo + CMP_CAP;                              // only touched when HAS_HI
    uint32_t *s_min = reinterpret_cast<uint32_t *>(s_klo + (HAS_HI ? 2 : 1) * CMP_CAP);
    uint32_t *s_mask = s_min + CMP_CAP;
    uint32_t *s_slot = s_mask + CMP_CAP;
    uint32_t *s_acc = s_slot + CMP_SLOTS;                           // [RT][32 rows][32 cols]
    __shared__ uint64_t s_jbeg[32];
    __shared__ uint32_t s_pref[33];
    constexpr int RT = HAS_HI ? CMP_RT_HI : CMP_RT;

    const bool compact = dev_dims != nullptr;
    if (compact) {
        if (blockIdx.x >= dev_dims[2]) return;
        row_begin = col_begin = 0;
        row_end = dev_dims[0]; col_end = dev_dims[1];
    }
    const uint2 unit = units[blockIdx.x];
    const uint32_t jb = unit.x & 0xFFFFu, ib0 = unit.x >> 16, n_ib = unit.y;
    const uint32_t col0 = col_begin + jb * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = CMP_THREADS / 32, RPW = 32 / NW;
    for (uint32_t i = threadIdx.x; i < (uint32_t)RT * 1024; i += CMP_THREADS) s_acc[i] = 0;

    for (uint32_t c = blockIdx.y; c < n_chunks; c += gridDim.y) {
        // column ranges of this chunk + exclusive prefix of their lengths
        if (warp == 0) {
            uint32_t j = col0 + lane;
            uint64_t b = 0, e = 0;
            if (j < col_end) {
                b = d.chunk_off[(uint64_t)j * (n_chunks + 1) + c];
                e = d.chunk_off[(uint64_t)j * (n_chunks + 1) + c + 1];
            }
            uint32_t len = (uint32_t)(e - b), inc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            s_jbeg[lane] = b;
            s_pref[lane + 1] = inc;
            if (lane == 0) s_pref[0] = 0;
        }
        __syncthreads();
        const uint32_t n_j = s_pref[32];
        for (uint32_t b0 = 0; b0 < n_j; b0 += CMP_CAP) {
            const uint32_t nb = min((uint32_t)CMP_CAP, n_j - b0);
            for (uint32_t i = threadIdx.x; i < CMP_SLOTS; i += CMP_THREADS) s_slot[i] = 0;
            for (uint32_t i = threadIdx.x; i < nb; i += CMP_THREADS) s_mask[i] = 0;
            // stage the batch
            for (uint32_t i = threadIdx.x; i < nb; i += CMP_THREADS) {
                uint32_t f = b0 + i;
                int jj = 0;
#pragma unroll
                for (int s = 16; s; s >>= 1) if (s_pref[jj + s] <= f) jj += s;
                uint64_t g = s_jbeg[jj] + (f - s_pref[jj]);
                s_klo[i] = d.klo[g];
                if (HAS_HI) s_khi[i] = d.khi[g];
                s_min[i] = d.minim[g];
            }
            __syncthreads();
            // insert: slot owns the first element that claimed it
            for (uint32_t i = threadIdx.x; i < nb; i += CMP_THREADS) {
                uint32_t f = b0 + i;
                int jj = 0;
#pragma unroll
                for (int s = 16; s; s >>= 1) if (s_pref[jj + s] <= f) jj += s;
                const uint64_t lo = s_klo[i], hi = HAS_HI ? s_khi[i] : 0;
                const uint32_t mn = s_min[i];
                uint32_t h = elem_hash(lo, hi, mn);
                for (;;) {
                    uint32_t prev = atomicCAS(&s_slot[h], 0u, i + 1);
                    uint32_t o = prev ? prev - 1 : i;
                    if (prev == 0 || (s_klo[o] == lo && s_min[o] == mn && (!HAS_HI || s_khi[o] == hi))) {
                        atomicOr(&s_mask[o], 1u << jj);
                        break;
                    }
                    h = (h + 1) & (CMP_SLOTS - 1);
                }
            }
            __syncthreads();
            // probe: the table serves every row tile of the unit; warp w streams rows w, w+NW, ... of each
            for (uint32_t t = 0; t < n_ib; t++) {
                const uint32_t row0 = row_begin + (ib0 + t) * 32;
#pragma unroll
                for (int r = 0; r < RPW; r++) {
                    const uint32_t i_sk = row0 + warp + r * NW;
                    if (i_sk >= row_end) continue;
                    const uint64_t ib = d.chunk_off[(uint64_t)i_sk * (n_chunks + 1) + c];
                    const uint64_t ie = d.chunk_off[(uint64_t)i_sk * (n_chunks + 1) + c + 1];
                    uint32_t acc = 0;
                    // software pipeline: the next 32 row elements are requested before the current ones are probed
                    uint64_t n_lo = 0, n_hi = 0;
                    uint32_t n_mn = 0;
                    if (ib + lane < ie) {
                        n_lo = d.klo[ib + lane]; n_mn = d.minim[ib + lane];
                        if (HAS_HI) n_hi = d.khi[ib + lane];
                    }
                    for (uint64_t e0 = ib; e0 < ie; e0 += 32) {
                        const uint64_t lo = n_lo, hi = n_hi;
                        const uint32_t mn = n_mn;
                        const bool have = e0 + lane < ie;
                        const uint64_t en = e0 + 32 + lane;
                        if (en < ie) {
                            n_lo = d.klo[en]; n_mn = d.minim[en];
                            if (HAS_HI) n_hi = d.khi[en];
                        }
                        uint32_t mask = 0;
                        if (have) {
                            uint32_t h = elem_hash(lo, hi, mn);
                            for (;;) {
                                uint32_t o = s_slot[h];
                                if (o == 0) break;
                                o--;
                                if (s_klo[o] == lo && s_min[o] == mn && (!HAS_HI || s_khi[o] == hi)) {
                                    mask = s_mask[o];
                                    break;
                                }
                                h = (h + 1) & (CMP_SLOTS - 1);
                            }
                        }
                        // 32 lane masks -> per-column counts (lane jj keeps column jj)
                        uint32_t any = __reduce_or_sync(0xffffffffu, mask);
                        if (__popc(any) > 8) {
                            // many columns hit (related sketches): transpose the 32x32 bit matrix with five
                            // butterfly shuffles; lane jj ends up with bit jj of every lane's mask
                            uint32_t v = __brev(mask);
#pragma unroll
                            for (int j = 16, m = 0x0000FFFF; j; j >>= 1, m ^= m << j) {
                                const uint32_t x = __shfl_xor_sync(0xffffffffu, v, j);
                                if (lane & j) v ^= ((x ^ (v >> j)) & (uint32_t)m) << j;
                                else v ^= (v ^ (x >> j)) & (uint32_t)m;
                            }
                            acc += __popc(v);
                        } else {
                            while (any) {
                                int jj = __ffs(any) - 1;
                                any &= any - 1;
                                uint32_t bb = __ballot_sync(0xffffffffu, (mask >> jj) & 1u);
                                if (lane == jj) acc += __popc(bb);
                            }
                        }
                    }
                    // lane jj of the warp that owns the row holds the count of pair (row, col0+jj); a (tile, row) is
                    // owned by exactly one warp, so this is a plain read-modify-write
                    if (acc) s_acc[(t * 32 + (warp + r * NW)) * 32 + lane] += acc;
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_ib * 1024; i += CMP_THREADS) {
        const uint32_t v = s_acc[i];
        if (!v) continue;
        const uint32_t t = i >> 10, rr = (i >> 5) & 31, cc = i & 31;
        const uint32_t i_sk = row_begin + (ib0 + t) * 32 + rr, j_sk = col0 + cc;
        if (i_sk >= row_end || j_sk >= col_end || i_sk == j_sk) continue;
        if (compact) atomicAdd(out + ((uint64_t)blockIdx.x * RT + t) * 1024 + rr * 32 + cc, v);
        else atomicAdd(out + (uint64_t)(i_sk - row_begin) * ld + (j_sk - col_begin), v);
    }
}

// ---------------------------------------------------------------- multi-GPU plan
//
// After the all-gather every rank holds every rank's header {sketches, queries, elements, flags} and size list.
// One CTA turns them into: the global sketch order (all queries rank-major, then all references rank-major --
// Comparator.cpp:7-21, 512-515 put the queries first), their element ranges inside the gathered arrays (rank r's
// elements start at r * e_cap), the job's dimensions, and the units of the tile grid dealt to this rank.
__global__ void __launch_bounds__(256)
exchange_plan_kernel(const uint64_t *__restrict__ hdrsz, uint32_t world, uint32_t rank, uint64_t n_cap, uint64_t e_cap,
                     int symmetric, uint32_t rt, uint32_t units_cap, uint64_t *__restrict__ sk_begin, uint64_t *__restrict__ sk_end,
                     uint64_t *__restrict__ sizes_compact, uint2 *__restrict__ units, uint32_t *__restrict__ dims)
{
    __shared__ uint64_t s_qoff[CMP_MAX_WORLD + 1], s_roff[CMP_MAX_WORLD + 1];
    const uint64_t stride = XHDR_WORDS + n_cap;
    if (threadIdx.x == 0) {
        uint64_t q = 0, r = 0;
        for (uint32_t w = 0; w < world; w++) {
            const uint64_t n_w = hdrsz[w * stride], q_w = hdrsz[w * stride + 1];
            s_qoff[w] = q; s_roff[w] = r;
            q += q_w; r += n_w - q_w;
        }
        s_qoff[world] = q; s_roff[world] = r;
    }
    __syncthreads();
    const uint64_t Q = s_qoff[world], N = Q + s_roff[world];
    for (uint32_t w = 0; w < world; w++) {
        const uint64_t *h = hdrsz + w * stride;
        const uint64_t n_w = h[0], q_w = h[1];
        // element offsets of rank w's sketches: exclusive scan of its sizes, done by thread 0 of each pass (n_w is
        // small next to the join) -- every thread then places its sketches
        __shared__ uint64_t s_carry;
        if (threadIdx.x == 0) s_carry = (uint64_t)w * e_cap;
        __syncthreads();
        for (uint64_t b0 = 0; b0 < n_w; b0 += blockDim.x) {
            const uint64_t j = b0 + threadIdx.x;
            const uint64_t sz = j < n_w ? h[XHDR_WORDS + j] : 0;
            // block-wide exclusive scan of sz
            __shared__ uint64_t s_w[8];
            uint64_t inc = sz;
            const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) s_w[wp] = inc;
            __syncthreads();
            uint64_t wb = 0, tot = 0;
            for (int x = 0; x < 8; x++) { if (x < wp) wb += s_w[x]; tot += s_w[x]; }
            const uint64_t beg = s_carry + wb + inc - sz;
            if (j < n_w) {
                const uint64_t gidx = j < q_w ? s_qoff[w] + j : Q + s_roff[w] + (j - q_w);
                sk_begin[gidx] = beg; sk_end[gidx] = beg + sz; sizes_compact[gidx] = sz;
            }
            __syncthreads();
            if (threadIdx.x == 0) s_carry += tot;
            __syncthreads();
        }
    }
    // units of the tile grid: column tile jb, row tiles in runs of rt; symmetric: row tiles 0 .. jb only
    const uint32_t rows = symmetric ? (uint32_t)N : (uint32_t)Q;
    const uint32_t nI = (rows + 31) / 32, nJ = ((uint32_t)N + 31) / 32;
    __shared__ uint32_t s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    // unit index = running number over (jb, run); thread jb computes the first index of its column tile in closed form
    for (uint32_t jb = threadIdx.x; jb < nJ; jb += blockDim.x) {
        // runs of column tile x: symmetric ceil((x + 1) / rt), else ceil(nI / rt)
        uint64_t first;
        if (symmetric) {
            // sum_{x < jb} ceil((x + 1) / rt) = sum_{y = 1..jb} ceil(y / rt)
            const uint64_t full = jb / rt, rem = jb % rt;
            first = rt * full * (full + 1) / 2 + rem * (full + 1);
        } else {
            first = (uint64_t)jb * ((nI + rt - 1) / rt);
        }
        const uint32_t lim = symmetric ? min(jb + 1, nI) : nI;
        uint32_t mine = 0;
        for (uint32_t ib0 = 0, u = 0; ib0 < lim; ib0 += rt, u++) {
            const uint64_t idx = first + u;
            if (idx % world != rank) continue;
            const uint64_t slot = idx / world;
            if (slot < units_cap) units[slot] = make_uint2(jb | (ib0 << 16), min(rt, lim - ib0));
            mine++;
        }
        if (mine) atomicAdd(&s_count, mine);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        dims[0] = rows; dims[1] = (uint32_t)N; dims[2] = min(s_count, units_cap); dims[3] = (uint32_t)Q;
        dims[4] = s_count;                                       // > units_cap: the host's capacity bound was wrong
    }
}

cudaError_t launch_exchange_plan(const uint64_t *d_hdrsz, uint32_t world, uint32_t rank, uint64_t n_cap, uint64_t e_cap,
                                 int symmetric, uint32_t rt, uint32_t units_cap, uint64_t *sk_begin, uint64_t *sk_end,
                                 uint64_t *sizes_compact, uint2 *units, uint32_t *dims, cudaStream_t st)
{
    if (world > CMP_MAX_WORLD) return cudaErrorInvalidValue;
    exchange_plan_kernel<<<1, 256, 0, st>>>(d_hdrsz, world, rank, n_cap, e_cap, symmetric, rt, units_cap, sk_begin, sk_end,
                                            sizes_compact, units, dims);
    return cudaGetLastError();
}

// Rank 0, after the owned tiles of every rank have arrived: one CTA per unit (column tile x run of row tiles)
// copies its compact 32 x 32 tiles to their place in the rows x columns matrix (leading dimension ld, zeroed by
// the caller).  Same unit enumeration as exchange_plan_kernel: unit idx belongs to rank idx % world, slot idx / world.
__global__ void __launch_bounds__(256)
exchange_assemble_kernel(const uint32_t *__restrict__ tiles, uint64_t tile_words, const uint32_t *__restrict__ dims,
                         uint32_t world, int symmetric, uint32_t rt, uint32_t rt_max, uint32_t *__restrict__ out, uint64_t ld)
{
    const uint32_t rows = dims[0], N = dims[1];
    const uint32_t nI = (rows + 31) / 32, nJ = (N + 31) / 32;
    const uint32_t jb = blockIdx.x, run = blockIdx.y;
    if (jb >= nJ) return;
    const uint32_t lim = symmetric ? min(jb + 1, nI) : nI;
    const uint32_t ib0 = run * rt;
    if (ib0 >= lim) return;
    uint64_t first;
    if (symmetric) {
        const uint64_t full = jb / rt, rem = jb % rt;
        first = rt * full * (full + 1) / 2 + rem * (full + 1);
    } else {
        first = (uint64_t)jb * ((nI + rt - 1) / rt);
    }
    const uint64_t idx = first + run;
    const uint32_t *src = tiles + (idx % world) * tile_words + (idx / world) * (uint64_t)rt_max * 1024;
    const uint32_t n_ib = min(rt, lim - ib0);
    for (uint32_t i = threadIdx.x; i < n_ib * 1024; i += blockDim.x) {
        const uint32_t t = i >> 10, rr = (i >> 5) & 31, cc = i & 31;
        const uint32_t r = (ib0 + t) * 32 + rr, c = jb * 32 + cc;
        if (r < rows && c < N) out[(uint64_t)r * ld + c] = src[i];
    }
}

cudaError_t launch_exchange_assemble(const uint32_t *d_tiles, uint64_t tile_words, const uint32_t *dev_dims, uint32_t world,
                                     int symmetric, uint32_t rt, uint32_t rt_max, uint32_t rows_cap, uint32_t cols_cap,
                                     uint32_t *d_out, uint64_t ld, cudaStream_t st)
{
    const uint32_t nJ = (cols_cap + 31) / 32, nI = (rows_cap + 31) / 32;
    const uint32_t runs = (nI + rt - 1) / rt;
    if (!nJ || !runs) return cudaSuccess;
    if (runs > 65535) return cudaErrorInvalidValue;
    exchange_assemble_kernel<<<dim3(nJ, runs), 256, 0, st>>>(d_tiles, tile_words, dev_dims, world, symmetric, rt, rt_max, d_out, ld);
    return cudaGetLastError();
}

cudaError_t launch_chunk_offsets(const CmpData &d, uint32_t n_sketches, const uint32_t *dev_dims, uint32_t n_chunks, int m,
                                 cudaStream_t st)
{
    uint64_t total = (uint64_t)n_sketches * (n_chunks + 1);          // n_sketches = capacity when dev_dims is given
    if (!total) return cudaSuccess;
    unsigned blocks = (unsigned)((total + 255) / 256);
    chunk_offsets_kernel<<<blocks, 256, 0, st>>>(d.minim, d.sk_off, d.sk_end, n_sketches, dev_dims, n_chunks, 1ULL << (2 * m),
                                                 d.chunk_off);
    return cudaGetLastError();
}

size_t hashjoin_smem_bytes(bool has_hi)
{
    return (size_t)CMP_CAP * (has_hi ? 16 : 8) + (size_t)CMP_CAP * 8 + (size_t)CMP_SLOTS * 4 +
           (size_t)(has_hi ? CMP_RT_HI : CMP_RT) * 1024 * 4;
}

cudaError_t launch_hashjoin(const CmpData &d, bool has_hi, const uint2 *d_units, uint32_t n_units, const uint32_t *dev_dims,
                            uint32_t n_chunks, uint32_t chunk_groups, uint32_t row_begin, uint32_t row_end,
                            uint32_t col_begin, uint32_t col_end, uint32_t *d_out, uint64_t ld, cudaStream_t st)
{
    if (!n_units) return cudaSuccess;
    size_t smem = hashjoin_smem_bytes(has_hi);
    {
        static PerDeviceOnce once[2];
        cudaError_t e = once[has_hi].run([&] {
            return has_hi
                ? cudaFuncSetAttribute(hashjoin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                : cudaFuncSetAttribute(hashjoin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        });
        if (e != cudaSuccess) return e;
    }
    dim3 grid(n_units, chunk_groups);
    if (has_hi)
        hashjoin_kernel<true><<<grid, CMP_THREADS, smem, st>>>(d, d_units, dev_dims, n_chunks, row_begin, row_end, col_begin,
                                                              col_end, d_out, ld);
    else
        hashjoin_kernel<false><<<grid, CMP_THREADS, smem, st>>>(d, d_units, dev_dims, n_chunks, row_begin, row_end, col_begin,
                                                               col_end, d_out, ld);
    return cudaGetLastError();
}

}  // namespace spsp
