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
// minimizers are hash-uniform so equal ranges balance).  A CTA owns one tile
// of 32 row sketches x 32 column sketches and a subset of the chunks.  Per
// chunk it builds, in shared memory, a hash table of the column sketches'
// elements (key -> 32-bit membership mask over the tile's columns), then every
// warp streams the elements of its row sketches through the table and turns
// the 32 lane masks into 32 per-column counts with warp ballots -- one probe
// serves 32 sketch pairs.  Column batches larger than the table are processed
// in several passes, so capacity never affects the result.
#include "common.cuh"
#include "compare.cuh"

namespace spsp {

__global__ void chunk_offsets_kernel(const uint32_t *__restrict__ minim, const uint64_t *__restrict__ sk_off,
                                     const uint64_t *__restrict__ sk_end, uint32_t n_sketches, uint32_t n_chunks, uint64_t space,
                                     uint64_t *__restrict__ chunk_off)
{
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

template <bool HAS_HI>
__global__ void __launch_bounds__(CMP_THREADS, 1)
hashjoin_kernel(CmpData d, const uint2 *__restrict__ tiles, uint32_t n_chunks, uint32_t row_begin,
                uint32_t row_end, uint32_t col_begin, uint32_t col_end, uint32_t *__restrict__ out, uint64_t ld)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *s_klo = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *s_khi = s_klo + CMP_CAP;                              // only touched when HAS_HI
    uint32_t *s_min = reinterpret_cast<uint32_t *>(s_klo + (HAS_HI ? 2 : 1) * CMP_CAP);
    uint32_t *s_mask = s_min + CMP_CAP;
    uint32_t *s_slot = s_mask + CMP_CAP;
    __shared__ uint64_t s_jbeg[32];
    __shared__ uint32_t s_pref[33];

    const uint2 tile = tiles[blockIdx.x];
    const uint32_t row0 = row_begin + tile.x * 32, col0 = col_begin + tile.y * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = CMP_THREADS / 32, RPW = 32 / NW;
    uint32_t acc[RPW];
#pragma unroll
    for (int r = 0; r < RPW; r++) acc[r] = 0;

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
            // probe: warp w streams rows w, w+NW, ...
#pragma unroll
            for (int r = 0; r < RPW; r++) {
                const uint32_t i_sk = row0 + warp + r * NW;
                if (i_sk >= row_end) continue;
                const uint64_t ib = d.chunk_off[(uint64_t)i_sk * (n_chunks + 1) + c];
                const uint64_t ie = d.chunk_off[(uint64_t)i_sk * (n_chunks + 1) + c + 1];
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
                        acc[r] += __popc(v);
                    } else {
                        while (any) {
                            int jj = __ffs(any) - 1;
                            any &= any - 1;
                            uint32_t bb = __ballot_sync(0xffffffffu, (mask >> jj) & 1u);
                            if (lane == jj) acc[r] += __popc(bb);
                        }
                    }
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }
    // lane jj of the warp that owns row r holds the count of pair (row, col0+jj)
#pragma unroll
    for (int r = 0; r < RPW; r++) {
        const uint32_t i_sk = row0 + warp + r * NW, j_sk = col0 + lane;
        if (i_sk < row_end && j_sk < col_end && i_sk != j_sk && acc[r])
            atomicAdd(out + (uint64_t)(i_sk - row_begin) * ld + (j_sk - col_begin), acc[r]);
    }
}

// One thread per rank: hdr_all[2r] = sketches of rank r, sizes_all[r * n_max + i] = elements of its sketch i,
// which start at element r * e_max of the gathered arrays.
__global__ void gathered_ranges_kernel(const uint64_t *__restrict__ hdr_all, const uint64_t *__restrict__ sizes_all,
                                       uint32_t world, uint64_t n_max, uint64_t e_max, uint64_t *__restrict__ sk_begin,
                                       uint64_t *__restrict__ sk_end, uint64_t *__restrict__ sizes_compact)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= world) return;
    uint64_t first = 0;
    for (uint32_t q = 0; q < r; q++) first += hdr_all[2 * q];
    uint64_t acc = r * e_max;
    const uint64_t n = hdr_all[2 * r];
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t sz = sizes_all[r * n_max + i];
        sk_begin[first + i] = acc;
        sk_end[first + i] = acc + sz;
        sizes_compact[first + i] = sz;
        acc += sz;
    }
}

cudaError_t launch_gathered_ranges(const uint64_t *d_hdr_all, const uint64_t *d_sizes_all, uint32_t world, uint64_t n_max,
                                   uint64_t e_max, uint64_t *sk_begin, uint64_t *sk_end, uint64_t *sizes_compact,
                                   cudaStream_t st)
{
    gathered_ranges_kernel<<<1, 32, 0, st>>>(d_hdr_all, d_sizes_all, world, n_max, e_max, sk_begin, sk_end, sizes_compact);
    return cudaGetLastError();
}

cudaError_t launch_chunk_offsets(const CmpData &d, uint32_t n_sketches, uint32_t n_chunks, int m, cudaStream_t st)
{
    uint64_t total = (uint64_t)n_sketches * (n_chunks + 1);
    if (!total) return cudaSuccess;
    unsigned blocks = (unsigned)((total + 255) / 256);
    chunk_offsets_kernel<<<blocks, 256, 0, st>>>(d.minim, d.sk_off, d.sk_end, n_sketches, n_chunks, 1ULL << (2 * m),
                                                 d.chunk_off);
    return cudaGetLastError();
}

size_t hashjoin_smem_bytes(bool has_hi)
{
    return (size_t)CMP_CAP * (has_hi ? 16 : 8) + (size_t)CMP_CAP * 8 + (size_t)CMP_SLOTS * 4;
}

cudaError_t launch_hashjoin(const CmpData &d, bool has_hi, const uint2 *d_tiles, uint32_t n_tiles,
                            uint32_t n_chunks, uint32_t chunk_groups, uint32_t row_begin, uint32_t row_end,
                            uint32_t col_begin, uint32_t col_end, uint32_t *d_out, uint64_t ld, cudaStream_t st)
{
    if (!n_tiles) return cudaSuccess;
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
    dim3 grid(n_tiles, chunk_groups);
    if (has_hi)
        hashjoin_kernel<true><<<grid, CMP_THREADS, smem, st>>>(d, d_tiles, n_chunks, row_begin, row_end, col_begin,
                                                              col_end, d_out, ld);
    else
        hashjoin_kernel<false><<<grid, CMP_THREADS, smem, st>>>(d, d_tiles, n_chunks, row_begin, row_end, col_begin,
                                                               col_end, d_out, ld);
    return cudaGetLastError();
}

}  // namespace spsp
