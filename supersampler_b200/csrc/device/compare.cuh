// Declarations shared by compare.cu and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spsp {

struct CmpData {
    const uint32_t *minim;      // [E]
    const uint64_t *klo;        // [E]
    const uint64_t *khi;        // [E] or nullptr (k <= 32)
    const uint64_t *sk_off;     // [N+1] element range of sketch s = [sk_off[s], sk_off[s+1]) ...
    const uint64_t *sk_end;     // ... or [sk_off[s], sk_end[s]) when given (gathered layout with padding between ranks)
    uint64_t *chunk_off;        // [N][C+1]
};

constexpr int CMP_THREADS = 512;      // 16 warps, 2 row sketches per warp
constexpr int CMP_CAP = 6144;         // column elements per hash-table pass
constexpr int CMP_SLOTS = 16384;      // open-addressing slots (load <= 0.375)
constexpr int CMP_RT = 8;             // row tiles that share one table build (k <= 32)
constexpr int CMP_RT_HI = 4;          // ... when the 128-bit keys take the shared memory (k > 32)
constexpr int CMP_MAX_WORLD = 64;     // ranks of a multi-GPU exchange
constexpr int XHDR_WORDS = 4;         // exchange header per rank: {sketches, queries, elements, flags}, then the size list

// Multi-GPU plan (after the all-gather): global sketch order (queries of all ranks, then references), element
// ranges inside the gathered arrays, job dimensions dims = {rows, columns, units of this rank, queries, units
// before capping} and this rank's units of the tile grid.
cudaError_t launch_exchange_plan(const uint64_t *d_hdrsz, uint32_t world, uint32_t rank, uint64_t n_cap, uint64_t e_cap,
                                 int symmetric, uint32_t rt, uint32_t units_cap, uint64_t *sk_begin, uint64_t *sk_end,
                                 uint64_t *sizes_compact, uint2 *units, uint32_t *dims, cudaStream_t st);
// Rank 0: the compact tiles of every rank ([world][tile_words], unit idx at rank idx % world, slot idx / world, rt_max
// tile slots per unit) -> the rows x columns matrix d_out (zeroed by the caller, leading dimension ld); grid sized
// from the capacities, dimensions read from dev_dims.
cudaError_t launch_exchange_assemble(const uint32_t *d_tiles, uint64_t tile_words, const uint32_t *dev_dims, uint32_t world,
                                     int symmetric, uint32_t rt, uint32_t rt_max, uint32_t rows_cap, uint32_t cols_cap,
                                     uint32_t *d_out, uint64_t ld, cudaStream_t st);
// n_sketches is the capacity when dev_dims (device: dims[1] = sketches) is given.
cudaError_t launch_chunk_offsets(const CmpData &d, uint32_t n_sketches, const uint32_t *dev_dims, uint32_t n_chunks, int m,
                                 cudaStream_t st);
size_t hashjoin_smem_bytes(bool has_hi);
// units[u] = {jb | ib0 << 16, n_ib}; grid = n_units x chunk_groups.  dev_dims != null: device-side dimensions and
// compact per-unit output tiles (see compare.cu).
cudaError_t launch_hashjoin(const CmpData &d, bool has_hi, const uint2 *d_units, uint32_t n_units, const uint32_t *dev_dims,
                            uint32_t n_chunks, uint32_t chunk_groups, uint32_t row_begin, uint32_t row_end,
                            uint32_t col_begin, uint32_t col_end, uint32_t *d_out, uint64_t ld, cudaStream_t st);

}  // namespace spsp
