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

// sk_begin/sk_end of the union of all ranks' sketches from the gathered per-rank size lists
// (sizes_all[r * n_max + i], elements of rank r start at r * e_max); also the compact size list.
cudaError_t launch_gathered_ranges(const uint64_t *d_hdr_all, const uint64_t *d_sizes_all, uint32_t world, uint64_t n_max,
                                   uint64_t e_max, uint64_t *sk_begin, uint64_t *sk_end, uint64_t *sizes_compact,
                                   cudaStream_t st);
cudaError_t launch_chunk_offsets(const CmpData &d, uint32_t n_sketches, uint32_t n_chunks, int m, cudaStream_t st);
size_t hashjoin_smem_bytes(bool has_hi);
cudaError_t launch_hashjoin(const CmpData &d, bool has_hi, const uint2 *d_tiles, uint32_t n_tiles,
                            uint32_t n_chunks, uint32_t chunk_groups, uint32_t row_begin, uint32_t row_end,
                            uint32_t col_begin, uint32_t col_end, uint32_t *d_out, uint64_t ld, cudaStream_t st);

}  // namespace spsp
