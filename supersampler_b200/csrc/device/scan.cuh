// Declarations shared by scan.cu and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../../include/spsp.h"

namespace spsp {

struct ScanOut {
    spsp_hit *hits;              // device, cap entries
    unsigned long long *count;   // device, total hits found (may exceed cap)
    unsigned long long cap;
};

// Aligned q-gram filter: probe every g-th position with the q-gram that
// starts there; bits = log2(table bits); hashed != 0 when 2q > bits;
// rep_log2 = log2 of the number of interleaved table copies in shared memory
// (copy = lane % R, so lanes of different copies never share a bank).
// kind 0: bit table (any key width, direct or hashed); kind 1: byte table of
// phase masks (g == 4, 2q <= 16, one byte per q-gram); kind 2 (m == 11 only):
// bank-private bit table over 15 bits of the aligned 2-byte key (every lane
// reads its own copy: one shared-memory wavefront per probe) + a hashed bit
// table of the selected forward m-mers with 2^hbits bits, see scan_rowbit_kernel.
struct FilterParams {
    int g, q, bits, hashed, rep_log2, kind, hbits;
};

constexpr int FILTER_MAX_BITS = 20;          // 128 KB of shared memory
constexpr size_t FILTER_MAX_SMEM = (size_t)1 << (FILTER_MAX_BITS - 3);   // table incl. replication
constexpr int FILTER_CH = 4;                 // 64-base chunks per lane per iteration (8192 bases per warp)
constexpr int FILTER_WQ = 64;                // queue entries (3 words each) per warp

constexpr int ROWBIT_ROWS = 1024;            // kind 2: 10 row bits + 5 bit-in-word bits of the 16-bit key (bit 5 dropped)
constexpr int ROWBIT_Q = 64;                 // ring entries per warp
constexpr int ROWBIT_EW = 7;                 // words per entry: chunk, flags, 5 sequence words
constexpr int ROWBIT_MIN_HBITS = 12, ROWBIT_MAX_HBITS = 17;   // level-2 table: 2^hbits bits (512 B .. 16 KB)

// Global image of the tables.  kind 2: [level-2 bit table, 2^hbits bits][compact level-1 bit table, 4 KB].
__host__ __device__ __forceinline__ size_t filter_table_bytes(const FilterParams &fp)
{
    if (fp.kind == 2) return ((size_t)1 << (fp.hbits - 3)) + (size_t)ROWBIT_ROWS * 4;
    return fp.kind == 1 ? ((size_t)1 << (2 * fp.q)) : ((size_t)1 << (fp.bits - 3));
}

cudaError_t launch_scan_dense(const uint32_t *d_packed, uint64_t n_bases, int m, uint64_t thr, ScanOut out,
                              cudaStream_t st);
// d_exact: bitmap over the 4^m forward m-mers (bit x set iff m-mer x is selected).
cudaError_t launch_filter_build(int m, uint64_t thr, FilterParams fp, uint32_t *d_table, uint32_t *d_exact,
                                unsigned long long *d_nsel, cudaStream_t st);
cudaError_t launch_scan_filter(const uint32_t *d_packed, uint64_t n_bases, int m, uint64_t thr, FilterParams fp,
                               const uint32_t *d_table, const uint32_t *d_exact, ScanOut out, cudaStream_t st);

}  // namespace spsp
