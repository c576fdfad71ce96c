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
// starts there; bits = log2(table bits); hashed != 0 when 2q > bits.
struct FilterParams {
    int g, q, bits, hashed;
};

constexpr int FILTER_MAX_BITS = 20;          // 128 KB of shared memory
constexpr int FILTER_THREADS = 1024;         // upper bound (launch bounds)
constexpr int FILTER_QUEUE = 4096;           // candidate queue entries per CTA

__host__ __device__ __forceinline__ uint32_t filter_index(uint32_t key, const FilterParams &fp)
{
    return fp.hashed ? (key * 0x9E3779B1u) >> (32 - fp.bits) : key;
}

cudaError_t launch_scan_dense(const uint32_t *d_packed, uint64_t n_bases, int m, uint64_t thr, ScanOut out,
                              cudaStream_t st);
cudaError_t launch_filter_build(int m, uint64_t thr, FilterParams fp, uint32_t *d_table,
                                unsigned long long *d_nsel, cudaStream_t st);
cudaError_t launch_scan_filter(const uint32_t *d_packed, uint64_t n_bases, int m, uint64_t thr, FilterParams fp,
                               const uint32_t *d_table, ScanOut out, cudaStream_t st);

}  // namespace spsp
