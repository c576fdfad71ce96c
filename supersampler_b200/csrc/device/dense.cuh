// Dense minimizer machine (see dense.cu): totals of the reference's per-base
// loop that depend on every position, not only on the selected ones.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spsp {

struct DenseBuffers;      // grow-only device scratch, owned by the slot
DenseBuffers *dense_buffers_create();
void dense_buffers_destroy(DenseBuffers *b);

struct DenseIn {
    const uint32_t *d_packed;        // 2-bit sequence of the whole batch
    uint64_t n_bases;
    const uint64_t *d_rec_begin;     // [n_rec] global base offsets, ascending, disjoint
    const uint64_t *d_rec_end;
    const uint32_t *d_rec_input;
    uint64_t n_rec;
    uint32_t n_inputs;
    int k, m;
    uint64_t thr;
};

// h_total_superkmers / h_selected_kmers: host arrays of n_inputs entries.
// Runs on `st` and synchronises it.  kernel_ms (may be null): CUDA-event time of the kernels.
cudaError_t dense_stats_run(DenseBuffers *b, const DenseIn &in, uint64_t *h_total_superkmers,
                            uint64_t *h_selected_kmers, float *kernel_ms, uint32_t *launched, cudaStream_t st);

}  // namespace spsp
