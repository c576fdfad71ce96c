// Device post-pass of the sketch stage: from the scan kernel's hit list to
// (a) the sketch bytes of every input of a batch and (b) the compare-stage
// elements, without leaving the GPU.  Mirrors csrc/host/postpass.cpp step by
// step (same reference citations); the host version stays as the exact
// cross-check in the tests.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/spsp.h"

namespace spsp {

struct PostpassBuffers;      // grow-only device scratch, owned by the context

struct PostpassIn {
    const uint32_t *d_packed;        // 2-bit sequence of the whole batch
    uint64_t n_bases;
    const spsp_hit *d_hits;          // scan output (unordered), hits_cap entries
    const unsigned long long *d_hit_count;   // device: hits the scan found (may exceed hits_cap: then nothing is used)
    uint64_t hits_cap;
    const uint64_t *d_rec_begin;     // [n_rec] global base offsets, ascending
    const uint64_t *d_rec_end;       // [n_rec]
    const uint32_t *d_rec_input;     // [n_rec] input (file) of each record
    uint64_t n_rec;
    uint32_t n_inputs;
    int k, m;
    unsigned abundance;
};

struct PostpassOut {
    // host-visible (pinned) results, valid until the next run on the same buffers
    const uint8_t *h_body;           // sketch bytes after the header line, all inputs back to back
    const uint64_t *h_body_off;      // [n_inputs + 1] byte range of each input inside h_body
    const uint64_t *h_selected;      // [n_inputs] selected k-mer occurrences (header field 3)
    const uint64_t *h_elem_off;      // [n_inputs + 1] compare elements of each input
    // device-resident compare elements (sketch -> compare hand-off)
    const uint32_t *d_minim;
    const uint64_t *d_klo, *d_khi;   // d_khi null when k <= 32
    uint64_t n_elems;
    uint64_t n_hits;                 // what the scan counted
    uint32_t kernels_launched;       // kernels of this file (cub::DeviceRadixSort's, large batches only, are not counted)
    int retry;                       // PP_RETRY_*: a capacity was too small, nothing above is valid, run again
};
enum { PP_RETRY_NONE = 0, PP_RETRY_HITS = 1 /* grow the hit buffer, rescan */, PP_RETRY_POSTPASS = 2 /* rerun the post-pass */ };

PostpassBuffers *postpass_buffers_create();
void postpass_buffers_destroy(PostpassBuffers *b);
// Enqueues the whole pass on `st` behind the scan (no host round trip in between: the hit count stays on the
// device), then synchronises once and delivers the results.  Returns cudaSuccess or the first error;
// out->retry != 0 asks the caller to run again (capacities are remembered in the buffers).
cudaError_t postpass_run(PostpassBuffers *b, const PostpassIn &in, PostpassOut *out, cudaStream_t st);

}  // namespace spsp
