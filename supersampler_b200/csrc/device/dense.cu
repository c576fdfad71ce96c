// Dense minimizer machine for sm_100a: the part of the reference's per-base
// loop (SubSampler.cpp:367-440) that the sparse hit path does not need -- the
// minimizer of EVERY k-mer -- restated so that it runs in parallel:
//
//   rows kernel      rolling canonical m-mer hash at every position (one position
//                    per lane, 32 per warp row) and the sliding-window minimum of
//                    the hashes over the previous k-m and k-m+1 positions, by
//                    log-step warp shuffles (windows of 2^t doubled t times, two
//                    overlapping power-of-two windows combined).  Two bits per
//                    position come out: "new minimum" (h[p] < min of the previous
//                    k-mer's k-m+1 m-mers: the branch SubSampler.cpp:374-388 takes)
//                    and "selected" (the k-mer ending here has a minimizer hash <= T:
//                    SURVEY App. A.2 closed form, the FracMinHash 1/s test).
//   segments kernel  a "new minimum" resets the machine's state to a value that
//                    depends on that position only (minimizer = the entering m-mer,
//                    position_min = its position), so the sequence splits into
//                    independent segments; each is replayed from its reset point:
//                    only rescans (regular_minimizer_pos, :81-169, with its position
//                    quirks) can happen inside, and each rescan ends a super-k-mer
//                    ("dump", :391-398 / :401).  Counting resets + rescans + one per
//                    record gives total_superkmer_number (:429-431, :452) exactly.
//
// Used for print_stat's totals and as an independent check of the hit path: the
// number of selected k-mers must equal the sketch header's third field.
#include "dense.cuh"

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace spsp {

struct DBufD {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 1024;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
    ~DBufD() { if (p) cudaFree(p); }
};

struct DenseBuffers {
    DBufD nm, sel, totals;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ~DenseBuffers()
    {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
};
DenseBuffers *dense_buffers_create() { return new DenseBuffers(); }
void dense_buffers_destroy(DenseBuffers *b) { delete b; }

constexpr int DN_ROWS = 128;          // rows (of 32 positions) per warp tile
constexpr int DN_WARM = 2;            // warm-up rows: windows reach back at most 61 positions

__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// Value of a row-distributed array `sh` (1..32) positions earlier: from this row or the previous one.  The
// sending lane knows which of the two its receiver wants, so one shuffle moves the right value.
__device__ __forceinline__ uint64_t shift_down(uint64_t cur, uint64_t prev, int sh, int lane)
{
    const uint64_t send = lane + sh < 32 ? cur : prev;       // receiver (lane + sh) & 31 is in this row or wrapped
    return __shfl_sync(0xffffffffu, send, (lane - sh) & 31);
}

// Canonical m-mer hash at global position p (positions past the end hash to +inf).
__device__ __forceinline__ uint64_t mmer_hash_at(const uint32_t *__restrict__ packed, long long p, uint64_t n_pos, int m)
{
    if (p < 0 || (uint64_t)p >= n_pos) return ~0ULL;
    const uint64_t w = (uint64_t)p >> 4;
    const uint32_t fw = window16(__ldg(packed + w), __ldg(packed + w + 1), (int)(p & 15)) >> (32 - 2 * m);
    const uint32_t rc = rc_mmer(fw, m);
    return xxh64_8(min(fw, rc));
}

__global__ void __launch_bounds__(256)
dense_rows_kernel(const uint32_t *__restrict__ packed, uint64_t n_bases, int k, int m, uint64_t thr,
                  uint32_t *__restrict__ nm_bits, uint32_t *__restrict__ sel_bits)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_pos = n_bases >= (uint64_t)m ? n_bases - m + 1 : 0;
    const uint64_t n_rows = (n_pos + 31) >> 5;
    const uint64_t row0 = warp * DN_ROWS;
    if (row0 >= n_rows) return;
    const uint64_t row1 = min(row0 + (uint64_t)DN_ROWS, n_rows);
    const int d = k - m;
    int T = 0;
    while ((2 << T) <= d) T++;                               // 2^T <= d < 2^(T+1)
    const int sA = 1 + d - (1 << T), sB = sA + 1;            // second window of size d / d+1, seen from p-1
    uint64_t prevA[6];                                       // previous row of every doubling level
#pragma unroll
    for (int t = 0; t < 6; t++) prevA[t] = ~0ULL;
    for (long long row = (long long)row0 - DN_WARM; row < (long long)row1; row++) {
        const long long p = row * 32 + lane;
        const uint64_t h = mmer_hash_at(packed, p, n_pos, m);
        // A_t[i] = min of the 2^t hashes ending at i
        uint64_t A = h;
#pragma unroll
        for (int t = 0; t < 5; t++) {
            if (t < T) {
                const uint64_t old = prevA[t];
                prevA[t] = A;
                A = umin64(A, shift_down(A, old, 1 << t, lane));
            }
        }
        const uint64_t oldT = prevA[5];
        prevA[5] = A;
        const uint64_t e1 = shift_down(A, oldT, 1, lane);                  // window ending at p-1
        const uint64_t md = umin64(e1, shift_down(A, oldT, sA, lane));     // min h[p-d .. p-1]
        const uint64_t mw = umin64(e1, shift_down(A, oldT, sB, lane));     // min h[p-d-1 .. p-1]
        const unsigned nm = __ballot_sync(0xffffffffu, h < mw);
        const unsigned sl = __ballot_sync(0xffffffffu, umin64(h, md) <= thr);
        if (row >= (long long)row0 && lane == 0) { nm_bits[row] = nm; sel_bits[row] = sl; }
    }
}

// regular_minimizer_pos (SubSampler.cpp:81-169) on the k-mer that starts at global base `g`:
// right-most m-mer first, strict '<' on the hash, position quirks kept (:88-93, :149-164).
__device__ uint64_t dense_rescan(const uint32_t *__restrict__ packed, uint64_t g, int d, int m)
{
    uint64_t hbest = 0, pos = 0;
    uint32_t best = 0;
    bool rev = false;
    for (int j = 0; j <= d; j++) {
        const uint64_t p = g + (uint64_t)(d - j);
        const uint64_t w = p >> 4;
        const uint32_t fw = window16(__ldg(packed + w), __ldg(packed + w + 1), (int)(p & 15)) >> (32 - 2 * m);
        const uint32_t rc = rc_mmer(fw, m);
        const uint32_t cn = min(fw, rc);
        const bool lrev = cn != fw;
        const uint64_t h = xxh64_8(cn);
        if (j == 0) {
            best = cn; rev = lrev; hbest = h;
            pos = lrev ? 0 : (uint64_t)d;
        } else if (hbest > h) {
            pos = (uint64_t)(d - j); best = cn; rev = lrev; hbest = h;
        } else if (cn == best && lrev == rev) {
            if (rev && pos > (uint64_t)j) pos = (uint64_t)j;
            if (!rev && pos > (uint64_t)(d - j)) pos = (uint64_t)(d - j);
        }
    }
    return pos;
}

// Relative position of the next "new minimum" iteration after bit `p_rel` of record [rb, re), or `limit` if none
// (positions are m-mer starts relative to rb; valid iterations enter m-mers d+1 .. n-m).
__device__ uint64_t next_flag(const uint32_t *__restrict__ nm_bits, uint64_t rb, uint64_t from_rel, uint64_t last_rel,
                              uint64_t none)
{
    if (from_rel > last_rel) return none;
    uint64_t g = rb + from_rel;
    const uint64_t g_last = rb + last_rel;
    uint64_t w = g >> 5;
    uint32_t bits = nm_bits[w] & (0xFFFFFFFFu << (g & 31));
    for (;;) {
        if (bits) {
            const uint64_t hit = (w << 5) + (uint64_t)(__ffs(bits) - 1);
            return hit <= g_last ? hit - rb : none;
        }
        w++;
        if ((w << 5) > g_last) return none;
        bits = nm_bits[w];
    }
}

// Rescans between a state with position_min = pm (relative k-mer index space) and iteration c_end (exclusive).
__device__ uint32_t replay_rescans(const uint32_t *__restrict__ packed, uint64_t rb, uint64_t pm, uint64_t c_end, int d, int m)
{
    uint32_t n = 0;
    for (uint64_t c = pm + 1; c < c_end; c = pm + 1) {       // iteration c rescans when c-1 >= position_min (:391)
        pm = c + dense_rescan(packed, rb + c, d, m);          // position_min += i + 1 (:397)
        n++;
    }
    return n;
}

// regular_minimizer_pos by a whole warp: lane l hashes m-mer j = l (and j = l + 32 when d >= 32) of the k-mer
// at global base g; warp reductions pick the winner of the right-to-left strict-'<' scan (smallest j among the
// minimal hashes) and apply the position quirks in closed form: with the winner at j0 in orientation r, start
// from (j0 == 0 ? (r ? 0 : d) : d - j0) and, over the later m-mers j > j0 with the same canonical value and
// orientation, forward takes d - max j, reverse takes min(start, min j)  (:88-93, :149-164).
__device__ __forceinline__ uint64_t warp_rescan(const uint32_t *__restrict__ packed, uint64_t g, int d, int m, int lane)
{
    uint64_t h[2] = {~0ULL, ~0ULL};
    uint32_t cn[2] = {0, 0};
    bool rv[2] = {false, false};
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int j = lane + 32 * s;
        if (j <= d && (s == 0 || d >= 32)) {
            const uint64_t p = g + (uint64_t)(d - j);
            const uint64_t w = p >> 4;
            const uint32_t fw = window16(__ldg(packed + w), __ldg(packed + w + 1), (int)(p & 15)) >> (32 - 2 * m);
            const uint32_t rc = rc_mmer(fw, m);
            cn[s] = min(fw, rc);
            rv[s] = cn[s] != fw;
            h[s] = xxh64_8(cn[s]);
        }
    }
    // minimal hash of the window (two 32-bit reductions)
    const uint64_t hl = umin64(h[0], h[1]);
    const uint32_t hi_min = __reduce_min_sync(0xffffffffu, (uint32_t)(hl >> 32));
    const uint32_t lo_min = __reduce_min_sync(0xffffffffu, (uint32_t)(hl >> 32) == hi_min ? (uint32_t)hl : 0xFFFFFFFFu);
    const uint64_t hmin = ((uint64_t)hi_min << 32) | lo_min;
    const unsigned w0 = __ballot_sync(0xffffffffu, h[0] == hmin), w1 = __ballot_sync(0xffffffffu, h[1] == hmin && lane + 32 <= d);
    const int j0 = w0 ? __ffs(w0) - 1 : 32 + __ffs(w1) - 1;
    const int src = j0 & 31;
    const uint32_t best = __shfl_sync(0xffffffffu, j0 < 32 ? cn[0] : cn[1], src);
    const bool rev = __shfl_sync(0xffffffffu, (int)(j0 < 32 ? rv[0] : rv[1]), src) != 0;
    const unsigned t0 = __ballot_sync(0xffffffffu, lane > j0 && lane <= d && cn[0] == best && rv[0] == rev);
    const unsigned t1 = __ballot_sync(0xffffffffu, lane + 32 > j0 && lane + 32 <= d && cn[1] == best && rv[1] == rev);
    int pos = j0 == 0 ? (rev ? 0 : d) : d - j0;
    if (!rev) {
        if (t1) pos = d - (32 + 31 - __clz(t1));
        else if (t0) pos = d - (31 - __clz(t0));
    } else {
        const int jm = t0 ? __ffs(t0) - 1 : (t1 ? 32 + __ffs(t1) - 1 : 0x7fffffff);
        if (jm < pos) pos = jm;
    }
    return (uint64_t)pos;
}

__device__ __forceinline__ uint64_t first_rec_ending_after(const uint64_t *__restrict__ rec_end, uint64_t n_rec, uint64_t pos)
{
    uint64_t lo = 0, hi = n_rec;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (rec_end[mid] <= pos) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One thread per 32-position row finds the segments that start at its "new minimum" bits and counts the
// selected k-mers among its positions; the segments are queued in shared memory and replayed by whole warps
// (the rescans inside a segment are sequential, each one is a warp-wide reduction).
// totals[2*input] += super-k-mer boundaries, totals[2*input+1] += selected k-mers.
constexpr int DS_QCAP = 96;                 // queued segments per warp; a lane adds at most 32 per row

struct DsItem {
    uint64_t rb, pm, c_end;                 // record start, position_min (relative), first iteration not replayed
};

__global__ void __launch_bounds__(256)
dense_segments_kernel(const uint32_t *__restrict__ packed, uint64_t n_bases, const uint32_t *__restrict__ nm_bits,
                      const uint32_t *__restrict__ sel_bits, const uint64_t *__restrict__ rec_begin,
                      const uint64_t *__restrict__ rec_end, const uint32_t *__restrict__ rec_input, uint64_t n_rec,
                      int k, int m, unsigned long long *__restrict__ totals)
{
    __shared__ unsigned long long s_acc[2];
    __shared__ uint32_t s_input;
    __shared__ DsItem s_q[8][DS_QCAP];
    __shared__ uint32_t s_qin[8][DS_QCAP];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const uint64_t n_pos = n_bases >= (uint64_t)m ? n_bases - m + 1 : 0;
    const uint64_t n_rows = (n_pos + 31) >> 5;
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int d = k - m;
    if (threadIdx.x == 0) {
        s_acc[0] = 0; s_acc[1] = 0; s_input = 0xFFFFFFFFu;
        // the block's common input = input of the first record that ends behind the block's first position
        if (n_rec) {
            const uint64_t r0 = first_rec_ending_after(rec_end, n_rec, (uint64_t)blockIdx.x * blockDim.x * 32);
            if (r0 < n_rec) s_input = rec_input[r0];
        }
    }
    __syncthreads();
    const uint32_t block_input = s_input;
    auto add = [&](uint32_t input, unsigned long long bnd, unsigned long long sel) {
        if (input == 0xFFFFFFFFu || (bnd | sel) == 0) return;
        if (input != block_input) {
            if (bnd) atomicAdd(totals + 2 * input, bnd);
            if (sel) atomicAdd(totals + 2 * input + 1, sel);
        } else {
            if (bnd) atomicAdd(&s_acc[0], bnd);
            if (sel) atomicAdd(&s_acc[1], sel);
        }
    };
    uint32_t qn = 0;                                             // queued segments of this warp (warp-uniform)
    // replay every queued segment with the whole warp
    auto drain = [&]() {
        __syncwarp();
        uint32_t cur_in = 0xFFFFFFFFu;
        unsigned long long cnt = 0;
        for (uint32_t i = 0; i < qn; i++) {
            const DsItem it = s_q[wi][i];
            const uint32_t in_i = s_qin[wi][i];
            if (in_i != cur_in) { if (lane == 0) add(cur_in, cnt, 0); cur_in = in_i; cnt = 0; }
            uint64_t pm = it.pm;
            for (uint64_t c = pm + 1; c < it.c_end; c = pm + 1) {   // iteration c rescans when c-1 >= position_min (:391)
                pm = c + warp_rescan(packed, it.rb + c, d, m, lane); // position_min += i + 1 (:397)
                cnt++;                                              // dump: the super-k-mer ends (:401)
            }
        }
        if (lane == 0) add(cur_in, cnt, 0);
        __syncwarp();
        qn = 0;
    };

    // per-lane state while walking the row
    uint32_t my_input = 0xFFFFFFFFu;
    unsigned long long my_bound = 0, my_sel = 0;
    uint64_t r = 0, r_stop = 0;
    uint32_t nmw = 0, slw = 0;
    const uint64_t g0 = row << 5, g1 = g0 + 31;
    if (row < n_rows && n_rec) {
        nmw = nm_bits[row]; slw = sel_bits[row];
        r = first_rec_ending_after(rec_end, n_rec, g0);
        r_stop = r;
        while (r_stop < n_rec && rec_begin[r_stop] <= g1) r_stop++;
    }
    // the warp walks in rounds: every lane contributes the segments of ONE record overlap per round
    for (;;) {
        const bool have = r < r_stop;
        if (!__any_sync(0xffffffffu, have)) break;
        uint32_t bits = 0;
        uint64_t rb = 0, n = 0, K = 0;
        uint32_t input = 0;
        if (have) {
            rb = rec_begin[r]; n = rec_end[r] - rb; input = rec_input[r];
            if (n >= (uint64_t)k) {
                K = n - k + 1;
                auto mask_range = [&](uint64_t lo_rel, uint64_t hi_rel) -> uint32_t {
                    const uint64_t lo = rb + lo_rel, hi = rb + hi_rel;
                    if (hi_rel < lo_rel || hi < g0 || lo > g1) return 0u;
                    const uint32_t a = lo > g0 ? (uint32_t)(lo - g0) : 0u, b = hi < g1 ? (uint32_t)(hi - g0) : 31u;
                    return (0xFFFFFFFFu << a) & (0xFFFFFFFFu >> (31 - b));
                };
                if (input != my_input) { add(my_input, my_bound, my_sel); my_input = input; my_bound = my_sel = 0; }
                // selected k-mers: the k-mer that starts at c ends its window at m-mer c + d
                my_sel += __popc(slw & mask_range((uint64_t)d, n - m));
                // new-minimum iterations c = 1 .. K-1 enter m-mer c + d
                if (K > 1) bits = nmw & mask_range((uint64_t)d + 1, n - m);
                my_bound += __popc(bits);                        // the minimizer changes there (:374-388, :401)
            }
            r++;
        }
        // queue the segments of this round (at most 32 per lane); drain first if they might not fit
        const uint32_t mine = __popc(bits);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t done_before = 0;                                // items of this round already queued (over all lanes)
        while (done_before < total) {
            if (qn == DS_QCAP) drain();
            const uint32_t room = DS_QCAP - qn;
            // lanes write the items whose round-index falls into [done_before, done_before + room)
            uint32_t idx = incl - mine;                          // round-index of my first item
            uint32_t b2 = bits;
            while (b2) {
                const uint32_t bit = __ffs(b2) - 1;
                b2 &= b2 - 1;
                if (idx >= done_before && idx < done_before + room) {
                    const uint64_t p_rel = g0 + bit - rb;        // the entering m-mer; position_min = p_rel
                    const uint64_t nx = next_flag(nm_bits, rb, p_rel + 1, n - m, ~0ULL);
                    DsItem it;
                    it.rb = rb; it.pm = p_rel; it.c_end = nx == ~0ULL ? K : nx - d;
                    const uint32_t slot = qn + (idx - done_before);
                    s_q[wi][slot] = it;
                    s_qin[wi][slot] = input;
                }
                idx++;
            }
            const uint32_t put = min(room, total - done_before);
            qn += put;
            done_before += put;
        }
    }
    drain();
    add(my_input, my_bound, my_sel);
    __syncthreads();
    if (threadIdx.x == 0 && block_input != 0xFFFFFFFFu) {
        if (s_acc[0]) atomicAdd(totals + 2 * block_input, s_acc[0]);
        if (s_acc[1]) atomicAdd(totals + 2 * block_input + 1, s_acc[1]);
    }
}

// One thread per record: its first k-mer's rescan (:359-365) and what follows until the first new minimum, plus
// the record's last super-k-mer (:441-454).
__global__ void dense_records_kernel(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ nm_bits,
                                     const uint64_t *__restrict__ rec_begin, const uint64_t *__restrict__ rec_end,
                                     const uint32_t *__restrict__ rec_input, uint64_t n_rec, int k, int m,
                                     unsigned long long *__restrict__ totals)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t rb = rec_begin[r], n = rec_end[r] - rb;
    if (n < (uint64_t)k) return;
    const int d = k - m;
    const uint64_t K = n - k + 1;
    unsigned long long bound = 1;                                           // the tail super-k-mer
    if (K > 1) {
        const uint64_t pm = dense_rescan(packed, rb, d, m);                 // position_min of k-mer 0
        const uint64_t nx = next_flag(nm_bits, rb, (uint64_t)d + 1, n - m, ~0ULL);
        const uint64_t c_end = nx == ~0ULL ? K : nx - d;
        bound += replay_rescans(packed, rb, pm, c_end, d, m);
    }
    atomicAdd(totals + 2 * rec_input[r], bound);
}

#define DN_CK(call)                                  \
    do {                                             \
        cudaError_t e_ = (call);                     \
        if (e_ != cudaSuccess) return e_;            \
    } while (0)

cudaError_t dense_stats_run(DenseBuffers *b, const DenseIn &in, uint64_t *h_total_superkmers,
                            uint64_t *h_selected_kmers, float *kernel_ms, uint32_t *launched, cudaStream_t st)
{
    const uint64_t n_pos = in.n_bases >= (uint64_t)in.m ? in.n_bases - in.m + 1 : 0;
    const uint64_t n_rows = (n_pos + 31) >> 5;
    const size_t nin = in.n_inputs ? in.n_inputs : 1;
    if (!b->ev0) { DN_CK(cudaEventCreate(&b->ev0)); DN_CK(cudaEventCreate(&b->ev1)); }
    DN_CK(b->nm.ensure((n_rows + 2) * 4)); DN_CK(b->sel.ensure((n_rows + 2) * 4)); DN_CK(b->totals.ensure(nin * 16));
    DN_CK(cudaMemsetAsync(b->totals.p, 0, nin * 16, st));
    DN_CK(cudaEventRecord(b->ev0, st));
    uint32_t nl = 0;
    if (n_rows && in.n_rec) {
        const uint64_t warps = (n_rows + DN_ROWS - 1) / DN_ROWS;
        dense_rows_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(in.d_packed, in.n_bases, in.k, in.m, in.thr,
                                                                       b->nm.as<uint32_t>(), b->sel.as<uint32_t>());
        dense_segments_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(in.d_packed, in.n_bases, b->nm.as<uint32_t>(),
            b->sel.as<uint32_t>(), in.d_rec_begin, in.d_rec_end, in.d_rec_input, in.n_rec, in.k, in.m,
            b->totals.as<unsigned long long>());
        dense_records_kernel<<<(unsigned)((in.n_rec + 127) / 128), 128, 0, st>>>(in.d_packed, b->nm.as<uint32_t>(), in.d_rec_begin,
            in.d_rec_end, in.d_rec_input, in.n_rec, in.k, in.m, b->totals.as<unsigned long long>());
        nl = 3;
    }
    DN_CK(cudaEventRecord(b->ev1, st));
    std::vector<unsigned long long> h(nin * 2);
    DN_CK(cudaMemcpyAsync(h.data(), b->totals.p, nin * 16, cudaMemcpyDeviceToHost, st));
    DN_CK(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < in.n_inputs; i++) {
        if (h_total_superkmers) h_total_superkmers[i] = h[2 * i];
        if (h_selected_kmers) h_selected_kmers[i] = h[2 * i + 1];
    }
    if (kernel_ms) DN_CK(cudaEventElapsedTime(kernel_ms, b->ev0, b->ev1));
    if (launched) *launched = nl;
    return cudaGetLastError();
}

}  // namespace spsp
