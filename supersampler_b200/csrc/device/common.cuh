// Device-side helpers shared by the scan and compare kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

namespace spsp {

// Function attributes (opt-in shared memory sizes) belong to a device, and a process may drive several GPUs
// from several host threads: one bit per device, set after the attribute call succeeded there.
struct PerDeviceOnce {
    std::atomic<uint64_t> done{0};
    template <class F> cudaError_t run(F &&f)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const uint64_t bit = 1ull << (dev & 63);
        if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
        e = f();
        if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
        return e;
    }
};

// XXH64 of one 8-byte little-endian word with seed 1312: what the reference's
// unrevhash computes (SubSampler.cpp:64-67 -> include/xxhash64.h:158-163,
// :115-148, :188-191).  A bijection on u64.
__host__ __device__ __forceinline__ uint64_t rotl64(uint64_t x, int r)
{
    return (x << r) | (x >> (64 - r));
}
__host__ __device__ __forceinline__ uint64_t xxh64_8(uint64_t x)
{
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL,
                   P3 = 1609587929392839161ULL, P4 = 9650029242287828579ULL,
                   P5 = 2870177450012600261ULL;
    uint64_t r = 1312ULL + P5 + 8ULL;
    r ^= rotl64(x * P2, 31) * P1;
    r = rotl64(r, 27) * P1 + P4;
    r ^= r >> 33;
    r *= P2;
    r ^= r >> 29;
    r *= P3;
    r ^= r >> 32;
    return r;
}

// Reverse the sixteen 2-bit bases of a word and complement them (code ^ 2,
// reference utils.cpp:20-22 nuc2intrc / :449-462 rcbc).
__device__ __forceinline__ uint32_t rc_word(uint32_t x)
{
    uint32_t r = __brev(x);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return r ^ 0xAAAAAAAAu;
}

// Reverse complement of an m-mer held right-aligned in a u32 (m <= 15).
__host__ __device__ __forceinline__ uint32_t rc_mmer(uint32_t x, int m)
{
#ifdef __CUDA_ARCH__
    return rc_word(x) >> (32 - 2 * m);
#else
    uint32_t r = 0;
    for (int i = 0; i < m; i++) { r = (r << 2) | ((x & 3u) ^ 2u); x >>= 2; }
    return r;
#endif
}

// 128-bit streaming load that does not allocate in L1 (read-once input).
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// Same as ld_stream_u4 with a memory clobber: later shared-memory reads cannot be scheduled
// above it, which keeps a prefetch at the top of a loop body.
__device__ __forceinline__ uint4 ld_stream_u4_pinned(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// 16 bases starting `o` bases into the word pair (w0 = earlier bases),
// top-aligned: first base in bits 31..30.
__device__ __forceinline__ uint32_t window16(uint32_t w0, uint32_t w1, int o)
{
    return __funnelshift_l(w1, w0, 2 * o);
}

// ---- TMA bulk copy (global -> shared) signalled on an mbarrier ---------------
// One thread arms the barrier with the byte count and issues cp.async.bulk
// (SASS: UBLKCP); every thread then waits on the barrier's phase.  dst, src and
// bytes must be multiples of 16.
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    const uint32_t piece = 16384;                 // keep single copies modest
    for (uint32_t off = 0; off < bytes; off += piece) {
        const uint32_t n = bytes - off < piece ? bytes - off : piece;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(static_cast<char *>(dst_smem) + off)),
                       "l"(static_cast<const char *>(src_gmem) + off), "r"(n), "r"(b)
                     : "memory");
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t b = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(b), "r"(parity) : "memory");
    }
}

}  // namespace spsp
