// Shared device helpers: tile geometry, mbarrier / bulk-copy (TMA) PTX wrappers, min-loc.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bopy {

// ---- tile geometry of the blocked solve ------------------------------------------------------
constexpr int BM = 128;        // training rows per block row (= size of an inverted diagonal block)
constexpr int BN = 128;        // candidates per tile (one CTA owns one tile at a time)
constexpr int NT = 256;        // compute threads per CTA (8 warps); the sweep kernel adds a producer warpgroup
constexpr int STAGES = 4;      // bulk-copy ring depth
constexpr int TILE_BYTES = 8192;  // one operand tile: BM x KC x sizeof(T) = BN x KC x sizeof(T)
constexpr int MAX_D = 32;

template <typename T> struct Geo {
    static constexpr int VEC = 16 / sizeof(T);           // elements per 128-bit shared load
    static constexpr int KC = TILE_BYTES / (BM * sizeof(T));  // contraction depth per tile: 8 (f64) / 16 (f32)
    static constexpr int CH = BM / KC;                   // tiles per 128 columns: 16 (f64) / 8 (f32)
};

// ---- argmin record -----------------------------------------------------------------------------
struct MinLoc {
    double val;
    long long idx;   // < 0: empty
};

// np.argmin ordering: a NaN beats every number, earlier index wins among equals / among NaNs.
__host__ __device__ __forceinline__ bool minloc_better(const MinLoc& a, const MinLoc& b) {
    if (a.idx < 0) return false;
    if (b.idx < 0) return true;
    const bool an = a.val != a.val, bn = b.val != b.val;
    if (an || bn) {
        if (an && bn) return a.idx < b.idx;
        return an;
    }
    if (a.val < b.val) return true;
    if (a.val > b.val) return false;
    return a.idx < b.idx;
}

#ifdef __CUDACC__
__device__ __forceinline__ MinLoc minloc_shfl_xor(const MinLoc& v, int mask) {
    MinLoc o;
    o.val = __shfl_xor_sync(0xffffffffu, v.val, mask);
    o.idx = __shfl_xor_sync(0xffffffffu, v.idx, mask);
    return o;
}

__device__ __forceinline__ MinLoc minloc_warp_reduce(MinLoc v) {
#pragma unroll
    for (int mask = 16; mask > 0; mask >>= 1) {
        MinLoc o = minloc_shfl_xor(v, mask);
        if (minloc_better(o, v)) v = o;
    }
    return v;
}

// ---- mbarrier + bulk async copy (TMA, 1-D) -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// orders generic-proxy accesses (st.global / st.shared done so far) before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// global -> shared bulk copy; completion is signalled as `bytes` transaction bytes on `bar`.
// dst/src 16-byte aligned, bytes a multiple of 16.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// bring `bytes` of global memory into L2 ahead of the bulk copy that will need them (no shared-memory cost, no completion signal)
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
#endif  // __CUDACC__

}  // namespace bopy
