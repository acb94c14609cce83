// tcgen05 / TMEM PTX wrappers for the fp32-mode engine (sm_100a): tensor-memory allocation, shared-memory matrix
// descriptors, the kind::tf32 MMA, commit -> mbarrier, and tensor-memory loads in the m16n8 fragment shape.
//
// Operand tiles are K-major in the canonical "interleaved" (no swizzle) layout of the tensor core:
//   element (mn, k) of a [128 mn][8 k] fp32 tile lives at float offset (mn/8)*64 + (k/4)*32 + (mn%8)*4 + (k%4)
// i.e. core matrices of 8 mn-rows x 16 bytes (4 k) = 128 contiguous bytes; the two core matrices of a row group that
// make up K = 8 are adjacent (leading byte offset 128), row groups follow at 256 bytes (stride byte offset).
// One tcgen05.mma kind::tf32 (K = 8) consumes exactly one such 4 KB tile per operand.
// (MN-major operands are not available to kind::tf32: measured on B200, the instruction then leaves zeros --
// tools/tcgen05_probe.cu, profiles/r02/tcgen05_probe.log.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bopy {
namespace tc {

constexpr int TF32_TILE_FLOATS = 128 * 8;       // one operand tile of one MMA: 128 (m or n) x 8 (k)
constexpr int TF32_TILE_BYTES = TF32_TILE_FLOATS * 4;

__host__ __device__ __forceinline__ int tile_index(int k, int mn) { return (mn >> 3) * 64 + (k >> 2) * 32 + (mn & 7) * 4 + (k & 3); }
constexpr uint32_t TILE_LBO = 128, TILE_SBO = 256;

// shared-memory matrix descriptor (tcgen05 "version 1"), no swizzle:
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4, [46,48) = 1
// K-major interleaved: LBO = distance between the core matrices adjacent in k, SBO = distance between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor of kind::tf32 with fp32 accumulation: c_format F32 (bit 4), a/b format TF32 (= 2) at bits 7 / 10,
// a/b major MN (bits 15 / 16), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#ifdef __CUDACC__
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one lane of a fully converged warp (elect.sync): the issuing lane of tcgen05.mma / commit / bulk copies when the
// whole warp runs the surrounding loop, so that counters and descriptors stay warp-uniform (uniform registers in SASS)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, px;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
                 : "memory");
}

// 16 lanes x 256 bit, repeated 8 times along the columns: 16 rows x 64 fp32 columns starting at (lane, column) of taddr.
// Thread t receives r[4 g + 2 h + e] = D[lane + t/4 + 8 h][column + 8 g + 2 (t%4) + e]  (the m16n8 accumulator fragment)
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 32 bit, repeated 32 times along the columns: thread t receives r[j] = D[lane + t][column + j]
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// cvt.rna.tf32.f32: round an fp32 to the nearest TF32 (kept in an fp32 container, low 13 bits zero)
__device__ __forceinline__ float to_tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// x (fp64) ~ hi + lo with hi, lo both TF32, both rounded to nearest: |x - hi - lo| <= 2^-22 |x|
__device__ __forceinline__ void tf32_pair(double x, float& hi, float& lo) {
    hi = to_tf32_rn(static_cast<float>(x));
    lo = to_tf32_rn(static_cast<float>(x - static_cast<double>(hi)));
}
// the same from the fp32 rounding of x: xf - hi is exact in fp32, so one conversion instead of three
// (|x - hi - lo| <= (2^-22 + 2^-24) |x|); used where the pair is formed per element in the hot loop
__device__ __forceinline__ void tf32_pair_fast(double x, float& hi, float& lo) {
    const float xf = static_cast<float>(x);
    hi = to_tf32_rn(xf);
    lo = to_tf32_rn(xf - hi);
}
#endif

}  // namespace tc
}  // namespace bopy
