// Stand-alone probe of the tcgen05 building blocks the fp32-mode engine relies on (run on a B200 through gpurun):
//   T1  D = A.B^T with MN-major interleaved operand tiles (layout / descriptor check against the host)
//   T2  how the tensor core rounds when it adds into its fp32 accumulator (chains of MMAs with crafted increments)
//   T3  issue rate of kind::tf32 M=128 N=128/256 K=8 from shared-memory operands, one CTA per SM
//   T4  tensor-memory load rate (16x256b.x8) with 8 warps
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bopy_b200/csrc -o tools/bin/tcgen05_probe tools/tcgen05_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

using namespace bopy;

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e__ = (x);                                                                     \
        if (e__ != cudaSuccess) {                                                                  \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__);       \
            exit(1);                                                                               \
        }                                                                                          \
    } while (0)

// ------------------------------------------------------------------------------------------------------------------
// T1 / T2: nmma MMAs (each K = 8) chained into one accumulator; tile t of A / B is used by MMA t.
// out_frag[128][N] written from the 16x256b fragments, out_row[128][N] from 32x32b loads (cross-check of both shapes).
__global__ void __launch_bounds__(256, 1) chain_kernel(const float* __restrict__ At, const float* __restrict__ Bt, int nmma, int N,
                                                        float* out_frag, float* out_row) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* As = reinterpret_cast<float*>(smem);                       // [nmma][1024]
    float* Bs = As + (size_t)nmma * tc::TF32_TILE_FLOATS;             // [nmma][N*8]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < nmma * tc::TF32_TILE_FLOATS; i += blockDim.x) As[i] = At[i];
    for (int i = tid; i < nmma * N * 8; i += blockDim.x) Bs[i] = Bt[i];
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, N, false, false);
        for (int t = 0; t < nmma; ++t) {
            const uint64_t ad = tc::smem_desc(smem_u32(As + (size_t)t * tc::TF32_TILE_FLOATS), tc::TILE_LBO, tc::TILE_SBO);
            const uint64_t bd = tc::smem_desc(smem_u32(Bs + (size_t)t * N * 8), tc::TILE_LBO, tc::TILE_SBO);
            tc::mma_tf32(tmem, ad, bd, idesc, t > 0);
        }
        tc::mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc::fence_after_thread_sync();
    const int rg = warp & 3, cg = warp >> 2;     // warp may only touch lanes 32 rg .. 32 rg + 31
    for (int c0 = 64 * cg; c0 < N; c0 += 128) {
        for (int mt = 0; mt < 2; ++mt) {
            uint32_t r[32];
            tc::tmem_ld_16x256b_x8(tmem + ((uint32_t)(32 * rg + 16 * mt) << 16) + c0, r);
            tc::tmem_wait_ld();
            for (int g = 0; g < 8; ++g)
                for (int h = 0; h < 2; ++h)
                    for (int e = 0; e < 2; ++e)
                        out_frag[(size_t)(32 * rg + 16 * mt + (lane >> 2) + 8 * h) * N + c0 + 8 * g + 2 * (lane & 3) + e] =
                            __uint_as_float(r[4 * g + 2 * h + e]);
        }
        for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tc::tmem_ld_32x32b_x32(tmem + ((uint32_t)(32 * rg) << 16) + c0 + 32 * half, r);
            tc::tmem_wait_ld();
            for (int j = 0; j < 32; ++j) out_row[(size_t)(32 * rg + lane) * N + c0 + 32 * half + j] = __uint_as_float(r[j]);
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// T3: issue rate.  `iters` rounds of 3 MMAs (the 3xTF32 pattern: lo.hi, hi.lo, hi.hi on different tiles) into one accumulator.
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int N, int nacc, long long* clocks_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* As = reinterpret_cast<float*>(smem);   // 2 A tiles (hi, lo)
    float* Bs = As + 2 * tc::TF32_TILE_FLOATS;    // 2 B tiles of N x 8
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 2 * tc::TF32_TILE_FLOATS + 2 * N * 8; i += blockDim.x) As[i] = 0.0f;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, N, false, false);
        const uint64_t ah = tc::smem_desc(smem_u32(As), tc::TILE_LBO, tc::TILE_SBO), al = tc::smem_desc(smem_u32(As + tc::TF32_TILE_FLOATS), tc::TILE_LBO, tc::TILE_SBO);
        const uint64_t bh = tc::smem_desc(smem_u32(Bs), tc::TILE_LBO, tc::TILE_SBO), bl = tc::smem_desc(smem_u32(Bs + N * 8), tc::TILE_LBO, tc::TILE_SBO);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem + (uint32_t)((it % nacc) * N);
            tc::mma_tf32(d, al, bh, idesc, 1);
            tc::mma_tf32(d, ah, bl, idesc, 1);
            tc::mma_tf32(d, ah, bh, idesc, 1);
        }
        tc::mma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        clocks_out[blockIdx.x] = t1 - t0;
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// T4: tensor-memory load rate: 8 warps, each `iters` x (two 16x256b.x8 loads = its 32 lanes x 64 columns)
__global__ void __launch_bounds__(256, 1) ldtm_kernel(int iters, long long* clocks_out, float* sink) {
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = tmem_base_s;
    const int rg = warp & 3, cg = warp >> 2;
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t r0[32], r1[32];
        const uint32_t col = (uint32_t)(((it & 3) * 128) + 64 * cg);
        tc::tmem_ld_16x256b_x8(tmem + ((uint32_t)(32 * rg) << 16) + col, r0);
        tc::tmem_ld_16x256b_x8(tmem + ((uint32_t)(32 * rg + 16) << 16) + col, r1);
        tc::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) clocks_out[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[tid] = acc;
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// T5: fp32 -> fp64 conversion + add rate (the fold of a tensor-memory partial sum into the fp64 residual)
__global__ void __launch_bounds__(256, 1) fold_kernel(int iters, long long* clocks_out, double* sink, const float* src) {
    double r[64];
    float f[64];
    for (int j = 0; j < 64; ++j) {
        r[j] = 0.0;
        f[j] = src[(threadIdx.x + j) & 255];
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            r[j] += static_cast<double>(f[j]);
            f[j] = __uint_as_float(__float_as_uint(f[j]) ^ (it & 1));
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) clocks_out[blockIdx.x] = t1 - t0;
    double s = 0;
    for (int j = 0; j < 64; ++j) s += r[j];
    if (s == 1.2345) sink[threadIdx.x] = s;
}

// T0: diagnostics.  (a) tensor-memory store/load round trip, (b) one MMA under several operand layouts / descriptor variants.
__device__ __forceinline__ void tmem_st_32x32b_x1(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__global__ void __launch_bounds__(128, 1) diag_kernel(const float* __restrict__ At, const float* __restrict__ Bt, int variant, uint32_t lboA,
                                                       uint32_t sboA, uint32_t lboB, uint32_t sboB, int a_mn, int b_mn, float* out, float* rt) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* As = reinterpret_cast<float*>(smem);
    float* Bs = As + 1024;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 1024; i += blockDim.x) {
        As[i] = At[i];
        Bs[i] = Bt[i];
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) rt[130] = __uint_as_float(tmem);
    // (a) round trip: lane l of the warp's quarter, column 300 + warp  <-  1000 * warp + lane
    tmem_st_32x32b_x1(tmem + ((uint32_t)(32 * warp) << 16) + 300 + warp, __float_as_uint(1000.0f * warp + lane));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    {
        uint32_t r[32];
        tc::tmem_ld_32x32b_x32(tmem + ((uint32_t)(32 * warp) << 16) + 288, r);
        tc::tmem_wait_ld();
        rt[tid] = __uint_as_float(r[12 + warp]);
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, 128, a_mn != 0, b_mn != 0);
        const uint64_t ad = tc::smem_desc(smem_u32(As), lboA, sboA);
        const uint64_t bd = tc::smem_desc(smem_u32(Bs), lboB, sboB);
        if (variant == 0) tc::mma_tf32(tmem, ad, bd, idesc, 0);
        tc::mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc::fence_after_thread_sync();
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld_32x32b_x32(tmem + ((uint32_t)(32 * warp) << 16) + c0, r);
        tc::tmem_wait_ld();
        for (int j = 0; j < 32; ++j) out[(size_t)(32 * warp + lane) * 128 + c0 + j] = __uint_as_float(r[j]);
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

static void run_diag() {
    // logical A[m][k], B[n][k] with distinctive values; D[m][n] = sum_k A[m][k] B[n][k]
    std::vector<float> Al(128 * 8), Bl(128 * 8);
    for (int m = 0; m < 128; ++m)
        for (int k = 0; k < 8; ++k) {
            Al[m * 8 + k] = (float)((m * 3 + k * 5) % 11 - 5);
            Bl[m * 8 + k] = (float)((m * 7 + k * 2) % 9 - 4);
        }
    std::vector<double> D(128 * 128);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
            double s = 0;
            for (int k = 0; k < 8; ++k) s += (double)Al[m * 8 + k] * Bl[n * 8 + k];
            D[m * 128 + n] = s;
        }
    float *dA, *dB, *dO, *dR;
    CK(cudaMalloc(&dA, 4096));
    CK(cudaMalloc(&dB, 4096));
    CK(cudaMalloc(&dO, 128 * 128 * 4));
    CK(cudaMalloc(&dR, 256 * 4));
    CK(cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    struct Var { const char* name; int mn; uint32_t lbo, sbo; int layout; };
    // layout 0: MN-major interleaved [mn/4][k][mn%4]   (core matrix = 8 k x 4 mn, core matrices 128 B apart along mn)
    // layout 1: K-major interleaved  [mn/8][k/4][mn%8][k%4] (core matrix = 8 mn x 4 k; the two k core matrices adjacent)
    // layout 2: MN-major, k-group-major [k][mn] plain row-major (k rows of 128 mn): core matrix rows 512 B apart -> not canonical, control
    const Var vars[] = {{"MN-major interleaved LBO=4096 SBO=128", 1, 4096, 128, 0}, {"MN-major interleaved LBO=128 SBO=4096", 1, 128, 4096, 0},
                        {"MN-major interleaved LBO=128 SBO=128", 1, 128, 128, 0},  {"K-major interleaved LBO=128 SBO=256", 0, 128, 256, 1},
                        {"K-major interleaved LBO=256 SBO=128", 0, 256, 128, 1}};
    for (const Var& v : vars) {
        std::vector<float> A(1024, 0.f), B(1024, 0.f);
        for (int mn = 0; mn < 128; ++mn)
            for (int k = 0; k < 8; ++k) {
                int idx = v.layout == 0 ? (mn >> 2) * 32 + k * 4 + (mn & 3) : tc::tile_index(k, mn);
                A[idx] = Al[mn * 8 + k];
                B[idx] = Bl[mn * 8 + k];
            }
        CK(cudaMemcpy(dA, A.data(), 4096, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), 4096, cudaMemcpyHostToDevice));
        CK(cudaMemset(dO, 0xff, 128 * 128 * 4));
        CK(cudaMemset(dR, 0xff, 256 * 4));
        diag_kernel<<<1, 128, 16384>>>(dA, dB, 0, v.lbo, v.sbo, v.lbo, v.sbo, v.mn, v.mn, dO, dR);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("T0 %s: CUDA error %s\n", v.name, cudaGetErrorString(e));
            exit(1);
        }
        std::vector<float> O(128 * 128), R(256);
        CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(R.data(), dR, R.size() * 4, cudaMemcpyDeviceToHost));
        double mx = 0;
        int nz = 0;
        for (int i = 0; i < 128 * 128; ++i) {
            mx = fmax(mx, fabs(O[i] - D[i]));
            nz += O[i] != 0.0f;
        }
        uint32_t tb;
        memcpy(&tb, &R[130], 4);
        printf("T0 %-42s: max err %.3g, nonzero outputs %d / 16384, tmem base 0x%08x, st/ld round trip lane0..2 of warp 0/1/3: %g %g %g | %g %g | %g\n", v.name, mx, nz, tb,
               R[0], R[1], R[2], R[32], R[33], R[96]);
        printf("      D[0][0..5] got %g %g %g %g %g %g want %g %g %g %g %g %g ; D[1][0] %g want %g ; D[8][0] %g want %g ; D[64][64] %g want %g\n", O[0], O[1], O[2], O[3],
               O[4], O[5], D[0], D[1], D[2], D[3], D[4], D[5], O[128], D[128], O[8 * 128], D[8 * 128], O[64 * 128 + 64], D[64 * 128 + 64]);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dR);
}

static float tf32_trunc_host(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev);
    run_diag();

    // ---------------- T1: layout / descriptor check -----------------------------------------------------------------
    for (int N : {128, 256}) {
        const int nmma = 3;
        std::vector<float> A(nmma * 1024), B((size_t)nmma * N * 8), Aref((size_t)nmma * 128 * 8), Bref((size_t)nmma * N * 8);
        srand(7);
        for (int t = 0; t < nmma; ++t)
            for (int k = 0; k < 8; ++k) {
                for (int m = 0; m < 128; ++m) {
                    const float v = (float)((rand() % 17) - 8) / 8.0f;
                    A[(size_t)t * 1024 + tc::tile_index(k, m)] = v;
                    Aref[((size_t)t * 8 + k) * 128 + m] = v;
                }
                for (int n = 0; n < N; ++n) {
                    const float v = (float)((rand() % 13) - 6) / 4.0f;
                    B[(size_t)t * N * 8 + tc::tile_index(k, n)] = v;
                    Bref[((size_t)t * 8 + k) * N + n] = v;
                }
            }
        float *dA, *dB, *dF, *dR;
        CK(cudaMalloc(&dA, A.size() * 4));
        CK(cudaMalloc(&dB, B.size() * 4));
        CK(cudaMalloc(&dF, (size_t)128 * N * 4));
        CK(cudaMalloc(&dR, (size_t)128 * N * 4));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dF, 0xff, (size_t)128 * N * 4));
        CK(cudaMemset(dR, 0xff, (size_t)128 * N * 4));
        const size_t smem = (A.size() + B.size()) * 4;
        CK(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chain_kernel<<<1, 256, smem>>>(dA, dB, nmma, N, dF, dR);
        CK(cudaDeviceSynchronize());
        std::vector<float> F((size_t)128 * N), R((size_t)128 * N);
        CK(cudaMemcpy(F.data(), dF, F.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(R.data(), dR, R.size() * 4, cudaMemcpyDeviceToHost));
        double ef = 0, er = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int t = 0; t < nmma; ++t)
                    for (int k = 0; k < 8; ++k) s += (double)Aref[((size_t)t * 8 + k) * 128 + m] * Bref[((size_t)t * 8 + k) * N + n];
                ef = fmax(ef, fabs(F[(size_t)m * N + n] - s));
                er = fmax(er, fabs(R[(size_t)m * N + n] - s));
            }
        printf("T1 N=%d: max |D - A.B^T| via 16x256b fragments %.3g, via 32x32b rows %.3g (exact small integers: expect 0)\n", N, ef, er);
        cudaFree(dA); cudaFree(dB); cudaFree(dF); cudaFree(dR);
    }

    // ---------------- T2: accumulator rounding ----------------------------------------------------------------------
    {
        const int N = 128, nmma = 65;
        const double u = ldexp(1.0, -23);
        // column n of MMA t >= 1 adds inc[n % NI] * u through ONE product (k = 0); MMA 0 sets the accumulator to 1.0 (rows 0..63)
        // or to -1.0 (rows 64..127).  Columns >= 64: the increment is spread over 8 products of inc/8 each.
        const double inc[8] = {0.75, 0.5, 0.25, -0.25, -0.5, -0.75, 1.5, 0.375};
        std::vector<float> A((size_t)nmma * 1024, 0.f), B((size_t)nmma * N * 8, 0.f);
        for (int m = 0; m < 128; ++m) A[tc::tile_index(0, m)] = m < 64 ? 1.0f : -1.0f;
        for (int n = 0; n < N; ++n) B[tc::tile_index(0, n)] = 1.0f;
        for (int t = 1; t < nmma; ++t)
            for (int n = 0; n < N; ++n) {
                const double v = inc[n % 8] * u;
                if (n < 64) {
                    B[(size_t)t * N * 8 + tc::tile_index(0, n)] = (float)v;
                    for (int m = 0; m < 128; ++m) A[(size_t)t * 1024 + tc::tile_index(0, m)] = 1.0f;
                } else {
                    for (int k = 0; k < 8; ++k) {
                        B[(size_t)t * N * 8 + tc::tile_index(k, n)] = (float)(v / 8.0);
                        for (int m = 0; m < 128; ++m) A[(size_t)t * 1024 + tc::tile_index(k, m)] = 1.0f;
                    }
                }
            }
        float *dA, *dB, *dF, *dR;
        CK(cudaMalloc(&dA, A.size() * 4));
        CK(cudaMalloc(&dB, B.size() * 4));
        CK(cudaMalloc(&dF, (size_t)128 * N * 4));
        CK(cudaMalloc(&dR, (size_t)128 * N * 4));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        // shared memory: 65 tiles of A and B do not fit -> run in chunks?  No: use nmma = 17 per launch (17 * 8 KB = 136 KB)
        const int per = 17;
        const size_t smem = (size_t)per * (1024 + N * 8) * 4;
        CK(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chain_kernel<<<1, 256, smem>>>(dA, dB, per, N, dF, dR);
        CK(cudaDeviceSynchronize());
        std::vector<float> F((size_t)128 * N);
        CK(cudaMemcpy(F.data(), dF, F.size() * 4, cudaMemcpyDeviceToHost));
        printf("T2 accumulator rounding: start +1.0 (row 0) / -1.0 (row 64), then %d MMAs each adding inc*ulp(1) = inc*2^-23; result in ulps from start\n", per - 1);
        printf("   exact would be %d*inc; round-to-nearest-even per step: +16*round(inc) for |inc|>0.5, 0 for |inc|<0.5; truncation: see sign asymmetry\n", per - 1);
        for (int n = 0; n < 8; ++n)
            printf("   inc %+6.3f one product : row0 %+8.2f  row64 %+8.2f | 8 products of inc/8: row0 %+8.2f row64 %+8.2f\n", inc[n],
                   (F[n] - 1.0) / u, (F[(size_t)64 * N + n] + 1.0) / u, (F[64 + n] - 1.0) / u, (F[(size_t)64 * N + 64 + n] + 1.0) / u);
        // random chain: products of random TF32 values, compare with the exact sum (double) -> signed mean error in ulps of the result
        srand(11);
        std::vector<double> exact((size_t)128 * N, 0.0);
        for (int t = 0; t < per; ++t)
            for (int k = 0; k < 8; ++k) {
                for (int m = 0; m < 128; ++m) A[(size_t)t * 1024 + tc::tile_index(k, m)] = tf32_trunc_host(0.5f + (float)rand() / RAND_MAX);
                for (int n = 0; n < N; ++n) B[(size_t)t * N * 8 + tc::tile_index(k, n)] = tf32_trunc_host(0.5f + (float)rand() / RAND_MAX);
            }
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int t = 0; t < per; ++t)
                    for (int k = 0; k < 8; ++k) s += (double)A[(size_t)t * 1024 + tc::tile_index(k, m)] * B[(size_t)t * N * 8 + tc::tile_index(k, n)];
                exact[(size_t)m * N + n] = s;
            }
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        for (int len : {1, 2, 4, 8, 17}) {
            chain_kernel<<<1, 256, smem>>>(dA, dB, len, N, dF, dR);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(F.data(), dF, F.size() * 4, cudaMemcpyDeviceToHost));
            double mean = 0, mx = 0, rms = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double s = 0;
                    for (int t = 0; t < len; ++t)
                        for (int k = 0; k < 8; ++k) s += (double)A[(size_t)t * 1024 + tc::tile_index(k, m)] * B[(size_t)t * N * 8 + tc::tile_index(k, n)];
                    const double rel = (F[(size_t)m * N + n] - s) / s;
                    mean += rel;
                    rms += rel * rel;
                    mx = fmax(mx, fabs(rel));
                }
            mean /= 128.0 * N;
            rms = sqrt(rms / (128.0 * N));
            printf("   random positive chain of %2d MMAs: relative error mean %+.3e rms %.3e max %.3e (2^-24 = 5.96e-08)\n", len, mean, rms, mx);
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dF); cudaFree(dR);
    }

    // ---------------- T3: MMA issue rate ------------------------------------------------------------------------------
    {
        long long* dclk;
        CK(cudaMalloc(&dclk, prop.multiProcessorCount * sizeof(long long)));
        std::vector<long long> clk(prop.multiProcessorCount);
        for (int N : {128, 256}) {
            for (int nacc : {1, 2}) {
                for (int grid : {1, prop.multiProcessorCount}) {
                    const int iters = 4000;
                    const size_t smem = (size_t)(2 * 1024 + 2 * N * 8) * 4;
                    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    rate_kernel<<<grid, 128, smem>>>(iters, N, nacc, dclk);
                    CK(cudaDeviceSynchronize());
                    cudaEvent_t e0, e1;
                    cudaEventCreate(&e0);
                    cudaEventCreate(&e1);
                    cudaEventRecord(e0);
                    rate_kernel<<<grid, 128, smem>>>(iters, N, nacc, dclk);
                    cudaEventRecord(e1);
                    CK(cudaDeviceSynchronize());
                    float ms = 0;
                    cudaEventElapsedTime(&ms, e0, e1);
                    CK(cudaMemcpy(clk.data(), dclk, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                    long long mxc = 0;
                    for (int i = 0; i < grid; ++i) mxc = clk[i] > mxc ? clk[i] : mxc;
                    const double per_mma = (double)mxc / (3.0 * iters);
                    const double tflops = 2.0 * 128 * N * 8 * 3.0 * iters * grid / (ms * 1e-3) / 1e12;
                    printf("T3 N=%3d accumulators %d grid %3d: %.1f clk per MMA (%.0f MAC/clk/SM), kernel %.3f ms -> %.1f TFLOP/s TF32 dense\n", N, nacc,
                           grid, per_mma, 128.0 * N * 8 / per_mma, ms, tflops);
                }
            }
        }
        // T4
        float* sink;
        CK(cudaMalloc(&sink, 256 * 4));
        for (int grid : {1, prop.multiProcessorCount}) {
            const int iters = 2000;
            ldtm_kernel<<<grid, 256>>>(iters, dclk, sink);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(clk.data(), dclk, grid * sizeof(long long), cudaMemcpyDeviceToHost));
            long long mxc = 0;
            for (int i = 0; i < grid; ++i) mxc = clk[i] > mxc ? clk[i] : mxc;
            printf("T4 grid %3d: 128x128 fp32 accumulator to registers (8 warps, 16x256b.x8 x2 each): %.1f clk per 64 KB (%.1f B/clk/SM)\n", grid,
                   (double)mxc / iters, 65536.0 * iters / mxc);
        }
        // T5
        double* dsink;
        float* dsrc;
        CK(cudaMalloc(&dsink, 256 * 8));
        CK(cudaMalloc(&dsrc, 256 * 4));
        CK(cudaMemset(dsrc, 0, 256 * 4));
        fold_kernel<<<prop.multiProcessorCount, 256>>>(500, dclk, dsink, dsrc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(clk.data(), dclk, sizeof(long long), cudaMemcpyDeviceToHost));
        printf("T5 fold 128x128 (cvt f32->f64 + DADD, 256 threads x 64): %.1f clk per fold\n", (double)clk[0] / 500);
        cudaFree(dclk);
    }
    printf("probe done\n");
    return 0;
}
