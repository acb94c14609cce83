// The fused hot path: kernel tile K(X*, X) -> blocked triangular solve against the Cholesky factor
// -> diagonal variance reduction -> LCB / EI / POI epilogue -> per-CTA arg-min.
//
// Reference arithmetic ($SK = sklearn/gaussian_process, scikit-learn 1.9.0; bopy = /root/reference/bopy):
//   K_trans = kernel_(X*, X_train)                $SK/_gpr.py:446, $SK/kernels.py:1569-1570 / 1720-1729
//   mean    = y_std * (K_trans @ alpha_) + y_mean $SK/_gpr.py:447-450
//   V       = L^-1 K_trans^T                      $SK/_gpr.py:460-462
//   var     = (k(x,x) - sum_i V[i,c]^2) * y_std^2 diagonal of $SK/_gpr.py:466-469 (never the m x m matrix)
//   acq     = LCB / EI / POI                      bopy/acquisition.py:83-85, 99-106, 123-128
//   argmin  = np.argmin rules                     bopy/optimizer.py:99-107
//
// One CTA owns a tile of BN = 128 candidates and walks the n/128 block rows of L in order:
//   R_I = K*_I - sum_{J<I} L_IJ V_J     GEMM, operands streamed by bulk async copies (TMA)
//   V_I = inv(L_II) R_I                 second GEMM against the pre-inverted 128x128 diagonal block
// L is stored pre-tiled (8 KB [k][row] tiles, off-diagonal tiles negated, diagonal blocks inverted) so that
// every operand tile is ONE contiguous cp.async.bulk transfer; V_J tiles live in a per-CTA workspace in the
// same tile layout.  Warp roles: warps 0-7 compute (setmaxnreg 232), warp 8 (one lane, setmaxnreg 40) is the TMA producer running ahead
// through a 4-stage mbarrier ring, across block-row boundaries.
//
// Two inner-product engines ("policies"):
//   DmmaPolicy      fp64, mma.sync.m8n8k4.f64 (DMMA), warp tile 32 rows x 64 candidates, XOR-swizzled tiles
//   FmaPolicy<T>    register-tiled FMA, thread tile 8 x 8 (used for the fp32 solve)
#pragma once
#include <type_traits>

#include "common.cuh"

namespace bopy {

enum { K_RBF = 0, K_M12 = 1, K_M32 = 2, K_M52 = 3 };
enum { A_NONE = -1, A_LCB = 0, A_EI = 1, A_POI = 2 };

constexpr int NT_ALL = NT + 128;  // 8 compute warps + one producer warpgroup (register budgets are per 4 warps)
constexpr int REGS_COMPUTE = 232, REGS_PRODUCER = 40;  // setmaxnreg: 8*32*232 + 4*32*40 <= 65536

struct SweepParams {
    const void* Lt;     // packed factor tiles (TILE_BYTES each), tile t of block row I at Engine::row_base(I) + t
    const double* Xt;   // [n_blocks][d+1][BM]: X/l (dimension-major) then alpha, zero padded
    void* Vws;          // [slots][n_pad][BN] solve workspace (policy tile layout)
    const double* Xs;   // candidates (m, d) row-major
    long long m, ntiles;
    int n, n_blocks, d;
    int slot_per_tile;  // 1: workspace slot = tile index (V is exported), 0: slot = blockIdx.x
    double ls[MAX_D];
    double amp, kss, y_mean, y_std, y_var;
    int acq;
    double eta, kappa;
    double* mean_out;
    double* var_out;
    double* acq_out;
    long long index_base;
    MinLoc* partials;   // [gridDim.x] or nullptr when no arg-min is wanted
    MinLoc* tile_records;  // [ntiles] per-tile arg-min records (for segmented arg-min) or nullptr
    int xrow_separate;  // 1: the X/l block row has its own shared-memory buffer and is prefetched one row ahead
    int tc_dstages;     // tensor-core engine: depth of the inv(L_II) ring
    int tc_fold;        // tensor-core engine: operand tiles (8 columns of L each) per hi.hi accumulation chain (4 or 8)
    long long* tc_prof; // tensor-core engine: optional [gridDim.x][16] cycle counters per phase (BOPY_B200_TC_PROF=1), else nullptr
    int tc_stages;      // tensor-core engine: permanent slots of the (L_IJ, V_J) ring (4, fewer when d is large)
    int nan_skip;       // 1: candidates whose acquisition value is NaN never win the arg-min (np.nanargmin); 0: np.argmin (first NaN wins)
    // group mode (sweep_group_kernel.cuh): group_size CTAs share one candidate tile
    int group_size;     // CTAs per group (0 / 1: sweep_kernel, one tile per CTA)
    int group_slots;    // workspace slots (tiles in flight) per group
    int group_lead;     // rows of the next tile interleaved with the last rows of the current one in the job order
    unsigned* gctl;     // control block (GroupCtl), zeroed before the launch
    double* gpart;      // [groups][slots][n_blocks][2][BN] partial mean / sum v^2 of every block row
    int xs_stage;       // group mode: candidate rows of a tile are staged in shared memory by a bulk copy (room + alignment allow it)
    int zigzag;         // sweep_kernel: 1 = odd block rows read V_{I-2} .. V_0 then V_{I-1} (L2 reuse of the per-CTA workspace);
                        // 0 = ascending J, the order of group mode (same summation order = bit-identical results)
};

// exp(x) for x <= 0, branch-free so that the 64 evaluations a thread makes per block row interleave instead of
// serialising on libdevice's special-case branches (the kernel-tile step was latency bound on them), and short: the
// exponentials run on the same FP64 pipe as the DMMAs of the solve (10 FP64 instructions here; the degree-13 Taylor
// version of round 1 took 25: 13 % of the pipe at n = 256).
//   x = (64 k + j) ln2/64 + r,  |r| <= ln2/128,  exp(x) = 2^k * T[j] * (1 + q(r)),  T[j] = 2^(j/64) (correctly rounded),
//   n = 64 k + j = rint(x 64/ln2) by the 1.5*2^52 trick, r = x - n ln2/64 with ln2/64 split hi (32 bits: n hi is exact) + lo,
//   q = r + r^2 (1/2 + r/6 + r^2 (1/24 + r/120))  (truncation r^6/720 < 3.5e-17),  result = fma(T, q, T) * 2^k.
// <= 1 ulp against the true exponential on [-708, 0] (T and the final fma are the only roundings that matter), exact at 0
// (restated in numpy in tools/exp_study.py).  Below -708 the true value is a denormal < 2.5e-308: returned as 0.  NaN propagates.
static __device__ const double EXP_TABLE_64[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951,
};
__device__ __forceinline__ double exp_nonpos(double x) {
    const double xc = fmax(x, -708.0);
    const double t = fma(xc, 92.33248261689366, 6755399441055744.0);
    const int n = __double2loint(t);                           // 64 k + j, in [-65372, 0]
    const double nd = t - 6755399441055744.0;
    double r = fma(nd, -0.01083042469326756, xc);              // ln2/64, leading 32 bits
    r = fma(nd, -2.9815858269852933e-12, r);                   // ... and the rest
    const double r2 = r * r;
    const double a1 = fma(1.6666666666666666e-01, r, 0.5);
    const double a2 = fma(8.333333333333333e-03, r, 4.1666666666666664e-02);
    const double q = fma(fma(a2, r2, a1), r2, r);
    const double T = __ldg(&EXP_TABLE_64[n & 63]);
    const double p = fma(T, q, T);                              // in (0.99, 2)
    double res = __hiloint2double(__double2hiint(p) + ((n >> 6) << 20), __double2loint(p));   // * 2^k, k in [-1022, 0]
    res = x < -708.0 ? 0.0 : res;
    return x != x ? x : res;
}

// base kernel as sklearn evaluates it from the squared scaled distance
template <int KIND> __device__ __forceinline__ double base_kernel(double d2) {
    if (KIND == K_RBF) return exp_nonpos(-0.5 * d2);
    const double r = sqrt(d2);
    if (KIND == K_M12) return exp_nonpos(-r);
    if (KIND == K_M32) {
        const double k = r * 1.7320508075688772;  // math.sqrt(3)
        return (1.0 + k) * exp_nonpos(-k);
    }
    const double k = r * 2.23606797749979;        // math.sqrt(5)
    return __dadd_rn(__dadd_rn(1.0, k), __ddiv_rn(__dmul_rn(k, k), 3.0)) * exp_nonpos(-k);
}

// scipy.special.ndtr (cephes): 0.5*erfc(-a/sqrt(2)) evaluated the way scipy branches it
__device__ __forceinline__ double ndtr_like_scipy(double a) {
    const double x = a * 0.70710678118654752440;
    const double z = fabs(x);
    if (z < 0.70710678118654752440) return 0.5 + 0.5 * erf(x);
    const double y = 0.5 * erfc(z);
    return x > 0.0 ? 1.0 - y : y;
}

// bopy/acquisition.py:83-85, 99-106, 123-128 on (mean, var); scipy.stats.norm NaN rule: scale > 0 or NaN
__device__ __forceinline__ double acquisition_value(int acq, double mean, double var, double eta, double kappa) {
    const double sd = sqrt(var);  // var < 0 -> NaN like np.sqrt
    if (acq == A_LCB) return __dadd_rn(mean, -__dmul_rn(kappa, sd));
    if (!(sd > 0.0)) return __longlong_as_double(0x7ff8000000000000LL);
    const double z = __ddiv_rn(__dadd_rn(eta, -mean), sd);
    const double cdf = ndtr_like_scipy(z);
    if (acq == A_POI) return __dadd_rn(1.0, -cdf);
    // pdf = exp(-z^2/2) / sqrt(2 pi) / sd
    const double pdf = __ddiv_rn(__ddiv_rn(exp(__ddiv_rn(-__dmul_rn(z, z), 2.0)), 2.5066282746310002), sd);
    return __dadd_rn(__dmul_rn(-var, pdf), __dmul_rn(__dadd_rn(mean, -eta), cdf));
}

template <typename T> struct VecOf;
template <> struct VecOf<double> { using type = double2; };
template <> struct VecOf<float> { using type = float4; };

// pack VEC consecutive register values into one 128-bit vector (explicit, so the array stays in registers)
__device__ __forceinline__ double2 pack_vec(const double* r) { return make_double2(r[0], r[1]); }
__device__ __forceinline__ float4 pack_vec(const float* r) { return make_float4(r[0], r[1], r[2], r[3]); }
__device__ __forceinline__ void unpack_vec(double* r, const double2& v) { r[0] = v.x; r[1] = v.y; }
__device__ __forceinline__ void unpack_vec(float* r, const float4& v) { r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; }

// ---------------------------------------------------------------------------------------------------------
// Policies: who owns which accumulator element, how operand tiles are laid out, how one tile is multiplied.
// A tile: logical [k][row] (k < KC, row < 128); B tile: logical [k][cand].  a_index / b_index give the
// physical element offset inside the 8 KB tile (and inside the [128][128] residual / workspace block rows).
// ---------------------------------------------------------------------------------------------------------
template <typename T> struct FmaPolicy {
    using Elem = T;
    static constexpr int VEC = Geo<T>::VEC, KC = Geo<T>::KC, CH = Geo<T>::CH;
    static constexpr int RI = 8, CJ = 8, CV = VEC;   // rows x candidates per thread, candidates per vector store
    static constexpr bool kSwizzled = false;
    int ty, tx, part;
    bool leader;
    __device__ explicit FmaPolicy(int tid) {
        const int lane = tid & 31, warp = tid >> 5;
        part = warp >> 1;
        ty = part * 4 + (lane >> 3);
        tx = (warp & 1) * 8 + (lane & 7);
        leader = (lane >> 3) == 0;
    }
    __device__ __forceinline__ int row_of(int i) const { return ty * VEC + (i / VEC) * (16 * VEC) + (i % VEC); }
    __device__ __forceinline__ int cand_of(int j) const { return tx * VEC + (j / VEC) * (16 * VEC) + (j % VEC); }
    __host__ __device__ static __forceinline__ int a_index(int k, int r) { return k * BM + r; }
    __host__ __device__ static __forceinline__ int b_index(int k, int c) { return k * BN + c; }
    // sum over the lanes that hold the same candidates but different rows
    __device__ __forceinline__ double reduce_rows(double v) const {
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        return v;
    }
    // acc += A (x) B over one tile; DIAG marks tile kc of the diagonal GEMM (dense here)
    template <bool DIAG>
    __device__ __forceinline__ void mma_tile(T (&acc)[RI][CJ], const T* __restrict__ As, const T* __restrict__ Bs,
                                             int /*kc*/) const {
        using V = typename VecOf<T>::type;
        constexpr int NV = 8 / VEC;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            T a[8], b[8];
#pragma unroll
            for (int g = 0; g < NV; ++g) {
                unpack_vec(&a[g * VEC], *reinterpret_cast<const V*>(&As[k * BM + ty * VEC + g * 16 * VEC]));
                unpack_vec(&b[g * VEC], *reinterpret_cast<const V*>(&Bs[k * BN + tx * VEC + g * 16 * VEC]));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
    }
};

// fp64 tensor-shaped MMA: D(8x8) += A(8x4) * B(4x8); a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// c0/c1 = C[lane/4][2*(lane%4) + 0/1]
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct DmmaPolicy {
    using Elem = double;
    static constexpr int VEC = 2, KC = 8, CH = 16;
    static constexpr int RI = 4, CJ = 16, CV = 2;
    static constexpr bool kSwizzled = true;
    int lane, rg, cg, part;
    int aoff[RI], boff[CJ / 2];   // fragment element offsets inside a tile for k = lane%4 (k+4: add 4*128)
    bool leader;
    __device__ explicit DmmaPolicy(int tid) {
        lane = tid & 31;
        const int warp = tid >> 5;
        rg = warp >> 1;   // owns the 8-row atoms {rg, 7-rg, 8+rg, 15-rg}: in the triangular diagonal GEMM atom a needs
                          // a+1 k-tiles, and each of the four sets sums to the same 34 tile-visits
        cg = warp & 1;    // candidates 64*cg .. 64*cg+63
        part = rg;
        leader = (lane >> 2) == 0;
        const int kq = lane & 3, q8 = lane >> 2;
#pragma unroll
        for (int i = 0; i < RI; ++i) aoff[i] = a_index(kq, 8 * atom(i) + q8);
#pragma unroll
        for (int jj = 0; jj < CJ / 2; ++jj) boff[jj] = b_index(kq, 64 * cg + 8 * jj + q8);
    }
    __device__ __forceinline__ int atom(int i) const { return (i & 2) * 4 + ((i & 1) ? 7 - rg : rg); }
    __device__ __forceinline__ int row_of(int i) const { return 8 * atom(i) + (lane >> 2); }
    __device__ __forceinline__ int cand_of(int j) const { return 64 * cg + 8 * (j >> 1) + 2 * (lane & 3) + (j & 1); }
    // XOR swizzle on the 128-wide axis keyed by k%4: the 4 x 8 fragment gather of a half-warp hits 32 banks
    __host__ __device__ static __forceinline__ int a_index(int k, int r) { return k * BM + (r ^ (4 * (k & 3))); }
    __host__ __device__ static __forceinline__ int b_index(int k, int c) { return k * BN + (c ^ (4 * (k & 3))); }
    __device__ __forceinline__ double reduce_rows(double v) const {
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        return v;
    }
    // rows (atoms) I0 .. RI-1 of the warp tile times one operand tile
    template <int I0>
    __device__ __forceinline__ void mma_rows(double (&acc)[RI][CJ], const double* __restrict__ As,
                                             const double* __restrict__ Bs) const {
#pragma unroll
        for (int s = 0; s < KC / 4; ++s) {
            double a[RI], b[CJ / 2];
#pragma unroll
            for (int jj = 0; jj < CJ / 2; ++jj) b[jj] = Bs[boff[jj] + s * 4 * BN];
#pragma unroll
            for (int i = I0; i < RI; ++i) a[i] = As[aoff[i] + s * 4 * BM];
#pragma unroll
            for (int i = I0; i < RI; ++i) {
#pragma unroll
                for (int jj = 0; jj < CJ / 2; ++jj) dmma_m8n8k4(acc[i][2 * jj], acc[i][2 * jj + 1], a[i], b[jj]);
            }
        }
    }
    // DIAG: inv(L_II) is lower triangular, so row atom a only needs the k tiles kc <= a.  The warp's atoms ascend with i, so
    // the atoms that still need tile kc are a suffix i0 .. RI-1; i0 is warp-uniform and selects one of RI fully unrolled
    // bodies by a real branch.  (Written as `if (atom(i) < kc) continue;` ptxas predicates the DMMAs and pads every
    // predicated-off one with NOPs: the skipped half of the triangle then still costs ~60 % of its time --
    // profiles/r02/ncu_sweep_tc_f32_c4_source.csv.gz.)
    template <bool DIAG>
    __device__ __forceinline__ void mma_tile(double (&acc)[RI][CJ], const double* __restrict__ As,
                                             const double* __restrict__ Bs, int kc) const {
        if (!DIAG) {
            mma_rows<0>(acc, As, Bs);
            return;
        }
        const int i0 = (atom(0) < kc) + (atom(1) < kc) + (atom(2) < kc) + (atom(3) < kc);
        switch (i0) {
            case 0: mma_rows<0>(acc, As, Bs); break;
            case 1: mma_rows<1>(acc, As, Bs); break;
            case 2: mma_rows<2>(acc, As, Bs); break;
            case 3: mma_rows<3>(acc, As, Bs); break;
            default: break;
        }
    }
};

// fp32 operands on the warp-level TF32 tensor instruction with the 3xTF32 split (error-compensated: x = hi + lo, both
// TF32; a.b ~ hi.hi + hi.lo + lo.hi, fp32 accumulate): mma.sync.m16n8k8 (SASS HMMA.1688.F32.TF32).  Same warp tile as
// DmmaPolicy (32 rows x 64 candidates, 64 accumulators per thread), operand tiles [k=16][128] floats with an XOR
// swizzle on the 128-wide axis keyed by k%4 so the 4 x 8 fragment gathers hit 32 banks.  Against the register-tiled
// FFMA policy this needs 5x fewer shared-memory wavefronts per MAC (24 LDS.32 per 16384 MACs of a warp against 128
// LDS.128 quarter-warp phases), which is what bounded that policy at ~55 % of the FP32 pipe.
// hi = the TF32 the tensor core reads from an fp32 register anyway (top 19 bits), lo = the exact remainder (the core
// reads its top 19 bits in turn): two full-rate ALU instructions per element.  (cvt.rna.tf32.f32 would round instead
// of truncate, but it issues at conversion-unit rate and made the split as expensive as the MMAs themselves.)
__device__ __forceinline__ void tf32_split(float x, unsigned& hi, unsigned& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void hmma_tf32_m16n8k8(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

struct Tf32x3Policy {
    using Elem = float;
    static constexpr int VEC = 4, KC = Geo<float>::KC, CH = Geo<float>::CH;   // 16 k per tile, 8 tiles per 128 columns
    static constexpr int RI = 4, CJ = 16, CV = 2;
    static constexpr bool kSwizzled = true;
    int lane, rg, cg, part;
    bool leader;
    __device__ explicit Tf32x3Policy(int tid) {
        lane = tid & 31;
        const int warp = tid >> 5;
        rg = warp >> 1;   // rows 32 rg .. 32 rg + 31
        cg = warp & 1;    // candidates 64 cg .. 64 cg + 63
        part = rg;
        leader = (lane >> 2) == 0;
    }
    // accumulator (i, j): m-tile i>>1, row half i&1 (the instruction's c0/c1 vs c2/c3), n-tile j>>1, column j&1
    __device__ __forceinline__ int row_of(int i) const { return 32 * rg + 16 * (i >> 1) + 8 * (i & 1) + (lane >> 2); }
    __device__ __forceinline__ int cand_of(int j) const { return 64 * cg + 8 * (j >> 1) + 2 * (lane & 3) + (j & 1); }
    __host__ __device__ static __forceinline__ int a_index(int k, int r) { return k * BM + (r ^ (8 * (k & 3))); }
    __host__ __device__ static __forceinline__ int b_index(int k, int c) { return k * BN + (c ^ (8 * (k & 3))); }
    __device__ __forceinline__ double reduce_rows(double v) const {
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        return v;
    }
    template <bool DIAG>
    __device__ __forceinline__ void mma_tile(float (&acc)[RI][CJ], const float* __restrict__ As,
                                             const float* __restrict__ Bs, int /*kc*/) const {
        const int kq = lane & 3, q8 = lane >> 2, swz = 8 * kq;   // k = kq (+4): k & 3 == kq for both
#pragma unroll
        for (int s = 0; s < KC / 8; ++s) {
            const float* const A0 = As + (8 * s + kq) * BM;       // k = 8 s + kq
            const float* const A1 = A0 + 4 * BM;                  // k + 4
            const float* const B0 = Bs + (8 * s + kq) * BN;
            const float* const B1 = B0 + 4 * BN;
            unsigned ah[2][4], al[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int r = 32 * rg + 16 * mt + q8;
                const float a[4] = {A0[r ^ swz], A0[(r + 8) ^ swz], A1[r ^ swz], A1[(r + 8) ^ swz]};
#pragma unroll
                for (int u = 0; u < 4; ++u) tf32_split(a[u], ah[mt][u], al[mt][u]);
            }
#pragma unroll
            for (int nt = 0; nt < CJ / 2; ++nt) {
                const int c = 64 * cg + 8 * nt + q8;
                const float b[2] = {B0[c ^ swz], B1[c ^ swz]};
                unsigned bh[2], bl[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) tf32_split(b[u], bh[u], bl[u]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    // the tensor core adds into its accumulator with truncation: the 8-term partial products are summed
                    // from zero and joined to the running fp32 sum by a round-to-nearest FADD (no drift over long k)
                    float c4[4] = {0.f, 0.f, 0.f, 0.f};
                    hmma_tf32_m16n8k8(c4, al[mt], bh);   // small terms first
                    hmma_tf32_m16n8k8(c4, ah[mt], bl);
                    hmma_tf32_m16n8k8(c4, ah[mt], bh);
                    acc[2 * mt][2 * nt] += c4[0];
                    acc[2 * mt][2 * nt + 1] += c4[1];
                    acc[2 * mt + 1][2 * nt] += c4[2];
                    acc[2 * mt + 1][2 * nt + 1] += c4[3];
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// Engine = (policy of the off-diagonal GEMM, policy of the diagonal GEMM).
//   Engine<DmmaPolicy, DmmaPolicy>          fp64 everywhere                                   (dtype f64)
//   Engine<FmaPolicy<double>, FmaPolicy<double>>  fp64 on the FMA pipe (A/B comparison only)
//   Engine<Tf32x3Policy, DmmaPolicy>        "mixed": 3xTF32 tensor MMAs (fp32-grade) for the n^2/2 off-diagonal work
//   (Engine<FmaPolicy<float>, DmmaPolicy>    the same with register-tiled FFMA: BOPY_B200_F32_ENGINE=fma), fp64 DMMA for the
//                                           diagonal solve; L_IJ and V are stored in fp32, inv(L_II) in fp64; the
//                                           residual tile lives in shared memory in fp64 (seeded with the fp64 K*),
//                                           fp32 partial sums span one 128-column block of L only   (dtype f32)
// All operand tiles are TILE_BYTES; block row I holds I*CHG off-diagonal tiles followed by CHD diagonal tiles.
// ---------------------------------------------------------------------------------------------------------
template <class PG_, class PD_> struct Engine {
    using PG = PG_;
    using PD = PD_;
    using TG = typename PG::Elem;
    using TD = typename PD::Elem;
    static constexpr bool kMixed = sizeof(TG) != sizeof(TD);
    static constexpr int CHG = PG::CH, CHD = PD::CH;
    __host__ __device__ static long long row_base(int I) {   // tiles stored before block row I
        return (long long)CHG * I * (I - 1) / 2 + (long long)CHD * I;
    }
    __host__ __device__ static long long total_tiles(int n_blocks) { return row_base(n_blocks); }
};
using EngineF64 = Engine<DmmaPolicy, DmmaPolicy>;
using EngineF64Fma = Engine<FmaPolicy<double>, FmaPolicy<double>>;
using EngineMixed = Engine<Tf32x3Policy, DmmaPolicy>;
using EngineMixedFma = Engine<FmaPolicy<float>, DmmaPolicy>;

constexpr int FLUSH_BLOCKS = 4;   // mixed engine: 128-column blocks of L per fp32 partial sum (256 terms)
constexpr size_t SMEM_LIMIT = 227 * 1024;
template <class E> constexpr size_t sweep_smem_base(int d) {
    return (size_t)2 * STAGES * TILE_BYTES + (size_t)BM * BN * sizeof(typename E::TD) +
           (size_t)d * BN * sizeof(double) + 192;
}
// the block row of X/l + alpha gets a buffer of its own (prefetched a row ahead) whenever shared memory allows
template <class E> constexpr bool sweep_xrow_separate(int d) {
    return sweep_smem_base<E>(d) + (size_t)(d + 1) * BM * sizeof(double) <= SMEM_LIMIT;
}
template <class E> constexpr size_t sweep_smem_bytes(int d) {
    return sweep_smem_base<E>(d) + (sweep_xrow_separate<E>(d) ? (size_t)(d + 1) * BM * sizeof(double) : 0);
}

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// store CV consecutive candidates of one row (converted to the destination element type)
__device__ __forceinline__ void store_cands(double* dst, const double* v, int) { *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]); }
__device__ __forceinline__ void store_cands(float* dst, const float* v, int) { *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void store_cands(float* dst, const double* v, int) {
    *reinterpret_cast<float2*>(dst) = make_float2(static_cast<float>(v[0]), static_cast<float>(v[1]));
}

template <class E, int KIND>
__global__ void __launch_bounds__(NT_ALL, 1) sweep_kernel(const SweepParams p) {
    using PG = typename E::PG;
    using PD = typename E::PD;
    using TG = typename E::TG;
    using TD = typename E::TD;
    constexpr int CHG = E::CHG, CHD = E::CHD;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* const stA = smem_raw;                                 // [STAGES] A tiles (L_IJ or inv(L_II))
    unsigned char* const stB = smem_raw + STAGES * TILE_BYTES;           // [STAGES] B tiles (V_J)
    unsigned char* const rs_raw = smem_raw + 2 * STAGES * TILE_BYTES;
    TD* const Rs = reinterpret_cast<TD*>(rs_raw);                        // [BM][BN] residual tile (B operand of the diagonal GEMM)
    // [(d+1)][BM] X/l block row + alpha: own buffer behind the barriers when it fits, else aliasing Rs
    double* const xrow_alias = reinterpret_cast<double*>(rs_raw);
    double* const partM = reinterpret_cast<double*>(rs_raw + 48 * 1024); // aliases Rs: [4][BN]
    double* const partS = reinterpret_cast<double*>(rs_raw + 52 * 1024); // aliases Rs: [4][BN]
    double* const xs_s = reinterpret_cast<double*>(rs_raw + (size_t)BM * BN * sizeof(TD));  // [d][BN] candidates / l
    unsigned char* const tail = reinterpret_cast<unsigned char*>(xs_s + (size_t)p.d * BN);
    uint64_t* const full = reinterpret_cast<uint64_t*>(tail);            // [STAGES] producer -> consumers
    uint64_t* const empty = full + STAGES;                               // [STAGES] consumers -> producer
    uint64_t* const xbar = empty + STAGES;                               // block row of X/l + alpha landed
    uint64_t* const vbar = xbar + 1;                                     // V_I published to the workspace
    MinLoc* const red = reinterpret_cast<MinLoc*>(tail + 128);           // [4]
    double* const xrow = p.xrow_separate ? reinterpret_cast<double*>(tail + 192) : xrow_alias;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_pad = p.n_blocks * BM;
    const unsigned char* const Lt = reinterpret_cast<const unsigned char*>(p.Lt);
    const long long slot_bytes = (long long)n_pad * BN * sizeof(TG);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NT / 32);
        }
        mbar_init(xbar, 1);
        mbar_init(vbar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    if (warp >= NT / 32) {
        // =============================== TMA producer (one lane) =========================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
        if (warp != NT / 32 || lane != 0) return;
        uint32_t g = 0, vphase = 0;
        for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const unsigned char* const Vt =
                reinterpret_cast<const unsigned char*>(p.Vws) + (p.slot_per_tile ? tile : (long long)blockIdx.x) * slot_bytes;
            for (int I = 0; I < p.n_blocks; ++I) {
                const int T_gemm = I * CHG, T_all = T_gemm + CHD;
                const unsigned char* const a_row = Lt + E::row_base(I) * TILE_BYTES;
                for (int t = 0; t < T_all; ++t, ++g) {
                    const uint32_t stage = g % STAGES;
                    mbar_wait(&empty[stage], ((g / STAGES) & 1u) ^ 1u);
                    if (t < T_gemm) {
                        // J order: even rows 0..I-1, odd rows I-2..0 then I-1 (zig-zag: the V slices read last by one
                        // block row are read first by the next, so they are still in L2; V_{I-1} always comes last)
                        const int jpos = t / CHG, c = t - jpos * CHG;
                        const int J = ((I & 1) && p.zigzag) ? (jpos < I - 1 ? I - 2 - jpos : I - 1) : jpos;
                        if (J == I - 1 && c == 0) {   // first touch of V_{I-1}: wait until the consumers published it
                            mbar_wait(vbar, vphase);
                            vphase ^= 1u;
                        }
                        const long long tt = (long long)J * CHG + c;
                        mbar_arrive_expect_tx(&full[stage], 2 * TILE_BYTES);
                        bulk_g2s(stA + stage * TILE_BYTES, a_row + tt * TILE_BYTES, TILE_BYTES, &full[stage]);
                        bulk_g2s(stB + stage * TILE_BYTES, Vt + tt * TILE_BYTES, TILE_BYTES, &full[stage]);
                    } else {
                        mbar_arrive_expect_tx(&full[stage], TILE_BYTES);
                        bulk_g2s(stA + stage * TILE_BYTES, a_row + (long long)t * TILE_BYTES, TILE_BYTES, &full[stage]);
                    }
                }
            }
        }
        return;
    }

    // ===================================== compute warps ==================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_COMPUTE));
    const PG pg(tid);
    const PD pd(tid);
    uint32_t gcount = 0;   // operand tiles consumed so far (selects ring stage and barrier parity)
    uint32_t xphase = 0;
    MinLoc best;
    best.val = 0.0;
    best.idx = -1;

    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const long long c0 = tile * BN;
        TG* const Vt = reinterpret_cast<TG*>(reinterpret_cast<unsigned char*>(p.Vws) +
                                             (p.slot_per_tile ? tile : (long long)blockIdx.x) * slot_bytes);

        auto issue_xrow = [&](int I) {
            const uint32_t bytes = (uint32_t)(p.d + 1) * BM * sizeof(double);
            mbar_arrive_expect_tx(xbar, bytes);
            bulk_g2s(xrow, p.Xt + (long long)I * (p.d + 1) * BM, bytes, xbar);
        };

        if (tid == 0 && !(p.xrow_separate && tile != blockIdx.x)) issue_xrow(0);   // (prefetched by the previous tile)
        // stage this tile's candidates, scaled like sklearn does (X / length_scale), dimension-major
        for (int e = tid; e < BN * p.d; e += NT) {
            const int c = e / p.d, q = e - c * p.d;
            const long long gc = c0 + c;
            const double v = gc < p.m ? p.Xs[gc * p.d + q] : 0.0;
            xs_s[q * BN + c] = __ddiv_rn(v, p.ls[q]);
        }
        double mean_c = 0.0, ss_c = 0.0;  // per-candidate accumulators, threads 0..BN-1
        consumer_sync();

        for (int I = 0; I < p.n_blocks; ++I) {
            const int T_gemm = I * CHG;

            // ---- kernel tile K*[block row I, this tile's candidates] and its share of the mean --------
            TG acc[PG::RI][PG::CJ];
            mbar_wait(xbar, xphase);
            xphase ^= 1;
            {
                // two halves of the thread's rows: 32 squared distances live at a time leave the scheduler enough
                // registers to interleave the (branch-free) kernel evaluations
                constexpr int RH = PG::RI / 2;
                double kq[E::kMixed ? PG::RI : 1][E::kMixed ? PG::CJ : 1];   // mixed engine: K* stays fp64
                double mp[PG::CJ];
#pragma unroll
                for (int j = 0; j < PG::CJ; ++j) mp[j] = 0.0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double d2[RH][PG::CJ];
#pragma unroll
                    for (int i = 0; i < RH; ++i)
#pragma unroll
                        for (int j = 0; j < PG::CJ; ++j) d2[i][j] = 0.0;
                    for (int q = 0; q < p.d; ++q) {
                        double xr[RH], xc[PG::CJ];
#pragma unroll
                        for (int i = 0; i < RH; ++i) xr[i] = xrow[q * BM + pg.row_of(h * RH + i)];
#pragma unroll
                        for (int j = 0; j < PG::CJ; ++j) xc[j] = xs_s[q * BN + pg.cand_of(j)];
#pragma unroll
                        for (int i = 0; i < RH; ++i)
#pragma unroll
                            for (int j = 0; j < PG::CJ; ++j) {
                                const double df = xc[j] - xr[i];
                                d2[i][j] = fma(df, df, d2[i][j]);   // cdist's summation order over the dimensions
                            }
                    }
#pragma unroll
                    for (int i = 0; i < RH; ++i) {
                        const int row = pg.row_of(h * RH + i);
                        // rows beyond n (identity padding) get amplitude 0: no branch, so that all the kernel
                        // evaluations of a half stay in one basic block and the scheduler interleaves them
                        const double amp_i = (I * BM + row < p.n) ? p.amp : 0.0;
                        const double a_i = xrow[p.d * BM + row];
#pragma unroll
                        for (int j = 0; j < PG::CJ; ++j) {
                            const double kv = __dmul_rn(amp_i, base_kernel<KIND>(d2[i][j]));
                            if constexpr (E::kMixed) kq[h * RH + i][j] = kv;
                            acc[h * RH + i][j] = E::kMixed ? static_cast<TG>(0) : static_cast<TG>(kv);
                            mp[j] = fma(kv, a_i, mp[j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < PG::CJ; ++j) mp[j] = pg.reduce_rows(mp[j]);
                if (pg.leader) {
#pragma unroll
                    for (int j = 0; j < PG::CJ; ++j) partM[pg.part * BN + pg.cand_of(j)] = mp[j];
                }
                consumer_sync();
                if (tid < BN) mean_c += ((partM[tid] + partM[BN + tid]) + partM[2 * BN + tid]) + partM[3 * BN + tid];
                consumer_sync();  // xrow / partM consumed: Rs may be overwritten from here on
                if (p.xrow_separate && tid == 0) {   // prefetch the next block row (or the next tile's first) of X/l
                    if (I + 1 < p.n_blocks) issue_xrow(I + 1);
                    else if (tile + gridDim.x < p.ntiles) issue_xrow(0);
                }
                if constexpr (E::kMixed) {
                    // fp64 residual tile R lives in shared memory (diagonal policy's B layout); every thread owns the
                    // same elements throughout, so the read-modify-write flushes below need no barrier
#pragma unroll
                    for (int i = 0; i < PG::RI; ++i) {
                        const int row = pg.row_of(i);
#pragma unroll
                        for (int jv = 0; jv < PG::CJ / 2; ++jv)
                            *reinterpret_cast<double2*>(&Rs[PD::b_index(row, pg.cand_of(jv * 2))]) =
                                make_double2(kq[i][jv * 2], kq[i][jv * 2 + 1]);
                    }
                }
            }

            // ---- R_I = K*_I - sum_J L_IJ V_J ---------------------------------------------------------------
            for (int t = 0; t < T_gemm; ++t, ++gcount) {
                const uint32_t stage = gcount % STAGES;
                mbar_wait(&full[stage], (gcount / STAGES) & 1u);
                pg.template mma_tile<false>(acc, reinterpret_cast<const TG*>(stA + stage * TILE_BYTES),
                                            reinterpret_cast<const TG*>(stB + stage * TILE_BYTES), -1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);   // this warp is done with the stage
                if constexpr (E::kMixed) {
                    // fp32 partial sums cover FLUSH_BLOCKS 128-column blocks of L only; they are folded into the fp64 tile
                    if ((t + 1) % (FLUSH_BLOCKS * CHG) == 0 || t + 1 == T_gemm) {
#pragma unroll
                        for (int i = 0; i < PG::RI; ++i) {
                            const int row = pg.row_of(i);
#pragma unroll
                            for (int jv = 0; jv < PG::CJ / 2; ++jv) {
                                double2* const dst = reinterpret_cast<double2*>(&Rs[PD::b_index(row, pg.cand_of(jv * 2))]);
                                double2 r = *dst;
                                r.x += static_cast<double>(acc[i][jv * 2]);
                                r.y += static_cast<double>(acc[i][jv * 2 + 1]);
                                *dst = r;
                                acc[i][jv * 2] = acc[i][jv * 2 + 1] = static_cast<TG>(0);
                            }
                        }
                    }
                }
            }

            // ---- the residual tile becomes the B operand of the diagonal GEMM (in the diagonal policy's layout) ----
            if constexpr (!E::kMixed) {
#pragma unroll
                for (int i = 0; i < PG::RI; ++i) {
                    const int row = pg.row_of(i);
#pragma unroll
                    for (int jv = 0; jv < PG::CJ / PG::CV; ++jv)
                        store_cands(reinterpret_cast<TG*>(Rs) + PD::b_index(row, pg.cand_of(jv * PG::CV)),
                                    &acc[i][jv * PG::CV], 0);
                }
            }
            consumer_sync();

            // ---- V_I = inv(L_II) R_I ------------------------------------------------------------------------------
            TD accd[PD::RI][PD::CJ];
#pragma unroll
            for (int i = 0; i < PD::RI; ++i)
#pragma unroll
                for (int j = 0; j < PD::CJ; ++j) accd[i][j] = static_cast<TD>(0);
            for (int kc = 0; kc < CHD; ++kc, ++gcount) {
                const uint32_t stage = gcount % STAGES;
                mbar_wait(&full[stage], (gcount / STAGES) & 1u);
                pd.template mma_tile<true>(accd, reinterpret_cast<const TD*>(stA + stage * TILE_BYTES),
                                           Rs + kc * PD::KC * BN, kc);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
            }

            // ---- V_I: publish to the workspace (GEMM policy's B layout), fold into sum v^2 -------------------------
            {
                const bool publish = (I + 1 < p.n_blocks) || p.slot_per_tile;
                TG* const Vrow = Vt + (long long)I * BM * BN;
                double sq[PD::CJ];
#pragma unroll
                for (int j = 0; j < PD::CJ; ++j) sq[j] = 0.0;
#pragma unroll
                for (int i = 0; i < PD::RI; ++i) {
                    if (publish) {
                        const int row = pd.row_of(i);
#pragma unroll
                        for (int jv = 0; jv < PD::CJ / PD::CV; ++jv)
                            store_cands(&Vrow[PG::b_index(row, pd.cand_of(jv * PD::CV))], &accd[i][jv * PD::CV], 0);
                    }
#pragma unroll
                    for (int j = 0; j < PD::CJ; ++j) {
                        const double v = static_cast<double>(accd[i][j]);
                        sq[j] = fma(v, v, sq[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < PD::CJ; ++j) sq[j] = pd.reduce_rows(sq[j]);
                if (pd.leader) {
#pragma unroll
                    for (int j = 0; j < PD::CJ; ++j) partS[pd.part * BN + pd.cand_of(j)] = sq[j];
                }
                fence_proxy_async();  // V stores (generic proxy) before the producer's bulk loads of them
            }
            consumer_sync();   // every warp finished the diagonal GEMM (Rs reads) and published its part of V_I
            if (tid == 0 && I + 1 < p.n_blocks) {
                mbar_arrive(vbar);                          // producer may now load V_I
                if (!p.xrow_separate) issue_xrow(I + 1);    // lands in the (now free) Rs region
            }
            if (tid < BN) ss_c += ((partS[tid] + partS[BN + tid]) + partS[2 * BN + tid]) + partS[3 * BN + tid];
        }

        // ---- epilogue: de-normalise, acquisition, arg-min -------------------------------------------------------
        MinLoc mine;
        mine.val = 0.0;
        mine.idx = -1;
        if (tid < BN) {
            const long long gc = c0 + tid;
            if (gc < p.m) {
                const double mean = __dadd_rn(__dmul_rn(p.y_std, mean_c), p.y_mean);
                const double var = __dmul_rn(__dadd_rn(p.kss, -ss_c), p.y_var);
                if (p.mean_out) p.mean_out[gc] = mean;
                if (p.var_out) p.var_out[gc] = var;
                if (p.acq != A_NONE) {
                    const double a = acquisition_value(p.acq, mean, var, p.eta, p.kappa);
                    if (p.acq_out) p.acq_out[gc] = a;
                    if (!(p.nan_skip && a != a)) {
                        mine.val = a;
                        mine.idx = p.index_base + gc;
                    }
                }
            }
        }
        if (p.partials != nullptr || p.tile_records != nullptr) {
            if (warp < BN / 32) {
                mine = minloc_warp_reduce(mine);
                if (lane == 0) red[warp] = mine;
            }
            consumer_sync();
            if (tid == 0) {
                MinLoc tbest = red[0];
                for (int w = 1; w < BN / 32; ++w)
                    if (minloc_better(red[w], tbest)) tbest = red[w];
                if (p.tile_records != nullptr) p.tile_records[tile] = tbest;
                if (minloc_better(tbest, best)) best = tbest;
            }
        }
        consumer_sync();  // xs_s, part buffers and red are reused by the next tile
    }
    if (tid == 0 && p.partials != nullptr) p.partials[blockIdx.x] = best;
}

// reduce the per-CTA records to the final (value, index); one block
__global__ void minloc_finalize_kernel(const MinLoc* __restrict__ parts, int nparts, double* val_out,
                                       long long* idx_out) {
    __shared__ MinLoc red[32];
    MinLoc v;
    v.val = 0.0;
    v.idx = -1;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x)
        if (minloc_better(parts[i], v)) v = parts[i];
    v = minloc_warp_reduce(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (minloc_better(red[w], v)) v = red[w];
        if (val_out) *val_out = v.val;
        if (idx_out) *idx_out = v.idx;
    }
}

}  // namespace bopy
