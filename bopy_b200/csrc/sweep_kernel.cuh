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
//   R_I = K*_I - sum_{J<I} L_IJ V_J     register-tiled FMA GEMM, operands streamed by bulk async copies
//   V_I = inv(L_II) R_I                 second GEMM against the pre-inverted 128x128 diagonal block
// L is stored pre-tiled ([k][row] tiles of 8 KB, off-diagonal tiles negated, diagonal blocks inverted) so
// that every operand tile is ONE contiguous cp.async.bulk (TMA) transfer; V_J tiles live in an L2-resident
// per-CTA workspace in the same tile layout.
#pragma once
#include "common.cuh"

namespace bopy {

enum { K_RBF = 0, K_M12 = 1, K_M32 = 2, K_M52 = 3 };
enum { A_NONE = -1, A_LCB = 0, A_EI = 1, A_POI = 2 };

struct SweepParams {
    const void* Lt;     // packed factor tiles, tile (I, Jc) at ((CH*I*(I+1)/2 + Jc) * KC*BM), layout [k][row]
    const double* Xt;   // [n_blocks][d+1][BM]: X/l (dimension-major) then alpha, zero padded
    void* Vws;          // [slots][n_pad][BN] solve workspace
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
};

// base kernel as sklearn evaluates it from the squared scaled distance
template <int KIND> __device__ __forceinline__ double base_kernel(double d2) {
    if (KIND == K_RBF) return exp(-0.5 * d2);
    const double r = sqrt(d2);
    if (KIND == K_M12) return exp(-r);
    if (KIND == K_M32) {
        const double k = r * 1.7320508075688772;  // math.sqrt(3)
        return (1.0 + k) * exp(-k);
    }
    const double k = r * 2.23606797749979;        // math.sqrt(5)
    return __dadd_rn(__dadd_rn(1.0, k), __ddiv_rn(__dmul_rn(k, k), 3.0)) * exp(-k);
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

// acc[8][8] += A[k][rows] (x) B[k][cands] over one tile of KC contraction steps
template <typename T>
__device__ __forceinline__ void mma_chunk(T (&acc)[8][8], const T* __restrict__ As, const T* __restrict__ Bs,
                                          int ty, int tx) {
    using G = Geo<T>;
    using V = typename VecOf<T>::type;
    constexpr int VEC = G::VEC, KC = G::KC, NV = 8 / VEC;
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

template <typename T> constexpr size_t sweep_smem_bytes(int d) {
    return (size_t)2 * STAGES * TILE_BYTES + (size_t)BM * BN * sizeof(T) + (size_t)d * BN * sizeof(double) + 128;
}

template <typename T, int KIND>
__global__ void __launch_bounds__(NT, 1) sweep_kernel(const SweepParams p) {
    using G = Geo<T>;
    using V = typename VecOf<T>::type;
    constexpr int VEC = G::VEC, KC = G::KC, CH = G::CH, NV = 8 / VEC;
    constexpr int TE = TILE_BYTES / sizeof(T);  // elements per operand tile

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* const stA = reinterpret_cast<T*>(smem_raw);
    T* const stB = reinterpret_cast<T*>(smem_raw + STAGES * TILE_BYTES);
    unsigned char* const rs_raw = smem_raw + 2 * STAGES * TILE_BYTES;
    T* const Rs = reinterpret_cast<T*>(rs_raw);                       // [BM][BN] residual tile (B operand of the diagonal GEMM)
    double* const xrow = reinterpret_cast<double*>(rs_raw);             // aliases Rs: [(d+1)][BM] X/l block row + alpha
    double* const partM = reinterpret_cast<double*>(rs_raw + 48 * 1024);  // aliases Rs: [4][BN]
    double* const partS = reinterpret_cast<double*>(rs_raw + 52 * 1024);  // aliases Rs: [4][BN]
    double* const xs_s = reinterpret_cast<double*>(rs_raw + (size_t)BM * BN * sizeof(T));  // [d][BN] candidates / l
    unsigned char* const tail = reinterpret_cast<unsigned char*>(xs_s + (size_t)p.d * BN);
    uint64_t* const full = reinterpret_cast<uint64_t*>(tail);           // [STAGES]
    uint64_t* const xbar = full + STAGES;
    MinLoc* const red = reinterpret_cast<MinLoc*>(tail + 64);           // [4]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wy = warp >> 1;
    const int ty = wy * 4 + (lane >> 3);
    const int tx = (warp & 1) * 8 + (lane & 7);
    const int n_pad = p.n_blocks * BM;
    const T* const Lt = reinterpret_cast<const T*>(p.Lt);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_init(xbar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t gcount = 0;   // operand tiles consumed so far (selects ring stage and barrier parity)
    uint32_t xphase = 0;
    MinLoc best;
    best.val = 0.0;
    best.idx = -1;

    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const long long c0 = tile * BN;
        T* const Vt = reinterpret_cast<T*>(p.Vws) + (p.slot_per_tile ? tile : (long long)blockIdx.x) * n_pad * BN;

        auto issue_xrow = [&](int I) {
            const uint32_t bytes = (uint32_t)(p.d + 1) * BM * sizeof(double);
            mbar_arrive_expect_tx(xbar, bytes);
            bulk_g2s(xrow, p.Xt + (long long)I * (p.d + 1) * BM, bytes, xbar);
        };
        auto issue = [&](int I, int t, uint32_t gc) {
            const uint32_t stage = gc % STAGES;
            const T* a_src = Lt + ((long long)CH * I * (I + 1) / 2 + t) * TE;
            if (t < I * CH) {
                mbar_arrive_expect_tx(&full[stage], 2 * TILE_BYTES);
                bulk_g2s(stA + stage * TE, a_src, TILE_BYTES, &full[stage]);
                bulk_g2s(stB + stage * TE, Vt + (long long)t * TE, TILE_BYTES, &full[stage]);
            } else {
                mbar_arrive_expect_tx(&full[stage], TILE_BYTES);
                bulk_g2s(stA + stage * TE, a_src, TILE_BYTES, &full[stage]);
            }
        };

        if (tid == 0) issue_xrow(0);
        // stage this tile's candidates, scaled like sklearn does (X / length_scale), dimension-major
        for (int e = tid; e < BN * p.d; e += NT) {
            const int c = e / p.d, q = e - c * p.d;
            const long long gc = c0 + c;
            const double v = gc < p.m ? p.Xs[gc * p.d + q] : 0.0;
            xs_s[q * BN + c] = __ddiv_rn(v, p.ls[q]);
        }
        double mean_c = 0.0, ss_c = 0.0;  // per-candidate accumulators, threads 0..BN-1
        __syncthreads();

        for (int I = 0; I < p.n_blocks; ++I) {
            const int T_gemm = I * CH, T_all = T_gemm + CH;
            if (tid == 0) {
                const int pre = T_all < STAGES ? T_all : STAGES;
                for (int t = 0; t < pre; ++t) issue(I, t, gcount + t);
            }

            // ---- kernel tile K*[block row I, this tile's candidates] and its share of the mean --------
            T acc[8][8];
            mbar_wait(xbar, xphase);
            xphase ^= 1;
            {
                double d2[8][8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) d2[i][j] = 0.0;
                for (int q = 0; q < p.d; ++q) {
                    double xr[8], xc[8];
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        xr[s] = xrow[q * BM + owned<T>(ty, s)];
                        xc[s] = xs_s[q * BN + owned<T>(tx, s)];
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const double df = __dadd_rn(xc[j], -xr[i]);
                            d2[i][j] = __dadd_rn(d2[i][j], __dmul_rn(df, df));  // cdist order, unfused
                        }
                }
                double mp[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) mp[j] = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = owned<T>(ty, i);
                    const bool live = I * BM + row < p.n;
                    const double a_i = xrow[p.d * BM + row];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double kv = live ? __dmul_rn(p.amp, base_kernel<KIND>(d2[i][j])) : 0.0;
                        acc[i][j] = static_cast<T>(kv);
                        mp[j] = fma(kv, a_i, mp[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    mp[j] += __shfl_xor_sync(0xffffffffu, mp[j], 8);
                    mp[j] += __shfl_xor_sync(0xffffffffu, mp[j], 16);
                }
                if ((lane >> 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) partM[wy * BN + owned<T>(tx, j)] = mp[j];
                }
            }
            __syncthreads();
            if (tid < BN) mean_c += ((partM[tid] + partM[BN + tid]) + partM[2 * BN + tid]) + partM[3 * BN + tid];
            __syncthreads();  // xrow / partM consumed: Rs may be overwritten from here on

            // ---- R_I = K*_I - sum_J L_IJ V_J, then V_I = inv(L_II) R_I ------------------------------------
            for (int t = 0; t < T_all; ++t) {
                if (t == T_gemm) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
#pragma unroll
                        for (int g = 0; g < NV; ++g)
                            *reinterpret_cast<V*>(&Rs[owned<T>(ty, i) * BN + tx * VEC + g * 16 * VEC]) =
                                pack_vec(&acc[i][g * VEC]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = static_cast<T>(0);
                    }
                    __syncthreads();
                }
                const uint32_t stage = gcount % STAGES;
                mbar_wait(&full[stage], (gcount / STAGES) & 1u);
                const T* Bs = t < T_gemm ? stB + stage * TE : Rs + (t - T_gemm) * KC * BN;
                mma_chunk<T>(acc, stA + stage * TE, Bs, ty, tx);
                __syncthreads();  // every warp is done with this stage (and with this slice of Rs)
                if (tid == 0 && t + STAGES < T_all) issue(I, t + STAGES, gcount + STAGES);
                ++gcount;
            }

            // ---- V_I: publish to the workspace, fold into sum v^2 --------------------------------------------
            if (tid == 0 && I + 1 < p.n_blocks) issue_xrow(I + 1);
            {
                const bool publish = (I + 1 < p.n_blocks) || p.slot_per_tile;
                T* const Vrow = Vt + (long long)I * BM * BN;
                double sq[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) sq[j] = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (publish) {
#pragma unroll
                        for (int g = 0; g < NV; ++g)
                            *reinterpret_cast<V*>(&Vrow[owned<T>(ty, i) * BN + tx * VEC + g * 16 * VEC]) =
                                pack_vec(&acc[i][g * VEC]);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double v = static_cast<double>(acc[i][j]);
                        sq[j] = fma(v, v, sq[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], 8);
                    sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], 16);
                }
                if ((lane >> 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) partS[wy * BN + owned<T>(tx, j)] = sq[j];
                }
                fence_proxy_async();  // V stores (generic proxy) before the bulk loads of the next block row
            }
            __syncthreads();
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
                    mine.val = a;
                    mine.idx = p.index_base + gc;
                }
            }
        }
        if (p.partials != nullptr) {
            if (warp < BN / 32) {
                mine = minloc_warp_reduce(mine);
                if (lane == 0) red[warp] = mine;
            }
            __syncthreads();
            if (tid == 0) {
                for (int w = 0; w < BN / 32; ++w)
                    if (minloc_better(red[w], best)) best = red[w];
            }
        }
        __syncthreads();  // xs_s, part buffers and red are reused by the next tile
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
