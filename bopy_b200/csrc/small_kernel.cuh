// The same path for SMALL training sets (n <= 32: BASELINE configs C1 / C2, the reference's examples and tests):
// one THREAD per candidate, nothing padded to 128-row blocks.
//
//   k_i  = amplitude * base_kernel(|x/l - X_i/l|^2)          i < n       $SK/_gpr.py:446, kernels.py:1569-1570 / 1720-1729
//   mean = y_std * sum_i k_i alpha_i + y_mean                             $SK/_gpr.py:447-450
//   v    = inv(L) k  (the one inverted diagonal block the packed state already holds; lower triangular: a mat-vec with
//          no sequential dependency, the same arithmetic the blocked kernels use for their diagonal blocks)   :460-462
//   var  = (k(x,x) - sum_i v_i^2) * y_std^2, not clamped                  diagonal of :466-469
//   LCB / EI / POI, arg-min with np.argmin's (or np.nanargmin's) rules    bopy/acquisition.py:83-128
//
// The blocked kernels spend 128 kernel evaluations and a 128^3/2 DMMA product per candidate tile whatever n is; at n = 10
// that is > 40x the algorithmic work.  Here a candidate costs n exps, n(n+1)/2 + n(3d+5) FMAs and the epilogue, so the
// kernel is bound by FP64 issue for the exps (~40 slots each), far from HBM (8 d bytes per candidate).
// inv(L) and X/l, alpha sit in shared memory and are read as warp-wide broadcasts; k and the squared distances live in
// registers (NP = n rounded up to 8 / 16 / 32 is a template parameter so that they are statically indexed).
#pragma once
#include "sweep_kernel.cuh"

namespace bopy {

constexpr int SMALL_N_MAX = 32;
constexpr int SMALL_CTAS_PER_SM = 8;   // the grid is capped at this many thread blocks per SM (persistent loop over the tiles)
constexpr int SMALL_INLINE_M = 8;  // candidates a host-buffer call can pass as kernel parameters
constexpr int SMALL_NT = 128;      // threads = candidates per tile (= BN, so that per-tile records line up with the blocked kernels)

struct SmallParams {
    const double* Xt;      // block 0 of the packed X: [(d+1)][BM], X/l dimension-major then alpha
    const double* Dinv;    // block 0 of the inverted diagonal blocks: [BM][BM] row-major, lower triangular
    const double* Xs;      // candidates (m, d) row-major, or nullptr: they are in xs_inline (m <= SMALL_INLINE_M)
    long long m, ntiles;
    int n, d;
    double ls[MAX_D];
    double amp, kss, y_mean, y_std, y_var;
    int acq;
    double eta, kappa;
    double* mean_out;
    double* var_out;
    double* acq_out;
    long long index_base;
    MinLoc* partials;       // [gridDim.x] or nullptr
    MinLoc* tile_records;   // [ntiles] or nullptr
    int nan_skip;
    // host-buffer entry (bopy_acq_eval_host, the DIRECT probe of the reference's own examples: n ~ 10, one point per call):
    // candidates as kernel parameters, outputs in mapped pinned host memory -- no memcpy call around the launch
    double xs_inline[SMALL_INLINE_M * MAX_D];
};

template <int NP, int KIND>
__global__ void __launch_bounds__(SMALL_NT) small_n_kernel(const SmallParams p) {
    __shared__ double Ds[NP * (NP + 1) / 2];   // packed lower triangle of inv(L): row i at i (i + 1) / 2
    __shared__ double Xn[MAX_D][NP];           // X/l
    __shared__ double al[NP];
    __shared__ MinLoc red[SMALL_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < NP * (NP + 1) / 2; e += SMALL_NT) {
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int j = e - i * (i + 1) / 2;
        Ds[e] = (i < p.n && j < p.n) ? p.Dinv[(size_t)i * BM + j] : 0.0;
    }
    for (int e = tid; e < p.d * NP; e += SMALL_NT) {
        const int q = e / NP, i = e - q * NP;
        Xn[q][i] = p.Xt[q * BM + i];
    }
    if (tid < NP) al[tid] = p.Xt[p.d * BM + tid];
    __syncthreads();

    MinLoc best;
    best.val = 0.0;
    best.idx = -1;
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const long long gc = tile * SMALL_NT + tid;
        MinLoc mine;
        mine.val = 0.0;
        mine.idx = -1;
        if (gc < p.m) {
            double d2[NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) d2[i] = 0.0;
            for (int q = 0; q < p.d; ++q) {
                const double xv = p.Xs != nullptr ? p.Xs[gc * p.d + q] : p.xs_inline[gc * p.d + q];
                const double xq = __ddiv_rn(xv, p.ls[q]);   // X / length_scale, like sklearn
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const double df = xq - Xn[q][i];
                    d2[i] = fma(df, df, d2[i]);   // cdist's summation order over the dimensions
                }
            }
            double k[NP];
            double mean_c = 0.0;
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                k[i] = i < p.n ? __dmul_rn(p.amp, base_kernel<KIND>(d2[i])) : 0.0;
                mean_c = fma(k[i], al[i], mean_c);
            }
            double ss = 0.0;
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j <= i; ++j) v = fma(Ds[i * (i + 1) / 2 + j], k[j], v);
                ss = fma(v, v, ss);
            }
            const double mean = __dadd_rn(__dmul_rn(p.y_std, mean_c), p.y_mean);
            const double var = __dmul_rn(__dadd_rn(p.kss, -ss), p.y_var);
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
        if (p.partials != nullptr || p.tile_records != nullptr) {
            mine = minloc_warp_reduce(mine);
            if (lane == 0) red[warp] = mine;
            __syncthreads();
            if (tid == 0) {
                MinLoc tbest = red[0];
                for (int w = 1; w < SMALL_NT / 32; ++w)
                    if (minloc_better(red[w], tbest)) tbest = red[w];
                if (p.tile_records != nullptr) p.tile_records[tile] = tbest;
                if (minloc_better(tbest, best)) best = tbest;
            }
            __syncthreads();
        }
    }
    if (tid == 0 && p.partials != nullptr) p.partials[blockIdx.x] = best;
}

}  // namespace bopy
