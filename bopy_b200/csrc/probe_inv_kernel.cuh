// Inverse path: the fused posterior -> acquisition -> arg-min for a HANDFUL of candidates (m <= 8), in ONE hop.
//
// The reference's only production caller probes the acquisition one point per call, up to 20 000 times per trial
// (DIRECT, bopy/optimizer.py:95-107); each probe is predict(return_cov=True) on a (1, d) array
// (bopy/surrogate.py:83-92 -> $SK/_gpr.py:446-473).  probe_kernel spreads the forward substitution of such a call
// over the block rows of L, which leaves a dependent chain of n/128 hops of ~4.4 us (77 us of device time at
// n = 2048).  A state that is probed thousands of times can afford to pay for W = L^-1 once (blocked TRTRI on the
// DMMA tile kernel tile_gemm_async_kernel: 0.64 ms at n = 2048, 7.8 ms at n = 8192): then
//
//   v = W k*        one matrix-vector product, no dependency between rows: rows are dealt to all warps of the grid
//   var = k(x,x) - sum_r v_r^2,   mean = k* . alpha
//
// is one launch whose length is a warp's share of the rows (n/64 coalesced 512-byte loads per warp, 16 MB out of L2 at
// n = 2048).  The explicit inverse is only ever multiplied from the left against k* in fp64; its forward error is the
// same cond(L) eps as the substitution's (measured on every golden set: <= 0.3 of the 1e-9 parity bound at
// cond(L) = 1.3e5, 3e-4 of it on the BASELINE configs: tools/inverse_path_study.py, profiles/r02/inverse_path_numerics.log).  K*, the de-normalisation and the
// acquisition are the code of the other two paths (base_kernel, acquisition_value); only the order in which the n
// products meet differs, so the paths agree to rounding, not bit for bit.  A candidate's value does not depend on m
// or on its position in the call: rows are dealt to warps by (row, n, grid) alone.
//
// Per CTA (512 threads): K* for ALL n rows (every CTA needs the whole column: n exps per candidate per CTA, <= 1 us),
// its rows of W, a fixed-order reduction of sum v^2 into part[CTA][candidate]; the CTA that draws the last ticket adds
// the partials in CTA order and runs the epilogue (threadfence reduction: no second launch, no cooperative launch).
#pragma once
#include "common.cuh"
#include "probe_kernel.cuh"

namespace bopy {

constexpr int INV_NT = 512;                 // 16 warps
constexpr int INV_WARPS = INV_NT / 32;
constexpr int INV_MAX_NC = 8;               // candidates per call
// 512-byte row segments in flight per warp: 8 for one or two candidates (64 KB per SM: what 6.5 TB/s x ~1 us of latency asks
// of each of 148 SMs when W streams from HBM, n = 8192), 4 beyond (register budget of 512-thread blocks)
template <int NC> __host__ __device__ constexpr int inv_unroll() { return NC <= 2 ? 8 : 4; }

struct InvParams {
    const double* W;           // [n_pad][n_pad] row-major L^-1 (lower triangular; rows >= n are never read)
    const double* Xt;          // [n_blocks][d+1][BM]: X / l dimension-major per block row, then alpha
    const double* Xs;          // candidates (m, d) row-major, or nullptr: they are in xs_inline
    int m;
    int n, n_blocks, d;
    double ls[MAX_D];
    double amp, kss, y_mean, y_std, y_var;
    int acq;
    double eta, kappa;
    double* mean_out;
    double* var_out;
    double* acq_out;
    long long index_base;
    int nan_skip;
    double* min_val;           // optional
    long long* min_idx;        // optional
    double* part;              // [gridDim.x][INV_MAX_NC] sum v^2 of a CTA's rows
    unsigned* ticket;          // monotonic over launches
    unsigned ticket_base;      // the CTA whose ticket is ticket_base + gridDim.x - 1 finishes the call
    // host-buffer entry (bopy_acq_eval_host): the candidates travel as kernel parameters and the outputs are mapped pinned
    // host memory, so that a DIRECT probe is one launch and one stream synchronisation with no memcpy call around it
    double xs_inline[INV_MAX_NC * MAX_D];
};

inline size_t inv_smem_bytes(int nc, int n_pad, int d) {
    return ((size_t)nc * n_pad + (size_t)d * nc + 2 * (size_t)INV_WARPS * nc + nc) * sizeof(double) + 16;
}

__device__ __forceinline__ double2 ld_nc_v2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <int NC, int KIND>
__global__ void __launch_bounds__(INV_NT, 1) probe_inv_kernel(const InvParams p) {
    constexpr int INV_UNROLL = inv_unroll<NC>();
    extern __shared__ __align__(16) unsigned char inv_smem[];
    const int n_pad = p.n_blocks * BM;
    double* const ks = reinterpret_cast<double*>(inv_smem);        // [NC][n_pad] K*
    double* const xs_s = ks + (size_t)NC * n_pad;                   // [d][NC] candidates / l
    double* const partS = xs_s + p.d * NC;                          // [warps][NC]
    double* const partM = partS + INV_WARPS * NC;                   // [warps][NC]
    double* const ss_s = partM + INV_WARPS * NC;                    // [NC]
    int* const last_s = reinterpret_cast<int*>(ss_s + NC);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total_warps = gridDim.x * INV_WARPS, gw = blockIdx.x * INV_WARPS + warp;

    // the head of this warp's first row of W does not depend on K*: its loads are in flight while K* is formed
    double2 w_head[INV_UNROLL];
    if (gw < p.n) {
        const double* const wr = p.W + (size_t)gw * n_pad;
#pragma unroll
        for (int u = 0; u < INV_UNROLL; ++u) {
            const int j = 2 * lane + 64 * u;
            w_head[u] = j < gw + 1 ? ld_nc_v2(wr + j) : make_double2(0.0, 0.0);
        }
    }

    for (int e = tid; e < NC * p.d; e += INV_NT) {
        const int c = e / p.d, q = e - c * p.d;
        const double v = c < p.m ? (p.Xs != nullptr ? p.Xs[(long long)c * p.d + q] : p.xs_inline[c * p.d + q]) : 0.0;
        xs_s[q * NC + c] = __ddiv_rn(v, p.ls[q]);
    }
    __syncthreads();

    // ---- K*[0..n) for every candidate of the call, and this thread's share of the mean --------------------------
    double mp[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) mp[c] = 0.0;
#pragma unroll 4
    for (int i = tid; i < n_pad; i += INV_NT) {
        const int I = i >> 7, row = i & (BM - 1);
        const double* const xrow = p.Xt + (long long)I * (p.d + 1) * BM;
        double d2[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) d2[c] = 0.0;
        for (int q = 0; q < p.d; ++q) {
            const double xr = xrow[q * BM + row];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const double df = xs_s[q * NC + c] - xr;
                d2[c] = fma(df, df, d2[c]);   // cdist's summation order over the dimensions
            }
        }
        const double amp_i = i < p.n ? p.amp : 0.0;
        const double a_i = xrow[p.d * BM + row];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double kv = __dmul_rn(amp_i, base_kernel<KIND>(d2[c]));
            ks[(size_t)c * n_pad + i] = kv;
            mp[c] = fma(kv, a_i, mp[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int mask = 16; mask > 0; mask >>= 1) mp[c] += __shfl_xor_sync(0xffffffffu, mp[c], mask);
        if (lane == 0) partM[warp * NC + c] = mp[c];
    }
    __syncthreads();

    // ---- v_r = W[r, 0..r] . K*[0..r] for this warp's rows, folded into sum v^2 -----------------------------------
    double ss[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) ss[c] = 0.0;
    // rows are dealt boustrophedon (0 .. T-1, 2T-1 .. T, 2T .. 3T-1, ...): row r costs r + 1 products, so every warp gets
    // about the same number of bytes (dealt round-robin the last warp would read 1.7x what the first does at n = 8192)
    for (int k = 0; k * total_warps < p.n; ++k) {
        const int r = (k & 1) ? (k + 1) * total_warps - 1 - gw : k * total_warps + gw;
        if (r >= p.n) continue;
        const double* const wr = p.W + (size_t)r * n_pad;
        const int len = r + 1;
        double acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = 0.0;
        for (int j0 = 2 * lane; j0 < len; j0 += 64 * INV_UNROLL) {
            double2 w[INV_UNROLL];
#pragma unroll
            for (int u = 0; u < INV_UNROLL; ++u) {
                const int j = j0 + 64 * u;
                if (k == 0 && j0 == 2 * lane) w[u] = w_head[u];
                else w[u] = j < len ? ld_nc_v2(wr + j) : make_double2(0.0, 0.0);
                if (j + 1 >= len) w[u].y = 0.0;   // the strict upper triangle is not part of L^-1
            }
#pragma unroll
            for (int u = 0; u < INV_UNROLL; ++u) {
                const int j = j0 + 64 * u;
                if (j < len) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const double2 kv = *reinterpret_cast<const double2*>(&ks[(size_t)c * n_pad + j]);
                        acc[c] = fma(w[u].x, kv.x, acc[c]);
                        acc[c] = fma(w[u].y, kv.y, acc[c]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            double v = acc[c];
#pragma unroll
            for (int mask = 16; mask > 0; mask >>= 1) v += __shfl_xor_sync(0xffffffffu, v, mask);
            ss[c] = fma(v, v, ss[c]);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < NC; ++c) partS[warp * NC + c] = ss[c];
    }
    __syncthreads();
    double mean_c = 0.0;
    if (tid < NC) {
        double s = 0.0;
        for (int w = 0; w < INV_WARPS; ++w) {
            s += partS[w * NC + tid];
            mean_c += partM[w * NC + tid];
        }
        p.part[(size_t)blockIdx.x * INV_MAX_NC + tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned t = atomicAdd(p.ticket, 1u) - p.ticket_base;
        *last_s = (t == gridDim.x - 1) ? 1 : 0;
        __threadfence();
    }
    __syncthreads();
    if (*last_s == 0) return;

    // ---- the last CTA: partials in CTA order, epilogue ------------------------------------------------------------
    if (warp < NC) {
        double s = 0.0;
        for (int g = lane; g < (int)gridDim.x; g += 32) s += ld_cg(p.part + (size_t)g * INV_MAX_NC + warp);
#pragma unroll
        for (int mask = 16; mask > 0; mask >>= 1) s += __shfl_xor_sync(0xffffffffu, s, mask);
        if (lane == 0) ss_s[warp] = s;
    }
    __syncthreads();
    if (warp == 0) {
        MinLoc mine;
        mine.val = 0.0;
        mine.idx = -1;
        if (tid < NC && tid < p.m) {
            const double mean = __dadd_rn(__dmul_rn(p.y_std, mean_c), p.y_mean);
            const double var = __dmul_rn(__dadd_rn(p.kss, -ss_s[tid]), p.y_var);
            if (p.mean_out) p.mean_out[tid] = mean;
            if (p.var_out) p.var_out[tid] = var;
            if (p.acq != A_NONE) {
                const double av = acquisition_value(p.acq, mean, var, p.eta, p.kappa);
                if (p.acq_out) p.acq_out[tid] = av;
                if (!(p.nan_skip && av != av)) {
                    mine.val = av;
                    mine.idx = p.index_base + tid;
                }
            }
        }
        if (p.min_val != nullptr || p.min_idx != nullptr) {
            mine = minloc_warp_reduce(mine);
            if (lane == 0) {
                if (p.min_val) *p.min_val = mine.val;
                if (p.min_idx) *p.min_idx = mine.idx;
            }
        }
    }
}

}  // namespace bopy
