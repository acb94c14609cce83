// Branch-and-bound for the arg-min sweep: a cheap, rigorous LOWER bound of the acquisition per candidate, then only the
// candidates whose bound does not exceed an incumbent value go through the full fused sweep.
//
// The posterior variance obeys 0 < var <= prior = (amplitude + noise) y_std^2 (k(x,x) - sum v^2 with sum v^2 >= 0,
// $SK/_gpr.py:466), and all three acquisitions of bopy/acquisition.py:83-85, 99-106, 123-128 are monotone in the
// standard deviation on the side that matters:
//   LCB  mean - kappa sd                      >= mean - kappa sd_max                       (kappa >= 0)
//   EI   -E[(eta - f)^+], f ~ N(mean, sd^2)   >= the same at sd_max  (more spread, more expected improvement)
//   POI  P(f > eta)                           >= the same at sd_max if mean > eta, >= 0 otherwise
// and increasing in the mean, so evaluating them at (mean - slack, sd_max) bounds the true value from below.  The mean
// costs n (3d + 1 exp) flops per candidate, 1/40 of the full posterior at n = 2048.  The arg-min over the survivors
// {bound <= incumbent} equals the arg-min over all candidates (index and value: survivors keep their order and the
// sweep's arithmetic does not depend on position) -- EXCEPT that a candidate whose variance rounds to <= 0 (NaN
// acquisition, which np.argmin would return first) can be pruned; the pruned entry point is therefore opt-in.
#pragma once
#include "aux_kernels.cuh"
#include "common.cuh"
#include "sweep_kernel.cuh"

namespace bopy {

constexpr int PRUNE_NT = 256;

// mean[c] = y_std * sum_i alpha_i k(x_c, X_i) + y_mean, bound[c] = acquisition lower bound (see above)
template <int KIND, int DPAD>
__global__ void __launch_bounds__(PRUNE_NT) mean_bound_kernel(const double* __restrict__ Xt, const double* __restrict__ Xs,
                                                             long long m, int n_blocks, int d, LsParam ls, double amp,
                                                             double y_mean, double y_std, double sd_max, int acq,
                                                             double eta, double kappa, double* __restrict__ mean_out,
                                                             double* __restrict__ bound_out) {
    extern __shared__ __align__(16) double xrow[];   // [(d+1)][BM]: X/l block row, then alpha
    const long long c = (long long)blockIdx.x * PRUNE_NT + threadIdx.x;
    double xs[DPAD];
#pragma unroll
    for (int q = 0; q < DPAD; ++q) xs[q] = (q < d && c < m) ? __ddiv_rn(Xs[c * d + q], ls.v[q]) : 0.0;
    double acc0 = 0.0, acc1 = 0.0, mag = 0.0;   // mag = sum |alpha_i k_i|: scale of the summation-order rounding
    for (int I = 0; I < n_blocks; ++I) {
        __syncthreads();
        const double* const src = Xt + (long long)I * (d + 1) * BM;
        for (int e = threadIdx.x; e < (d + 1) * BM; e += PRUNE_NT) xrow[e] = src[e];
        __syncthreads();
        for (int r = 0; r < BM; r += 2) {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int q = 0; q < DPAD; ++q) {
                if (q < d) {
                    const double2 xr = *reinterpret_cast<const double2*>(&xrow[q * BM + r]);
                    const double f0 = xs[q] - xr.x, f1 = xs[q] - xr.y;
                    d0 = fma(f0, f0, d0);
                    d1 = fma(f1, f1, d1);
                }
            }
            const double2 al = *reinterpret_cast<const double2*>(&xrow[d * BM + r]);   // 0 beyond n
            const double t0 = base_kernel<KIND>(d0) * al.x, t1 = base_kernel<KIND>(d1) * al.y;
            acc0 += t0;
            acc1 += t1;
            mag += fabs(t0) + fabs(t1);
        }
    }
    if (c >= m) return;
    const double mean = fma(y_std, amp * (acc0 + acc1), y_mean);
    if (mean_out) mean_out[c] = mean;
    if (bound_out) {
        // the sweep adds the same terms alpha_i k_i in another order: its mean differs from this one by at most
        // ~n eps sum |terms| (worst case); bound from a mean lowered by that much, and a spread raised by an ulp or two
        const double n_eps = 4.0 * (double)(n_blocks * BM) * 2.220446049250313e-16;
        const double lo = mean - n_eps * (fabs(y_std) * amp * mag + fabs(mean));
        sd_max = sd_max * (1.0 + 1e-12);
        double b;
        if (acq == A_LCB) {
            b = kappa >= 0.0 ? lo - kappa * sd_max : lo;
        } else if (acq == A_EI) {
            b = acquisition_value(A_EI, lo, sd_max * sd_max, eta, kappa);
        } else {
            b = lo > eta ? acquisition_value(A_POI, lo, sd_max * sd_max, eta, kappa) : 0.0;
        }
        // one more guard band for the epilogue's own rounding
        bound_out[c] = b - 1e-12 * fabs(b);
    }
}

// ---- order-preserving stream compaction of {i : bound[i] <= thr} (NaN bounds survive) ------------------------------
constexpr int COMPACT_ITEMS = 8;   // candidates per thread

__device__ __forceinline__ bool prune_keep(double b, double thr) { return !(b > thr); }

__global__ void __launch_bounds__(PRUNE_NT) compact_count_kernel(const double* __restrict__ bound, long long m, const double* __restrict__ thr_dev,
                                                                long long seg_len, unsigned* __restrict__ block_counts) {
    __shared__ unsigned warp_sums[PRUNE_NT / 32];
    // thr_dev[i / seg_len]: the incumbent of candidate i's segment, still on the device (NaN: the segment survives whole)
    const long long base = ((long long)blockIdx.x * PRUNE_NT + threadIdx.x) * COMPACT_ITEMS;
    unsigned cnt = 0;
#pragma unroll
    for (int k = 0; k < COMPACT_ITEMS; ++k)
        if (base + k < m && prune_keep(bound[base + k], thr_dev[(base + k) / seg_len])) ++cnt;
#pragma unroll
    for (int mask = 16; mask > 0; mask >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, mask);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned s = 0;
        for (int w = 0; w < PRUNE_NT / 32; ++w) s += warp_sums[w];
        block_counts[blockIdx.x] = s;
    }
}

// exclusive scan of the block counts in place (one block; nblocks <= a few thousand), total -> *total_out
__global__ void __launch_bounds__(1024) compact_scan_kernel(unsigned* block_counts, int nblocks, long long* total_out) {
    __shared__ unsigned part[1024];
    const int per = (nblocks + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(nblocks, lo + per);
    unsigned s = 0;
    for (int i = lo; i < hi; ++i) s += block_counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {   // Hillis-Steele inclusive scan
        const unsigned v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = threadIdx.x == 0 ? 0u : part[threadIdx.x - 1];
    for (int i = lo; i < hi; ++i) {
        const unsigned cnt = block_counts[i];
        block_counts[i] = run;
        run += cnt;
    }
    if (threadIdx.x == 1023) *total_out = (long long)part[1023];
}

// idx_out[offset...] = ascending indices of the survivors of this block; rows_out (optional) = their candidate rows
__global__ void __launch_bounds__(PRUNE_NT) compact_scatter_kernel(const double* __restrict__ bound, long long m, const double* __restrict__ thr_dev,
                                                                  long long seg_len, const unsigned* __restrict__ block_offsets,
                                                                  const double* __restrict__ Xs, int d,
                                                                  long long* __restrict__ idx_out,
                                                                  double* __restrict__ rows_out) {
    __shared__ unsigned warp_offs[PRUNE_NT / 32];
    const long long base = ((long long)blockIdx.x * PRUNE_NT + threadIdx.x) * COMPACT_ITEMS;
    bool keep[COMPACT_ITEMS];
    unsigned cnt = 0;
#pragma unroll
    for (int k = 0; k < COMPACT_ITEMS; ++k) {
        keep[k] = base + k < m && prune_keep(bound[base + k], thr_dev[(base + k) / seg_len]);
        cnt += keep[k] ? 1u : 0u;
    }
    // exclusive scan of cnt over the block: within the warp by shuffles, across warps through shared memory
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) warp_offs[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int w = 0; w < PRUNE_NT / 32; ++w) {
            const unsigned v = warp_offs[w];
            warp_offs[w] = run;
            run += v;
        }
    }
    __syncthreads();
    long long pos = (long long)block_offsets[blockIdx.x] + warp_offs[warp] + (incl - cnt);
#pragma unroll
    for (int k = 0; k < COMPACT_ITEMS; ++k) {
        if (!keep[k]) continue;
        idx_out[pos] = base + k;
        if (rows_out)
            for (int q = 0; q < d; ++q) rows_out[pos * d + q] = Xs[(base + k) * d + q];
        ++pos;
    }
}

// out[s] = Xs[s * stride] (rows of d doubles): the strided sample that supplies the incumbent
__global__ void strided_rows_kernel(const double* __restrict__ Xs, int d, long long stride, long long S, double* out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S * d) return;
    const long long s = e / d;
    out[e] = Xs[s * stride * d + (e - s * d)];
}

// inc[s] = minimum (np.argmin ordering: a NaN wins) of the per_seg consecutive sample values of segment s
__global__ void segment_incumbent_kernel(const double* __restrict__ vals, long long nvals, long long per_seg, long long nseg,
                                         double* inc) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    MinLoc best;
    best.val = 0.0;
    best.idx = -1;
    for (long long j = s * per_seg; j < (s + 1) * per_seg && j < nvals; ++j) {
        MinLoc c;
        c.val = vals[j];
        c.idx = j;
        if (minloc_better(c, best)) best = c;
    }
    inc[s] = best.idx >= 0 ? best.val : __longlong_as_double(0x7ff8000000000000LL);
}

// per-segment arg-min over the survivors (idx_list ascending, vals[i] = acquisition of candidate idx_list[i]): one thread
// per segment finds its slice of the list by bisection and scans it with np.argmin's rules
__global__ void segment_argmin_survivors_kernel(const long long* __restrict__ idx_list, const double* __restrict__ vals,
                                                long long count, long long seg_len, long long nseg, long long index_base,
                                                double* val_out, long long* idx_out) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const long long first = s * seg_len, last = first + seg_len;
    long long lo = 0, hi = count;
    while (lo < hi) {   // first position with idx_list[pos] >= first
        const long long mid = (lo + hi) >> 1;
        if (idx_list[mid] < first) lo = mid + 1;
        else hi = mid;
    }
    MinLoc best;
    best.val = 0.0;
    best.idx = -1;
    for (long long pos = lo; pos < count && idx_list[pos] < last; ++pos) {
        MinLoc c;
        c.val = vals[pos];
        c.idx = index_base + idx_list[pos];
        if (minloc_better(c, best)) best = c;
    }
    val_out[s] = best.val;
    idx_out[s] = best.idx;
}

// min_idx (local index into the survivor list) -> original index
__global__ void remap_index_kernel(const long long* __restrict__ idx_list, long long count, long long index_base,
                                   long long* min_idx) {
    const long long local = *min_idx;
    *min_idx = (local >= 0 && local < count) ? index_base + idx_list[local] : -1;
}

}  // namespace bopy
