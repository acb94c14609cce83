// One-time packing of the fitted state into the solve layout, the full-covariance view, the
// counter-based candidate generator and the FMA/MMA peak microbenchmarks.
#pragma once
#include "common.cuh"
#include "sweep_tc_kernel.cuh"

namespace bopy {

// L with identity padding beyond n (rows/cols >= n): the padded system solves to v = 0 there.
// (ld = leading dimension of the row-major matrix: n for a caller's factor, n_pad for the handle's own copy)
__device__ __forceinline__ double L_at(const double* __restrict__ L, int n, int ld, int r, int c) {
    return (r < n && c < n) ? L[(size_t)r * ld + c] : (r == c ? 1.0 : 0.0);
}

// Dinv[I] = inv(L_II), 128 x 128 lower triangular, fp64 forward substitution (column j per thread).
// Dinv doubles as the column store: thread j only re-reads what it wrote itself.
__global__ void dinv_kernel(const double* __restrict__ L, int n, int ld, double* Dinv) {
    const int I = blockIdx.x, j = threadIdx.x, base = I * BM;
    double* D = Dinv + (size_t)I * BM * BM;
    for (int r = 0; r < BM; ++r) {
        double x = 0.0;
        if (r >= j) {
            double s = (r == j) ? 1.0 : 0.0;
            for (int k = j; k < r; ++k) s = fma(-L_at(L, n, ld, base + r, base + k), D[(size_t)k * BM + j], s);
            x = s / L_at(L, n, ld, base + r, base + r);
        }
        D[(size_t)r * BM + j] = x;
    }
}

// Operand tile t of block row I: t < I*CHG is the off-diagonal tile [k][row] = -L[I*BM+row][t*KCG+k] in the GEMM
// policy's element type and physical layout; the CHD tiles after it hold Dinv[I][row][.] in the diagonal policy's.
template <class E>
__global__ void pack_tiles_kernel(const double* __restrict__ L, int n, int ld, const double* __restrict__ Dinv,
                                  unsigned char* out) {
    using PG = typename E::PG;
    using PD = typename E::PD;
    constexpr int KMAX = PG::KC > PD::KC ? PG::KC : PD::KC;
    const int I = blockIdx.y, t = blockIdx.x;
    if (t >= I * E::CHG + E::CHD) return;
    const bool offdiag = t < I * E::CHG;
    const int KC = offdiag ? PG::KC : PD::KC;
    __shared__ double tmp[KMAX][BM + 1];
    for (int e = threadIdx.x; e < BM * KC; e += blockDim.x) {
        const int r = e / KC, k = e - r * KC;
        double v;
        if (offdiag)
            v = -L_at(L, n, ld, I * BM + r, t * PG::KC + k);
        else
            v = Dinv[((size_t)I * BM + r) * BM + (t - I * E::CHG) * PD::KC + k];
        tmp[k][r] = v;
    }
    __syncthreads();
    unsigned char* dst = out + (E::row_base(I) + t) * (long long)TILE_BYTES;
    for (int e = threadIdx.x; e < BM * KC; e += blockDim.x) {
        const int k = e / BM, r = e - k * BM;
        if (offdiag) {
            if constexpr (std::is_same<PG, TcPolicy>::value) {
                // tensor-core operand tile: [hi 4 KB | lo 4 KB], K-major canonical layout, both halves rounded to nearest
                float hi, lo;
                tc::tf32_pair(tmp[k][r], hi, lo);
                reinterpret_cast<float*>(dst)[tc::tile_index(k, r)] = hi;
                reinterpret_cast<float*>(dst)[tc::TF32_TILE_FLOATS + tc::tile_index(k, r)] = lo;
            } else {
                reinterpret_cast<typename PG::Elem*>(dst)[PG::a_index(k, r)] = static_cast<typename PG::Elem>(tmp[k][r]);
            }
        } else
            reinterpret_cast<typename PD::Elem*>(dst)[PD::a_index(k, r)] = static_cast<typename PD::Elem>(tmp[k][r]);
    }
}

// Latency path: M_I = inv(L_II) L_{I,I-1}, stored NEGATED in the DMMA operand-tile layout ([k = 8][row = 128] per tile,
// 16 tiles per block row), so that  V_I = inv(L_II) P_I + (-M_I) V_{I-1}.  Block I = blockIdx.x + 1; thread (ty, tx) owns
// the 8 x 8 patch rows 8 ty.., columns 8 tx.. (= tile tx).  inv(L_II) is lower triangular: k runs to the row only.
__global__ void __launch_bounds__(256) pack_m_kernel(const double* __restrict__ L, int n, int ld, const double* __restrict__ Dinv,
                                                    unsigned char* out) {
    const int I = blockIdx.x + 1, ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const double* const D = Dinv + (size_t)I * BM * BM;
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
    const int kmax = 8 * ty + 8;   // columns of inv(L_II) that rows 8 ty .. 8 ty + 7 can reach
    for (int k = 0; k < kmax; ++k) {
        double dk[8], lk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dk[i] = D[(size_t)(8 * ty + i) * BM + k];          // 0 above the diagonal
#pragma unroll
        for (int j = 0; j < 8; ++j) lk[j] = L_at(L, n, ld, I * BM + k, (I - 1) * BM + 8 * tx + j);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fma(dk[i], lk[j], acc[i][j]);
    }
    double* const dst = reinterpret_cast<double*>(out + ((size_t)I * 16 + tx) * TILE_BYTES);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[DmmaPolicy::a_index(j, 8 * ty + i)] = -acc[i][j];
}

// out[i] = W[i][0]: column 0 of a latency-path workspace ([n_pad][8]) as a plain vector
__global__ void extract_column_kernel(const double* __restrict__ W, int n, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = W[(size_t)i * 8];
}

struct LsParam {
    double v[MAX_D];
};

// Xt[I][q][r] = X[I*BM+r][q] / l_q (q < d), Xt[I][d][r] = alpha[I*BM+r]; zero beyond n
__global__ void pack_x_kernel(const double* __restrict__ X, const double* __restrict__ alpha, int n, int d,
                              LsParam ls, double* Xt) {
    const int I = blockIdx.x, r = threadIdx.x, row = I * BM + r;
    double* dst = Xt + (size_t)I * (d + 1) * BM;
    for (int q = 0; q < d; ++q) dst[q * BM + r] = row < n ? __ddiv_rn(X[(size_t)row * d + q], ls.v[q]) : 0.0;
    dst[d * BM + r] = row < n ? alpha[row] : 0.0;
}

// cov[a][b] = (k(x_a, x_b) - sum_i V[i,a] V[i,b]) * y_std^2   ($SK/_gpr.py:466-469); V in the sweep's tile layout
template <class E, int KIND>
__global__ void cov_kernel(const typename E::TG* __restrict__ Vws, int n_pad, int n, const double* __restrict__ Xs,
                           long long m, int d, LsParam ls, double amp, double kss, double y_var, double* cov) {
    using T = typename E::TG;
    const long long a = (long long)blockIdx.y * blockDim.y + threadIdx.y;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= m || b >= m) return;
    const T* Va = Vws + (a / BN) * (long long)n_pad * BN * v_elems_per_entry<E>();
    const T* Vb = Vws + (b / BN) * (long long)n_pad * BN * v_elems_per_entry<E>();
    const int ca = (int)(a % BN), cb = (int)(b % BN);
    double s = 0.0;
    for (int i = 0; i < n; ++i) s = fma(v_value<E>(Va, i, ca), v_value<E>(Vb, i, cb), s);
    double prior = kss;
    if (a != b) {
        double d2 = 0.0;
        for (int q = 0; q < d; ++q) {
            const double df = __dadd_rn(__ddiv_rn(Xs[a * d + q], ls.v[q]), -__ddiv_rn(Xs[b * d + q], ls.v[q]));
            d2 = __dadd_rn(d2, __dmul_rn(df, df));
        }
        prior = __dmul_rn(amp, base_kernel<KIND>(d2));
    }
    cov[a * m + b] = __dmul_rn(__dadd_rn(prior, -s), y_var);
}

// the same with V as the latency path leaves it (probe_kernel with keep_v: [batch][NA][n_pad][8], 8 NA candidates per batch)
template <int KIND>
__global__ void cov_probe_kernel(const double* __restrict__ V, int na, int n_pad, int n, const double* __restrict__ Xs,
                                 long long m, int d, LsParam ls, double amp, double kss, double y_var, double* cov) {
    const long long a = (long long)blockIdx.y * blockDim.y + threadIdx.y;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= m || b >= m) return;
    const int nc = 8 * na;
    const double* Va = V + ((a / nc) * na + ((a % nc) >> 3)) * (long long)n_pad * 8 + (a & 7);
    const double* Vb = V + ((b / nc) * na + ((b % nc) >> 3)) * (long long)n_pad * 8 + (b & 7);
    double s = 0.0;
    for (int i = 0; i < n; ++i) s = fma(Va[(long long)i * 8], Vb[(long long)i * 8], s);
    double prior = kss;
    if (a != b) {
        double d2 = 0.0;
        for (int q = 0; q < d; ++q) {
            const double df = __dadd_rn(__ddiv_rn(Xs[a * d + q], ls.v[q]), -__ddiv_rn(Xs[b * d + q], ls.v[q]));
            d2 = __dadd_rn(d2, __dmul_rn(df, df));
        }
        prior = __dmul_rn(amp, base_kernel<KIND>(d2));
    }
    cov[a * m + b] = __dmul_rn(__dadd_rn(prior, -s), y_var);
}

// acquisition epilogue on given moments + block-wide arg-min; one block (small m)
__global__ void moments_kernel(int acq, double eta, double kappa, const double* __restrict__ mean,
                               const double* __restrict__ var, long long m, double* acq_out, long long index_base,
                               double* min_val, long long* min_idx) {
    __shared__ MinLoc red[32];
    MinLoc v;
    v.val = 0.0;
    v.idx = -1;
    for (long long i = threadIdx.x; i < m; i += blockDim.x) {
        MinLoc c;
        c.val = acquisition_value(acq, mean[i], var[i], eta, kappa);
        c.idx = index_base + i;
        if (acq_out) acq_out[i] = c.val;
        if (minloc_better(c, v)) v = c;
    }
    v = minloc_warp_reduce(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (minloc_better(red[w], v)) v = red[w];
        if (min_val) *min_val = v.val;
        if (min_idx) *min_idx = v.idx;
    }
}

// segment s = tile records [s*tiles_per_seg, (s+1)*tiles_per_seg): its arg-min under np.argmin's ordering
__global__ void segment_minloc_kernel(const MinLoc* __restrict__ tile_records, long long ntiles, int tiles_per_seg,
                                      long long nseg, double* val_out, long long* idx_out) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    MinLoc v;
    v.val = 0.0;
    v.idx = -1;
    for (int t = 0; t < tiles_per_seg; ++t) {
        const long long g = s * tiles_per_seg + t;
        if (g < ntiles && minloc_better(tile_records[g], v)) v = tile_records[g];
    }
    val_out[s] = v.val;
    idx_out[s] = v.idx;
}

struct BoxParam {
    double lo[MAX_D], hi[MAX_D];
};

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void candidates_kernel(unsigned long long seed, long long index_base, long long m, int d, BoxParam box,
                                  double* out) {
    const long long total = m * d;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / d;
        const int j = (int)(e - i * d);
        const unsigned long long ctr = (unsigned long long)(index_base + i) * (unsigned long long)d + j + 1ULL;
        const unsigned long long z = splitmix64(seed + 0x9E3779B97F4A7C15ULL * ctr);
        const double u = __dmul_rn((double)(z >> 11), 1.0 / 9007199254740992.0);
        out[e] = __dadd_rn(box.lo[j], __dmul_rn(u, __dadd_rn(box.hi[j], -box.lo[j])));
    }
}

// local candidate clouds for multi-start refinement: row s*P + p is start s itself for p == 0 (the incumbent), else
// start s + (2u - 1) * halfwidth, clipped to the box; u from the same counter-based generator
__global__ void candidates_around_kernel(unsigned long long seed, const double* __restrict__ starts, long long S,
                                         int P, int d, BoxParam half, BoxParam box, double* out) {
    const long long total = S * P * (long long)d;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / d;             // output row
        const int j = (int)(e - i * d);
        const long long s = i / P;
        const int pidx = (int)(i - s * P);
        const double c = starts[s * d + j];
        double x = c;
        if (pidx != 0) {
            const unsigned long long ctr = (unsigned long long)i * (unsigned long long)d + j + 1ULL;
            const unsigned long long z = splitmix64(seed + 0x9E3779B97F4A7C15ULL * ctr);
            const double u = __dmul_rn((double)(z >> 11), 1.0 / 9007199254740992.0);
            x = __dadd_rn(c, __dmul_rn(__dadd_rn(__dmul_rn(2.0, u), -1.0), half.lo[j]));
            x = fmin(fmax(x, box.lo[j]), box.hi[j]);
        }
        out[e] = x;
    }
}

// out[s][:] = xs[idx[s] - index_base][:]
__global__ void gather_rows_kernel(const double* __restrict__ xs, const long long* __restrict__ idx, long long S, int d,
                                   long long index_base, long long m, double* out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S * d) return;
    const long long s = e / d;
    const int j = (int)(e - s * d);
    const long long r = idx[s] - index_base;
    out[e] = (r >= 0 && r < m) ? xs[r * d + j] : __longlong_as_double(0x7ff8000000000000LL);
}

// ---- batched multi-start refinement: one projected-gradient step of every start ------------------------------------
// Monotone projected gradient with Barzilai-Borwein step lengths, all starts in lock step (one thread per start):
//   trial point xt = P(xc - alpha gc) has just been evaluated -> (ft, gt).
//   accept (ft <= fc, or fc is NaN and ft is not): s = xt - xc, y = gt - gc, alpha = s.s / s.y if s.y > 0 else 4 alpha
//           (clamped to [1e-12, 1e12]); (xc, fc, gc) <- (xt, ft, gt)
//   reject: alpha <- alpha / 4
//   next trial xt <- P(xc - alpha gc), P = clip to the box.  NaN values are never accepted over numbers.
// first != 0: (xt, ft, gt) is the evaluated START: it becomes (xc, fc, gc) and alpha = 0.1 min_q(hi-lo) / max_q |g_q|.
// Restated in oracle/gp_oracle.py (multistart_step) operation by operation.
__global__ void multistart_step_kernel(long long S, int d, BoxParam box, double* xc, double* fc, double* gc, double* xt,
                                       const double* __restrict__ ft, const double* __restrict__ gt, double* alpha,
                                       int first) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double* const xcs = xc + s * d;
    double* const gcs = gc + s * d;
    double* const xts = xt + s * d;
    const double* const gts = gt + s * d;
    double a = alpha[s];
    const double f_new = ft[s], f_cur = fc[s];
    bool accept;
    if (first) {
        accept = true;
        double gmax = 0.0, span = box.hi[0] - box.lo[0];
        for (int q = 0; q < d; ++q) {
            gmax = fmax(gmax, fabs(gts[q]));          // fmax ignores NaN: a NaN gradient leaves gmax unchanged
            span = fmin(span, box.hi[q] - box.lo[q]);
        }
        a = gmax > 0.0 ? 0.1 * span / gmax : 1.0;
    } else {
        const bool new_nan = f_new != f_new, cur_nan = f_cur != f_cur;
        accept = !new_nan && (cur_nan || f_new <= f_cur);
        if (accept) {
            double ss = 0.0, sy = 0.0;
            for (int q = 0; q < d; ++q) {
                const double sq = xts[q] - xcs[q], yq = gts[q] - gcs[q];
                ss = __dadd_rn(ss, __dmul_rn(sq, sq));   // unfused: the oracle restates this in plain fp64 arithmetic
                sy = __dadd_rn(sy, __dmul_rn(sq, yq));
            }
            a = sy > 0.0 ? ss / sy : 4.0 * a;
            a = fmin(fmax(a, 1e-12), 1e12);
        } else {
            a = 0.25 * a;
        }
    }
    if (accept) {
        fc[s] = f_new;
        for (int q = 0; q < d; ++q) {
            xcs[q] = xts[q];
            gcs[q] = gts[q];
        }
    }
    alpha[s] = a;
    for (int q = 0; q < d; ++q) {
        const double gq = gcs[q];
        const double step = gq == gq ? __dmul_rn(a, gq) : 0.0;  // NaN gradient (sigma <= 0): stay
        xts[q] = fmin(fmax(__dadd_rn(xcs[q], -step), box.lo[q]), box.hi[q]);
    }
}

// ---- register-resident peak microbenchmarks ---------------------------------------------------------
template <typename T> __global__ void peak_fma_kernel(T* out, int iters, T x, T y) {
    T a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = static_cast<T>(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = fma(a[k], x, y);
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    if (s == static_cast<T>(-1.2345)) out[0] = s;
}

// mma.sync m8n8k4 f64 (DMMA): 256 MACs per warp instruction; 32 independent accumulator pairs per warp,
// the same instruction-level parallelism the sweep kernel's 32 x 64 warp tile has
__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters, double x, double y) {
    double c[32][2];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        c[k][0] = threadIdx.x + k;
        c[k][1] = threadIdx.x - k;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1])
                         : "d"(x), "d"(y));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) s += c[k][0] + c[k][1];
    if (s == -1.2345) out[0] = s;
}

// legacy warp-level TF32 tensor instruction (mma.sync.m16n8k8, SASS HMMA.1688.F32.TF32): 1024 MACs per warp
// instruction; 16 independent accumulator quads per warp.  Measures what the non-tcgen05 tensor path sustains.
__global__ void __launch_bounds__(256) peak_tf32_mma_kernel(float* out, int iters, float x, float y) {
    float c[16][4];
#pragma unroll
    for (int k = 0; k < 16; ++k)
#pragma unroll
        for (int u = 0; u < 4; ++u) c[k][u] = threadIdx.x + k - u;
    const unsigned a0 = __float_as_uint(x), a1 = __float_as_uint(y), b0 = __float_as_uint(y), b1 = __float_as_uint(x);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            asm volatile(
                "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                : "+f"(c[k][0]), "+f"(c[k][1]), "+f"(c[k][2]), "+f"(c[k][3])
                : "r"(a0), "r"(a1), "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
    if (s == -1.2345f) out[0] = s;
}

__global__ void fill_kernel(double* dst, long long n, double v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

// tcgen05.mma kind::tf32 M=128 N=128 K=8 from shared-memory operands into two tensor-memory accumulators, one CTA per SM:
// the rate the fp32-mode engine's off-diagonal products can be issued at (its roofline denominator).
__global__ void __launch_bounds__(128, 1) peak_tcgen05_tf32_kernel(int iters) {
    extern __shared__ __align__(128) unsigned char psm[];
    float* const ops = reinterpret_cast<float*>(psm);          // 2 A tiles + 2 B tiles of 4 KB
    uint64_t* const bar = reinterpret_cast<uint64_t*>(psm + 4 * tc::TF32_TILE_BYTES);
    uint32_t* const slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4 * tc::TF32_TILE_FLOATS; i += blockDim.x) ops[i] = 0.0f;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(slot, 256);
    fence_proxy_async();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        constexpr uint32_t idesc = tc::idesc_tf32(BM, BN, false, false);
        const uint32_t base = smem_u32(ops);
        const uint64_t ah = tc::smem_desc(base, tc::TILE_LBO, tc::TILE_SBO), al = tc::smem_desc(base + tc::TF32_TILE_BYTES, tc::TILE_LBO, tc::TILE_SBO);
        const uint64_t bh = tc::smem_desc(base + 2 * tc::TF32_TILE_BYTES, tc::TILE_LBO, tc::TILE_SBO),
                       bl = tc::smem_desc(base + 3 * tc::TF32_TILE_BYTES, tc::TILE_LBO, tc::TILE_SBO);
        for (int it = 0; it < iters; ++it) {
            tc::mma_tf32(tmem + 128u, al, bh, idesc, it > 0);
            tc::mma_tf32(tmem + 128u, ah, bl, idesc, 1);
            tc::mma_tf32(tmem, ah, bh, idesc, it > 0);
        }
        tc::mma_commit(bar);
        mbar_wait(bar, 0);
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// ---- one-shot batch selection on the device (bopy/optimizer.py:186-232, 271-276): the k best evaluations that keep a
// minimum distance.  One pass per pick: candidates closer than min_distance to the newest pick die, the best survivor is
// the next pick (np.argmin ordering among the living, NaN values never picked).
__global__ void topk_pass_kernel(const double* __restrict__ x, const double* __restrict__ a, long long N, int d, LsParam scale,
                                 double min_dist2, const long long* __restrict__ picked, int npicked, unsigned char* alive,
                                 MinLoc* partials) {
    __shared__ MinLoc red[32];
    double newest[MAX_D];
    bool have = false;
    if (npicked > 0) {
        const long long pi = picked[npicked - 1];
        have = pi >= 0;
        if (have)
            for (int q = 0; q < d; ++q) newest[q] = x[pi * d + q];
    }
    MinLoc v;
    v.val = 0.0;
    v.idx = -1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        unsigned char al = npicked == 0 ? (unsigned char)1 : alive[i];
        const double ai = a[i];
        if (npicked == 0 && ai != ai) al = 0;
        if (al && have) {
            double d2 = 0.0;
            for (int q = 0; q < d; ++q) {
                const double df = (x[i * d + q] - newest[q]) / scale.v[q];
                d2 = fma(df, df, d2);
            }
            if (d2 < min_dist2 || i == picked[npicked - 1]) al = 0;
        }
        alive[i] = al;
        if (al) {
            MinLoc c;
            c.val = ai;
            c.idx = i;
            if (minloc_better(c, v)) v = c;
        }
    }
    v = minloc_warp_reduce(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (minloc_better(red[w], v)) v = red[w];
        partials[blockIdx.x] = v;
    }
}

}  // namespace bopy
