// Fixed-hyper-parameter fit on the device: Gram matrix K(X,X) + alpha I, blocked Cholesky, alpha_ = K^-1 y.
//
// Reference arithmetic ($SK = sklearn/gaussian_process, scikit-learn 1.9.0):
//   K = kernel_(X_train_); K[diag] += alpha          $SK/_gpr.py:349-350  (pdist-based kernel, unit diagonal)
//   L_ = cholesky(K, lower=True)                     $SK/_gpr.py:352
//   alpha_ = cho_solve((L_, True), y_train_)         $SK/_gpr.py:363-367
// This is SURVEY.md section 8(f) rank 1 ("the step immediately before the path"): it removes the host fit and
// the upload of the n x n factor from every BO trial that keeps its hyper-parameters fixed.
//
// Right-looking blocked Cholesky with 128-wide block columns; per block column J:
//   chol_block_kernel   L_JJ = chol(A_JJ) in shared memory, and Dinv_J = inv(L_JJ) (needed by the sweep anyway)
//   gemm_nt_kernel      panel  L_IJ = A_IJ Dinv_J^T           (I > J)
//   gemm_nt_kernel      trailing A_IK -= L_IJ L_KJ^T          (I >= K > J), DMMA 128 x 128 tiles
#pragma once
#include "aux_kernels.cuh"
#include "common.cuh"
#include "sweep_kernel.cuh"

namespace bopy {

// lower triangle of K + (noise + alpha) I, row-major n x n (the strict upper triangle is left untouched)
template <int KIND>
__global__ void gram_kernel(const double* __restrict__ X, int n, int d, LsParam ls, double amp, double diag_value,
                            double* __restrict__ A) {
    const int i = blockIdx.y * blockDim.y + threadIdx.y;   // row
    const int j = blockIdx.x * blockDim.x + threadIdx.x;   // column
    if (i >= n || j > i) return;
    double v = diag_value;
    if (i != j) {
        double d2 = 0.0;
        for (int q = 0; q < d; ++q) {   // pdist(X / l): u = X[j] (the earlier row), v = X[i]
            const double df = __dadd_rn(__ddiv_rn(X[(size_t)j * d + q], ls.v[q]), -__ddiv_rn(X[(size_t)i * d + q], ls.v[q]));
            d2 = __dadd_rn(d2, __dmul_rn(df, df));
        }
        v = __dmul_rn(amp, base_kernel<KIND>(d2));
    }
    A[(size_t)i * n + j] = v;
}

// In-place Cholesky of the diagonal block J (identity padded beyond n) + its inverse.  One CTA, 128 threads.
// status[0] is set to J+1 if a non-positive pivot is met (matrix not positive definite).
__global__ void __launch_bounds__(BM) chol_block_kernel(double* A, int n, int J, double* Dinv, int* status) {
    extern __shared__ double sm[];          // [BM][BM+1]
    const int t = threadIdx.x, base = J * BM;
    constexpr int LD = BM + 1;
    for (int r = 0; r < BM; ++r) {          // thread t owns column t
        const int gr = base + r, gc = base + t;
        double v = (r == t) ? 1.0 : 0.0;
        if (gr < n && gc < n && t <= r) v = A[(size_t)gr * n + gc];
        sm[r * LD + t] = (t <= r) ? v : 0.0;
    }
    __syncthreads();
    for (int k = 0; k < BM; ++k) {
        // column k: pivot, scale
        const double akk = sm[k * LD + k];
        if (!(akk > 0.0)) {
            if (t == 0) status[0] = J + 1;
            return;
        }
        const double lkk = sqrt(akk);
        __syncthreads();
        if (t >= k) sm[t * LD + k] = (t == k) ? lkk : sm[t * LD + k] / lkk;   // thread t = row t of column k
        __syncthreads();
        // trailing update of the lower triangle: column t (> k), rows r >= t
        if (t > k) {
            const double ltk = sm[t * LD + k];
            for (int r = t; r < BM; ++r) sm[r * LD + t] = fma(-sm[r * LD + k], ltk, sm[r * LD + t]);
        }
        __syncthreads();
    }
    for (int r = 0; r < BM; ++r) {
        const int gr = base + r, gc = base + t;
        if (gr < n && gc < n && t <= r) A[(size_t)gr * n + gc] = sm[r * LD + t];
    }
    // Dinv_J = inv(L_JJ): forward substitution, column t per thread (same recurrence as dinv_kernel)
    double* D = Dinv + (size_t)J * BM * BM;
    for (int r = 0; r < BM; ++r) {
        double x = 0.0;
        if (r >= t) {
            double s = (r == t) ? 1.0 : 0.0;
            for (int k = t; k < r; ++k) s = fma(-sm[r * LD + k], D[(size_t)k * BM + t], s);
            x = s / sm[r * LD + r];
        }
        D[(size_t)r * BM + t] = x;
    }
}

// C_tile = beta * C_tile + sign * A_tile * B_tile^T on 128 x 128 x 128 tiles of row-major matrices, fp64 DMMA.
//   mode 0 (panel):    tile I in (J, nb):   A = Amat[I][J] (input), B = Dinv_J, C = Amat[I][J] (in place), beta 0
//   mode 1 (trailing): tiles I >= K > J:    A = Amat[I][J], B = Amat[K][J], C = Amat[I][K], beta 1, sign -1
// Rows / columns beyond n are treated as zero on load and skipped on store.
__global__ void __launch_bounds__(NT) gemm_nt_kernel(double* Amat, int n, int J, const double* Dinv, int mode) {
    __shared__ __align__(16) double As[2][DmmaPolicy::KC * BM];
    __shared__ __align__(16) double Bs[2][DmmaPolicy::KC * BN];
    int I, K;
    if (mode == 0) {
        I = J + 1 + blockIdx.x;
        K = J;
    } else {
        // enumerate the lower-triangular tile set {(I, K): J < K <= I < nb} with a linear index
        const int idx = blockIdx.x;
        int rrow = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while ((rrow + 1) * (rrow + 2) / 2 <= idx) ++rrow;
        while (rrow * (rrow + 1) / 2 > idx) --rrow;
        I = J + 1 + rrow;
        K = J + 1 + (idx - rrow * (rrow + 1) / 2);
    }
    const int tid = threadIdx.x;
    const DmmaPolicy pol(tid);
    const double* Ablk = Amat + (size_t)I * BM * n + (size_t)J * BM;             // A[r][k], ld n
    const double* Bblk = mode == 0 ? Dinv + (size_t)J * BM * BM : Amat + (size_t)K * BM * n + (size_t)J * BM;
    const int ldb = mode == 0 ? BM : n;
    const int rowsA = n - I * BM, rowsB = mode == 0 ? BM : n - K * BM, colsK = n - J * BM;   // valid extents
    // global -> register staging: 128 rows x 8 k per chunk, 4 doubles per thread per operand
    const int lr = tid >> 1, lk = (tid & 1) * 4;
    double ra[4], rb[4];
    auto load_chunk = [&](int kc) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = kc * DmmaPolicy::KC + lk + e;
            ra[e] = (lr < rowsA && k < colsK) ? Ablk[(size_t)lr * n + k] : 0.0;
            rb[e] = (lr < rowsB && (mode == 0 || k < colsK)) ? Bblk[(size_t)lr * ldb + k] : 0.0;
        }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            As[buf][DmmaPolicy::a_index(lk + e, lr)] = ra[e];
            Bs[buf][DmmaPolicy::b_index(lk + e, lr)] = rb[e];
        }
    };
    double acc[DmmaPolicy::RI][DmmaPolicy::CJ];
#pragma unroll
    for (int i = 0; i < DmmaPolicy::RI; ++i)
#pragma unroll
        for (int j = 0; j < DmmaPolicy::CJ; ++j) acc[i][j] = 0.0;
    constexpr int NCH = BM / DmmaPolicy::KC;
    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int kc = 0; kc < NCH; ++kc) {
        if (kc + 1 < NCH) load_chunk(kc + 1);
        pol.mma_tile<false>(acc, As[kc & 1], Bs[kc & 1], -1);
        if (kc + 1 < NCH) store_chunk((kc + 1) & 1);
        __syncthreads();
    }
    double* Cblk = Amat + (size_t)I * BM * n + (size_t)K * BM;
    const int colsC = n - K * BM;
#pragma unroll
    for (int i = 0; i < DmmaPolicy::RI; ++i) {
        const int r = pol.row_of(i);
        if (r >= rowsA) continue;
#pragma unroll
        for (int j = 0; j < DmmaPolicy::CJ; ++j) {
            const int c = pol.cand_of(j);
            if (c >= colsC || c >= BM) continue;
            double* dst = &Cblk[(size_t)r * n + c];
            if (mode == 0)
                *dst = acc[i][j];
            else if (I != K || c <= r)       // diagonal tiles: lower triangle only
                *dst = *dst - acc[i][j];
        }
    }
}

// alpha = L^-T (L^-1 y) with the inverted diagonal blocks; one CTA walks the block rows.  z is a scratch (n_pad).
__global__ void __launch_bounds__(1024) solve_alpha_kernel(const double* __restrict__ L, int n, int nb,
                                                           const double* __restrict__ Dinv,
                                                           const double* __restrict__ y, double* z, double* alpha) {
    __shared__ double rhs[BM];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    // forward: z_I = Dinv_I (y_I - sum_{J<I} L_IJ z_J)
    for (int I = 0; I < nb; ++I) {
        for (int r = warp; r < BM; r += nwarp) {
            const int gr = I * BM + r;
            double s = 0.0;
            if (gr < n)
                for (int k = lane; k < I * BM; k += 32) s = fma(L[(size_t)gr * n + k], z[k], s);
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
            if (lane == 0) rhs[r] = gr < n ? y[gr] - s : 0.0;
        }
        __syncthreads();
        if (tid < BM) {
            double s = 0.0;
            for (int k = 0; k <= tid; ++k) s = fma(Dinv[((size_t)I * BM + tid) * BM + k], rhs[k], s);
            z[I * BM + tid] = s;
        }
        __syncthreads();
    }
    // backward: a_I = Dinv_I^T (z_I - sum_{J>I} L_JI^T a_J)
    for (int I = nb - 1; I >= 0; --I) {
        if (tid < BM) {
            const int gc = I * BM + tid;
            double s = 0.0;
            if (gc < n)
                for (int k = (I + 1) * BM; k < n; ++k) s = fma(L[(size_t)k * n + gc], alpha[k], s);
            rhs[tid] = z[I * BM + tid] - s;
        }
        __syncthreads();
        if (tid < BM) {
            double s = 0.0;
            for (int k = tid; k < BM; ++k) s = fma(Dinv[((size_t)I * BM + k) * BM + tid], rhs[k], s);
            if (I * BM + tid < n) alpha[I * BM + tid] = s;
        }
        __syncthreads();
    }
}

}  // namespace bopy
