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

// lower triangle of K + (noise + alpha) I, row-major with leading dimension ld >= n (the strict upper triangle is
// left untouched); rows n..n_fill-1 get a unit diagonal (identity padding for the blocked algorithms)
template <int KIND>
__global__ void gram_kernel(const double* __restrict__ X, int n, int d, LsParam ls, double amp, double diag_value,
                            double* __restrict__ A, int ld, int n_fill) {
    const int i = blockIdx.y * blockDim.y + threadIdx.y;   // row
    const int j = blockIdx.x * blockDim.x + threadIdx.x;   // column
    if (i >= n) {
        if (i < n_fill && j == i) A[(size_t)i * ld + j] = 1.0;
        return;
    }
    if (j > i) return;
    double v = diag_value;
    if (i != j) {
        double d2 = 0.0;
        for (int q = 0; q < d; ++q) {   // pdist(X / l): u = X[j] (the earlier row), v = X[i]
            const double df = __dadd_rn(__ddiv_rn(X[(size_t)j * d + q], ls.v[q]), -__ddiv_rn(X[(size_t)i * d + q], ls.v[q]));
            d2 = __dadd_rn(d2, __dmul_rn(df, df));
        }
        v = __dmul_rn(amp, base_kernel<KIND>(d2));
    }
    A[(size_t)i * ld + j] = v;
}

// In-place inverse of the lower-triangular block held in sm[BM][BM+1] (rdg[k] = 1 / L[k][k] precomputed), for a CTA of
// CHOL_NT = 256 threads; on return X[r][j] (r >= j) sits at sm[j][r].  T: scratch of 64 x 65 doubles.  Ends with a
// block-wide barrier.  Two levels: inv([[A, 0], [C, B]]) = [[inv A, 0], [-inv(B) C inv(A), inv B]] with 64 x 64 halves --
// the two diagonal inverses run side by side (4 warps each) and the off-diagonal block is two dense 64^3 products; the
// one-level column sweep had a dependent chain of 127 steps on warp 0 (41 % of chol_block_kernel).
constexpr int CHOL_NT = 256;
constexpr int CHOL_PANEL = 16;
constexpr size_t chol_smem_bytes() {
    return ((size_t)BM * (BM + 1) + BM + 2 + CHOL_PANEL * (CHOL_PANEL + 1) + 64 * 65) * sizeof(double);
}
__device__ __forceinline__ void tri_inverse_in_place(double* sm, const double* rdg, double* T) {
    constexpr int LD = BM + 1, H = BM / 2, TLD = H + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // diagonal halves: X = inv(L_hh), one column j per lane PAIR (lane, lane ^ 16) of a warp -- warp w owns columns
    // 16 w .. 16 w + 15, the two lanes of a pair take the even / odd k of  X[r][j] = -(sum_{k=j..r-1} L[r][k] X[k][j]) / L[r][r]
    // and join by one shuffle.  X[r][j] (r >= j) is parked at sm[j][r] (diagonal and strict upper triangle of row j, which
    // nobody else touches: L is only read strictly below its diagonal from here on).  The lanes of a warp walk r and k
    // in lock step, so the reads of L[r][k] are broadcasts and a warp-level barrier per r orders the pair's exchange.
    {
        const int j = 16 * warp + (lane & 15), half = lane >> 4;
        if (half == 0) sm[j * LD + j] = rdg[j];
        __syncwarp();
        const int jw = 16 * warp;   // no column of this warp has entries above row jw
        const int rend = warp < 4 ? H : BM;
        for (int r = jw + 1; r < rend; ++r) {
            double s0 = 0.0, s1 = 0.0;
            int k = jw + half;
            for (; k + 2 < r; k += 4) {
                const double x0 = k >= j ? sm[j * LD + k] : 0.0;
                const double x1 = k + 2 >= j ? sm[j * LD + k + 2] : 0.0;
                s0 = fma(sm[r * LD + k], x0, s0);
                s1 = fma(sm[r * LD + k + 2], x1, s1);
            }
            if (k < r) {
                const double x0 = k >= j ? sm[j * LD + k] : 0.0;
                s0 = fma(sm[r * LD + k], x0, s0);
            }
            double sacc = s0 + s1;
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 16);
            if (half == 0 && r > j) sm[j * LD + r] = -sacc * rdg[r];
            __syncwarp();
        }
    }
    __syncthreads();
    // off-diagonal block: thread (ty, tx) owns rows ty + 16 i and columns tx + 16 jj (i, jj < 4) of the 64 x 64 results
    const int ty = tid >> 4, tx = tid & 15;
    double acc[4][4];
    // T = C inv(A):  T[r][j] = sum_{k >= j} L[64 + r][k] X[k][j],  X[k][j] parked at sm[j][k]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
    for (int k = 0; k < H; ++k) {
        double lv[4], xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) lv[i] = sm[(H + ty + 16 * i) * LD + k];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) xv[jj] = k >= tx + 16 * jj ? sm[(tx + 16 * jj) * LD + k] : 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fma(lv[i], xv[jj], acc[i][jj]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) T[(ty + 16 * i) * TLD + tx + 16 * jj] = acc[i][jj];
    __syncthreads();
    // X21 = -inv(B) T:  X[64 + r][j] = -sum_{s <= r} Xb[r][s] T[s][j],  Xb[r][s] parked at sm[64 + s][64 + r]; the result goes
    // to sm[j][64 + r] (rows < 64, columns >= 64: the untouched upper-right quarter)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
    for (int sidx = 0; sidx < H; ++sidx) {
        double xv[4], tv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = sidx <= ty + 16 * i ? sm[(H + sidx) * LD + H + ty + 16 * i] : 0.0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) tv[jj] = T[sidx * TLD + tx + 16 * jj];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fma(xv[i], tv[jj], acc[i][jj]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) sm[(tx + 16 * jj) * LD + H + ty + 16 * i] = -acc[i][jj];
    __syncthreads();
}

// (a) of chol_block_kernel, one warp: Cholesky of the 16 x 16 block at (k0, k0) of sm in registers (lane r = row r; lanes
// 16..31 mirror lanes 0..15 and do not store), rdg[k0 + r] = 1 / L_rr, and W = inv(L16) to w16 (row-major, leading dimension
// 17).  A function of its own (not inlined): inside the big kernel the unroller gives up on the triangular loops and the
// row array goes to local memory -- 17 k cycles per panel instead of 4 k.
__device__ __noinline__ void chol16_factor_and_invert(double* sm, double* rdg, double* w16, int* fail, int k0, int lane) {
    constexpr int LD = BM + 1, PW = CHOL_PANEL, WLD = PW + 1;
    const int r = lane & (PW - 1);
    double a[PW];
#pragma unroll
    for (int c = 0; c < PW; ++c) a[c] = c <= r ? sm[(k0 + r) * LD + k0 + c] : 0.0;
    bool bad = false;
    double rinv_own = 1.0;                  // 1 / L[r][r] of this lane's own row
#pragma unroll
    for (int k = 0; k < PW; ++k) {
        const double akk = __shfl_sync(0xffffffffu, a[k], k);
        bad = bad || !(akk > 0.0);
        // L_kk = sqrt(a_kk) from the reciprocal square root + one correction step (<= 1 ulp), 1 / L_kk for free
        const double ri = rsqrt(akk);
        double lkk = akk * ri;
        lkk = fma(fma(-lkk, lkk, akk), 0.5 * ri, lkk);
        if (r == k) {
            a[k] = lkk;
            rinv_own = ri;
        } else if (r > k) {
            a[k] = a[k] * ri;
        }
#pragma unroll
        for (int c = k + 1; c < PW; ++c) {
            const double lck = __shfl_sync(0xffffffffu, a[k], c);   // L[c][k]
            if (r >= c) a[c] = fma(-a[k], lck, a[c]);
        }
    }
    if (bad) {
        if (lane == 0) *fail = 1;
        return;
    }
    if (lane < PW) {
#pragma unroll
        for (int c = 0; c < PW; ++c)
            if (c <= r) sm[(k0 + r) * LD + k0 + c] = a[c];
        rdg[k0 + r] = rinv_own;
    }
    // W = inv(L16), lane j = column j: W[i][j] = -(sum_{k=j..i-1} L[i][k] W[k][j]) / L[i][i], the entries L[i][k] by
    // shuffle from lane i's row
    const int j = r;
    double x[PW];
#pragma unroll
    for (int i = 0; i < PW; ++i) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) {
            const double lik = __shfl_sync(0xffffffffu, a[k], i);
            if (k & 1) s1 = fma(lik, x[k], s1);
            else s0 = fma(lik, x[k], s0);
        }
        const double ri = __shfl_sync(0xffffffffu, rinv_own, i);
        x[i] = i < j ? 0.0 : (i == j ? ri : -(s0 + s1) * ri);
    }
    if (lane < PW) {
#pragma unroll
        for (int i = 0; i < PW; ++i) w16[i * WLD + j] = x[i];
    }
}

// In-place Cholesky of the diagonal block J (identity padded beyond n) + its inverse, all in shared memory.
// One CTA of CHOL_NT threads.  status[0] is set to J+1 if a non-positive pivot is met (not positive definite).
// This kernel is the serial chain of the device fit (n/128 of them, nothing else can run beside a diagonal block's
// factorisation), so it is organised around latency.  Panel-blocked, 16 columns at a time, 3 block-wide barriers per panel:
//   (a) one warp factorises the 16 x 16 diagonal block of the panel in REGISTERS (lane r = row r, column entries by
//       shuffles) with a reciprocal square root per pivot -- sqrt + divide were 2 x ~400 cycles on each of the 16 dependent
//       steps, 40 % of the kernel (ncu source page) -- and then inverts it (W = inv(L16), lane j = column j);
//   (b) the rows below become a small product, X = A_panel W^T: 2 threads per row, 8 independent dot products each (as a
//       forward substitution it was one dependent chain of 136 FMAs per row);
//   (c) the trailing lower triangle takes the rank-16 update, only the 16 x 16 blocks on or below the diagonal of what is
//       left (the update used to cover the full 128 x 128 square for every panel).
// The inverse of the whole block is column-parallel and barrier-free: thread j forward-substitutes column j of
// inv(L_JJ), parking it in row j of the (unused) strict upper triangle of the working block.

// (c) for NBK remaining 16-row blocks: thread (ty, tx) owns the elements (t0 + ty + 16 i, t0 + tx + 16 j), j <= i < NBK
template <int NBK>
__device__ __forceinline__ void chol_trailing_update(double* sm, int k0) {
    constexpr int LD = BM + 1, PW = CHOL_PANEL;
    const int tid = threadIdx.x, t0 = k0 + PW, c0 = t0 + (tid & 15), r0 = t0 + (tid >> 4);
    double upd[NBK][NBK];
#pragma unroll
    for (int i = 0; i < NBK; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) upd[i][j] = 0.0;
#pragma unroll 4
    for (int k = 0; k < PW; ++k) {
        const int kk = k0 + k;
        double lck[NBK], lrk[NBK];
#pragma unroll
        for (int j = 0; j < NBK; ++j) lck[j] = sm[(c0 + 16 * j) * LD + kk];
#pragma unroll
        for (int i = 0; i < NBK; ++i) lrk[i] = sm[(r0 + 16 * i) * LD + kk];
#pragma unroll
        for (int i = 0; i < NBK; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) upd[i][j] = fma(lrk[i], lck[j], upd[i][j]);
    }
#pragma unroll
    for (int i = 0; i < NBK; ++i) {
        const int r = r0 + 16 * i;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const int c = c0 + 16 * j;
            if (c <= r) sm[r * LD + c] -= upd[i][j];
        }
    }
}

__global__ void __launch_bounds__(CHOL_NT) chol_block_kernel(double* A, int n, int ld, int J, double* Dinv, int* status,
                                                             long long* prof = nullptr) {   // prof: optional [40] clock stamps (BOPY_B200_CHOL_PROF)
    extern __shared__ double sm[];          // [BM][BM+1] working block, rdg[BM] (1 / L_kk), fail flag, w16[16][17], T[64][65]
    constexpr int LD = BM + 1, PW = CHOL_PANEL, WLD = PW + 1;
    double* const rdg = sm + BM * LD;
    int* const fail = reinterpret_cast<int*>(rdg + BM);
    double* const w16 = rdg + BM + 2;       // inverse of the panel's 16 x 16 diagonal block, row-major
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, base = J * BM;
    for (int e = tid; e < BM * BM; e += CHOL_NT) {
        const int r = e / BM, c = e - r * BM, gr = base + r, gc = base + c;
        double v = (r == c) ? 1.0 : 0.0;
        if (gr < n && gc < n && c <= r) v = A[(size_t)gr * ld + gc];
        sm[r * LD + c] = (c <= r) ? v : 0.0;
    }
    if (tid == 0) *fail = 0;
    __syncthreads();
#define BOPY_CHOL_STAMP(slot)                                      \
    do {                                                           \
        if (prof != nullptr && tid == 0) prof[slot] = clock64();   \
    } while (0)
    BOPY_CHOL_STAMP(0);
    for (int k0 = 0; k0 < BM; k0 += PW) {
        // (a) Cholesky of the PW x PW diagonal block in registers, then its inverse
        if (warp == 0) chol16_factor_and_invert(sm, rdg, w16, fail, k0, lane);
        __syncthreads();
        BOPY_CHOL_STAMP(1 + 3 * (k0 / PW));
        if (*fail) {
            if (tid == 0) status[0] = J + 1;
            return;
        }
        // (b) panel rows below the diagonal block: X = A_panel W^T, X[row][c] = sum_{k <= c} A[row][k] W[c][k];
        //     thread = (row, 8 of the 16 columns)
        {
            const int row = k0 + PW + (tid >> 1), ch = tid & 1;
            double x[8];
            if (row < BM) {
                double av[PW];
#pragma unroll
                for (int k = 0; k < PW; ++k) av[k] = sm[row * LD + k0 + k];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int c = 8 * ch + cc;
                    double sacc = 0.0;
#pragma unroll
                    for (int k = 0; k < PW; ++k) sacc = fma(av[k], (k <= c) ? w16[c * WLD + k] : 0.0, sacc);
                    x[cc] = sacc;
                }
            }
            __syncwarp();   // the two threads of a row sit in one warp: both have read the row before either overwrites it
            if (row < BM) {
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) sm[row * LD + k0 + 8 * ch + cc] = x[cc];
            }
        }
        __syncthreads();
        BOPY_CHOL_STAMP(2 + 3 * (k0 / PW));
        // (c) trailing lower triangle -= panel panel^T
        switch ((BM - k0 - PW) / 16) {
            case 7: chol_trailing_update<7>(sm, k0); break;
            case 6: chol_trailing_update<6>(sm, k0); break;
            case 5: chol_trailing_update<5>(sm, k0); break;
            case 4: chol_trailing_update<4>(sm, k0); break;
            case 3: chol_trailing_update<3>(sm, k0); break;
            case 2: chol_trailing_update<2>(sm, k0); break;
            case 1: chol_trailing_update<1>(sm, k0); break;
            default: break;
        }
        __syncthreads();
        BOPY_CHOL_STAMP(3 + 3 * (k0 / PW));
    }
    for (int e = tid; e < BM * BM; e += CHOL_NT) {
        const int r = e / BM, c = e - r * BM, gr = base + r, gc = base + c;
        if (gr < n && gc < n && c <= r) A[(size_t)gr * ld + gc] = sm[r * LD + c];
    }
    if (tid < BM) rdg[tid] = 1.0 / sm[tid * LD + tid];
    __syncthreads();
    BOPY_CHOL_STAMP(25);
    tri_inverse_in_place(sm, rdg, w16 + PW * WLD);
    BOPY_CHOL_STAMP(26);
    double* D = Dinv + (size_t)J * BM * BM;
    for (int e = tid; e < BM * BM; e += CHOL_NT) {
        const int r = e / BM, c = e - r * BM;
        D[e] = c <= r ? sm[c * LD + r] : 0.0;
    }
    BOPY_CHOL_STAMP(27);
#undef BOPY_CHOL_STAMP
}

// Dinv[I] = inv(L_II) for ONE diagonal block of a factor that is already there (the appended row changed block I only)
__global__ void __launch_bounds__(CHOL_NT) dinv_block_kernel(const double* __restrict__ L, int n, int ld, int I, double* Dinv) {
    extern __shared__ double sm[];          // [BM][BM+1], then rdg[BM], ... (chol_smem_bytes(): the layout of chol_block_kernel)
    constexpr int LD = BM + 1;
    double* const rdg = sm + BM * LD;
    const int tid = threadIdx.x, base = I * BM;
    for (int e = tid; e < BM * BM; e += CHOL_NT) {
        const int r = e / BM, c = e - r * BM;
        sm[r * LD + c] = c <= r ? L_at(L, n, ld, base + r, base + c) : 0.0;
    }
    __syncthreads();
    if (tid < BM) rdg[tid] = 1.0 / sm[tid * LD + tid];
    __syncthreads();
    tri_inverse_in_place(sm, rdg, rdg + BM + 2 + CHOL_PANEL * (CHOL_PANEL + 1));
    double* D = Dinv + (size_t)I * BM * BM;
    for (int e = tid; e < BM * BM; e += CHOL_NT) {
        const int r = e / BM, c = e - r * BM;
        D[e] = c <= r ? sm[c * LD + r] : 0.0;
    }
}

// Row n of the factor of the data set grown by one point: L[n][0..n-1] = v = L^-1 k(X, x_new) (the latency path's forward
// solve of x_new, column 0 of its [n_pad][8] workspace), L[n][n] = sqrt(k(x,x) + noise + alpha - v.v).  One CTA.
// status[0] = 1 if the new pivot is not positive.
__global__ void __launch_bounds__(256) append_row_kernel(const double* __restrict__ V, int n, int ld, double diag,
                                                        double* L, int* status) {
    __shared__ double red[256];
    double s = 0.0;
    for (int k = threadIdx.x; k < n; k += 256) {
        const double v = V[(size_t)k * 8];
        L[(size_t)n * ld + k] = v;
        s = fma(v, v, s);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double piv = diag - red[0];
        if (!(piv > 0.0)) status[0] = 1;
        L[(size_t)n * ld + n] = sqrt(piv);
    }
}

// C_tile = beta * C_tile + sign * A_tile * B_tile^T on 128 x 128 x 128 tiles of row-major matrices, fp64 DMMA.
//   mode 0 (panel):    tile I in (J, nb):   A = Amat[I][J] (input), B = Dinv_J, C = Amat[I][J] (in place), beta 0
//   mode 1 (trailing): tiles I >= K > J:    A = Amat[I][J], B = Amat[K][J], C = Amat[I][K], beta 1, sign -1
//   mode 2 / 3: the same restricted to K = J+1 / to K >= J+2 (look-ahead: see launch_cholesky)
// Rows / columns beyond n are treated as zero on load and skipped on store.
__global__ void __launch_bounds__(NT) gemm_nt_kernel(double* Amat, int n, int ld, int J, const double* Dinv, int mode) {
    __shared__ __align__(16) double As[2][DmmaPolicy::KC * BM];
    __shared__ __align__(16) double Bs[2][DmmaPolicy::KC * BN];
    int I, K;
    if (mode == 0) {
        I = J + 1 + blockIdx.x;
        K = J;
    } else if (mode == 2) {   // trailing update of block column J+1 only (what the next diagonal block / panel needs)
        I = J + 1 + blockIdx.x;
        K = J + 1;
        mode = 1;
    } else {
        // enumerate the lower-triangular tile set {(I, K): K0 <= K <= I < nb} with a linear index; K0 = J+1 (mode 1: the
        // whole trailing matrix) or J+2 (mode 3: everything but block column J+1)
        const int K0 = mode == 3 ? J + 2 : J + 1;
        const int idx = blockIdx.x;
        int rrow = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while ((rrow + 1) * (rrow + 2) / 2 <= idx) ++rrow;
        while (rrow * (rrow + 1) / 2 > idx) --rrow;
        I = K0 + rrow;
        K = K0 + (idx - rrow * (rrow + 1) / 2);
        mode = 1;
    }
    const int tid = threadIdx.x;
    const DmmaPolicy pol(tid);
    const double* Ablk = Amat + (size_t)I * BM * ld + (size_t)J * BM;            // A[r][k]
    const double* Bblk = mode == 0 ? Dinv + (size_t)J * BM * BM : Amat + (size_t)K * BM * ld + (size_t)J * BM;
    const int ldb = mode == 0 ? BM : ld;
    const int rowsA = n - I * BM, rowsB = mode == 0 ? BM : n - K * BM, colsK = n - J * BM;   // valid extents
    // global -> register staging: 128 rows x 8 k per chunk, 4 doubles per thread per operand
    const int lr = tid >> 1, lk = (tid & 1) * 4;
    double ra[4], rb[4];
    auto load_chunk = [&](int kc) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = kc * DmmaPolicy::KC + lk + e;
            ra[e] = (lr < rowsA && k < colsK) ? Ablk[(size_t)lr * ld + k] : 0.0;
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
    double* Cblk = Amat + (size_t)I * BM * ld + (size_t)K * BM;
    const int colsC = n - K * BM;
#pragma unroll
    for (int i = 0; i < DmmaPolicy::RI; ++i) {
        const int r = pol.row_of(i);
        if (r >= rowsA) continue;
#pragma unroll
        for (int j = 0; j < DmmaPolicy::CJ; ++j) {
            const int c = pol.cand_of(j);
            if (c >= colsC || c >= BM) continue;
            double* dst = &Cblk[(size_t)r * ld + c];
            if (mode == 0)
                *dst = acc[i][j];
            else if (I != K || c <= r)       // diagonal tiles: lower triangle only
                *dst = *dst - acc[i][j];
        }
    }
}

// alpha = L^-T (L^-1 y) with the inverted diagonal blocks; one CTA of 1024 threads walks the block rows.
// z is a scratch vector (n_pad).  Threads are arranged as 8 k-groups x 128 columns so that every matrix access is a
// contiguous 1 KB row segment; partial sums are combined through shared memory in a fixed order.
__global__ void __launch_bounds__(1024) solve_alpha_kernel(const double* __restrict__ L, int n, int ld, int nb,
                                                           const double* __restrict__ Dinv,
                                                           const double* __restrict__ y, double* z, double* alpha) {
    __shared__ double part[8][BM];
    __shared__ double rhs[BM];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int kg = tid >> 7, c = tid & (BM - 1);
    // forward: z_I = Dinv_I (y_I - sum_{J<I} L_IJ z_J); rows of L are contiguous in k: one warp per row
    for (int I = 0; I < nb; ++I) {
        for (int r = warp; r < BM; r += nwarp) {
            const int gr = I * BM + r;
            double s = 0.0;
            if (gr < n)
                for (int k = lane; k < I * BM; k += 32) s = fma(L[(size_t)gr * ld + k], z[k], s);
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
            if (lane == 0) rhs[r] = gr < n ? y[gr] - s : 0.0;
        }
        __syncthreads();
        for (int r = warp; r < BM; r += nwarp) {      // z_I[r] = Dinv_I[r][0..r] . rhs
            double s = 0.0;
            for (int k = lane; k <= r; k += 32) s = fma(Dinv[((size_t)I * BM + r) * BM + k], rhs[k], s);
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
            if (lane == 0) z[I * BM + r] = s;
        }
        __syncthreads();
    }
    // backward: a_I = Dinv_I^T (z_I - sum_{J>I} L_JI^T a_J); for a fixed row k of L the 128 columns of block I are
    // contiguous: thread (kg, c) sums rows k = kg (mod 8)
    for (int I = nb - 1; I >= 0; --I) {
        const int gc = I * BM + c;
        double s = 0.0;
        if (gc < n)
            for (int k = (I + 1) * BM + kg; k < n; k += 8) s = fma(L[(size_t)k * ld + gc], alpha[k], s);
        part[kg][c] = s;
        __syncthreads();
        if (tid < BM) {
            double t = 0.0;
            for (int g = 0; g < 8; ++g) t += part[g][tid];
            rhs[tid] = z[I * BM + tid] - t;
        }
        __syncthreads();
        s = 0.0;                                         // a_I[c] = sum_{k >= c} Dinv_I[k][c] rhs[k]
        for (int k = c + kg; k < BM; k += 8) s = fma(Dinv[((size_t)I * BM + k) * BM + c], rhs[k], s);
        part[kg][c] = s;
        __syncthreads();
        if (tid < BM) {
            double t = 0.0;
            for (int g = 0; g < 8; ++g) t += part[g][tid];
            if (I * BM + tid < n) alpha[I * BM + tid] = t;
        }
        __syncthreads();
    }
}

// =====================================================================================================
// Log marginal likelihood and its gradient w.r.t. the log hyper-parameters ($SK/_gpr.py:541-656), SURVEY 8(f) rank 4
// =====================================================================================================

// out[0] = -0.5 y.alpha - sum log L_ii - n/2 log(2 pi)
__global__ void __launch_bounds__(1024) lml_value_kernel(const double* __restrict__ L, int n, int ld,
                                                         const double* __restrict__ y, const double* __restrict__ alpha,
                                                         double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += -0.5 * y[i] * alpha[i] - log(L[(size_t)i * ld + i]);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        out[0] = t - 0.5 * n * 1.8378770664093453;   // log(2 pi)
    }
}

// W diagonal blocks = Dinv (the rest of W is zero-filled by the caller)
__global__ void copy_diag_blocks_kernel(const double* __restrict__ Dinv, double* W, int ld) {
    const int I = blockIdx.x;
    for (int e = threadIdx.x; e < BM * BM; e += blockDim.x) {
        const int r = e / BM, c = e - r * BM;
        W[((size_t)I * BM + r) * ld + (size_t)I * BM + c] = Dinv[(size_t)I * BM * BM + e];
    }
}

// One 128 x 128 output tile C(I,J) = scale * sum_{kb in [kb0,kb1)} Aop(I,kb) * Bop(kb,J), fp64 DMMA, on padded
// (multiple-of-128, zero/identity filled) matrices, so no bounds checks.  Jobs:
//   0  TRTRI off-diagonal sum   T_IJ = sum_{K=J}^{I-1} L_IK W_KJ          I = J + delta        -> W_IJ
//   1  TRTRI scale              W_IJ = -Dinv_I T_IJ                        (in place)
//   2  K^-1 = W^T W             Kinv_IJ = sum_{K=I}^{nb-1} W_KI^T W_KJ     I >= J               -> out
__global__ void __launch_bounds__(NT) tile_gemm_kernel(int job, int nb, int delta, const double* __restrict__ Lmat,
                                                       double* W, const double* __restrict__ Dinv, double* out,
                                                       int ld) {
    __shared__ __align__(16) double As[2][DmmaPolicy::KC * BM];
    __shared__ __align__(16) double Bs[2][DmmaPolicy::KC * BN];
    int I, J, kb0, kb1;
    if (job == 2) {
        const int idx = blockIdx.x;   // lower-triangular tile enumeration
        int rrow = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while ((rrow + 1) * (rrow + 2) / 2 <= idx) ++rrow;
        while (rrow * (rrow + 1) / 2 > idx) --rrow;
        I = rrow;
        J = idx - rrow * (rrow + 1) / 2;
        kb0 = I;
        kb1 = nb;
    } else {
        J = blockIdx.x;
        I = J + delta;
        kb0 = job == 0 ? J : 0;
        kb1 = job == 0 ? I : 1;
    }
    const int tid = threadIdx.x;
    const DmmaPolicy pol(tid);
    // operand element addressing: A(r,k), B(k,c) for k-block kb
    //   job 0: A = L[I*BM+r][kb*BM+k]  (k contiguous)      B = W[kb*BM+k][J*BM+c]   (c contiguous)
    //   job 1: A = Dinv[I][r][k]       (k contiguous)      B = W[I*BM+k][J*BM+c]    (c contiguous)
    //   job 2: A = W[kb*BM+k][I*BM+r]  (r contiguous)      B = W[kb*BM+k][J*BM+c]   (c contiguous)
    const int lr = tid >> 1, lk = (tid & 1) * 4;       // k-contiguous loads: row lr, 4 consecutive k
    const int rk = tid >> 5, rx = (tid & 31) * 4;      // row-contiguous loads: k row rk, 4 consecutive x
    double ra[4], rb[4];
    auto load_chunk = [&](int kb, int kc) {
        const int k0 = kc * DmmaPolicy::KC;
        if (job == 2) {
            const double* src = W + ((size_t)kb * BM + k0 + rk) * ld + (size_t)I * BM + rx;
#pragma unroll
            for (int e = 0; e < 4; ++e) ra[e] = src[e];
        } else {
            const double* src = job == 0 ? Lmat + ((size_t)I * BM + lr) * ld + (size_t)kb * BM + k0 + lk
                                         : Dinv + ((size_t)I * BM + lr) * BM + k0 + lk;
#pragma unroll
            for (int e = 0; e < 4; ++e) ra[e] = src[e];
        }
        const size_t brow = job == 1 ? (size_t)I * BM : (size_t)kb * BM;
        const double* bsrc = W + (brow + k0 + rk) * ld + (size_t)J * BM + rx;
#pragma unroll
        for (int e = 0; e < 4; ++e) rb[e] = bsrc[e];
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (job == 2)
                As[buf][DmmaPolicy::a_index(rk, rx + e)] = ra[e];
            else
                As[buf][DmmaPolicy::a_index(lk + e, lr)] = ra[e];
            Bs[buf][DmmaPolicy::b_index(rk, rx + e)] = rb[e];
        }
    };
    double acc[DmmaPolicy::RI][DmmaPolicy::CJ];
#pragma unroll
    for (int i = 0; i < DmmaPolicy::RI; ++i)
#pragma unroll
        for (int j = 0; j < DmmaPolicy::CJ; ++j) acc[i][j] = 0.0;
    constexpr int NCH = BM / DmmaPolicy::KC;
    const int total = (kb1 - kb0) * NCH;
    load_chunk(kb0, 0);
    store_chunk(0);
    __syncthreads();
    for (int it = 0; it < total; ++it) {
        const int nx = it + 1;
        if (nx < total) load_chunk(kb0 + nx / NCH, nx % NCH);
        pol.mma_tile<false>(acc, As[it & 1], Bs[it & 1], -1);
        if (nx < total) store_chunk(nx & 1);
        __syncthreads();
    }
    double* C = (job == 2 ? out : W) + (size_t)I * BM * ld + (size_t)J * BM;
    const double scale = job == 1 ? -1.0 : 1.0;
#pragma unroll
    for (int i = 0; i < DmmaPolicy::RI; ++i) {
        const int r = pol.row_of(i);
#pragma unroll
        for (int jv = 0; jv < DmmaPolicy::CJ / 2; ++jv) {
            const int c = pol.cand_of(2 * jv);
            *reinterpret_cast<double2*>(&C[(size_t)r * ld + c]) =
                make_double2(scale * acc[i][2 * jv], scale * acc[i][2 * jv + 1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same 128 x 128 fp64 DMMA output tile, organised for throughput: 16-wide k chunks travel global -> shared memory by
// cp.async (no register staging) through a 4-stage ring, one block-wide barrier per chunk.  tile_gemm_kernel above stages
// 8-wide chunks through registers one chunk ahead and reaches ~25 % of the DMMA rate (68 us per 128^3 product); this one
// is what the triangular inversion below and K^-1 = W^T W run on.
//
// Triangular inversion W = L^-1 as a recursion over 2 x 2 block partitions instead of one block diagonal at a time:
//   inv([[A, 0], [C, B]]) = [[inv A, 0], [-inv(B) C inv(A), inv B]]
// Level h = 1, 2, 4, ... pairs the already inverted h-block diagonal squares; per level TWO launches cover every pair:
//   phase 1   T_IJ = sum_{K=J}^{top end}   L_IK W_KJ       (W11 lower triangular)          -> scratch
//   phase 2   W_IJ = -sum_{K=bottom start}^{I} W_IK T_KJ   (W22 lower triangular)          -> W
// so the critical path is ~2 (n/128) tile products instead of (n/128)^2 / 2 and every level exposes all its tiles at once.
// The scratch is the strict upper block triangle of W itself: T_IJ is parked at block (J, I), which nothing else reads
// (probe_inv_kernel masks columns beyond the row, K^-1 = W^T W sums over lower blocks only).  Block counts that are not a
// power of two: a pair whose bottom half lies beyond the last block row simply has no tiles.
//   mode 3    Kinv_IJ = sum_{K=I}^{nb-1} W_KI^T W_KJ, I >= J (lower-triangular tile enumeration)  -> out
constexpr int TG_KCH = 16;        // k per chunk
constexpr int TG_STAGES = 4;
constexpr size_t TG_SMEM_BYTES = (size_t)TG_STAGES * TG_KCH * (BM + BN) * sizeof(double);   // 128 KB

__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(NT) tile_gemm_async_kernel(int mode, int nb, int h, const double* __restrict__ Lmat,
                                                             double* W, double* out, int ld) {
    extern __shared__ __align__(16) unsigned char tg_smem[];
    double* const As = reinterpret_cast<double*>(tg_smem);                 // [STAGES][KCH][BM] (swizzled, k-major)
    double* const Bs = As + (size_t)TG_STAGES * TG_KCH * BM;               // [STAGES][KCH][BN]
    int I, J, kb0, kb1;
    if (mode == 3) {
        const int idx = blockIdx.x;   // lower-triangular tile enumeration
        int rrow = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while ((rrow + 1) * (rrow + 2) / 2 <= idx) ++rrow;
        while (rrow * (rrow + 1) / 2 > idx) --rrow;
        I = rrow;
        J = idx - rrow * (rrow + 1) / 2;
        kb0 = I;
        kb1 = nb;
    } else {
        const int pair = blockIdx.x / (h * h), rem = blockIdx.x - pair * h * h;
        const int top = 2 * h * pair, bottom = top + h;
        I = bottom + rem / h;
        J = top + rem % h;
        if (I >= nb) return;
        kb0 = mode == 1 ? J : bottom;
        kb1 = mode == 1 ? bottom : I + 1;
    }
    const int tid = threadIdx.x;
    const DmmaPolicy pol(tid);
    // operand element addressing for k-block kb:
    //   mode 1: A(r,k) = L[I*BM+r][kb*BM+k]  (k contiguous)     B(k,c) = W[kb*BM+k][J*BM+c]
    //   mode 2: A(r,k) = W[I*BM+r][kb*BM+k]  (k contiguous)     B(k,c) = T(kb,J)[k][c], parked at W block (J, kb)
    //   mode 3: A(r,k) = W[kb*BM+k][I*BM+r]  (r contiguous)     B(k,c) = W[kb*BM+k][J*BM+c]
    auto issue = [&](int chunk, int stage) {
        const int kb = kb0 + chunk / (BM / TG_KCH), k0 = (chunk % (BM / TG_KCH)) * TG_KCH;
        double* const as = As + (size_t)stage * TG_KCH * BM;
        double* const bs = Bs + (size_t)stage * TG_KCH * BN;
        if (mode == 3) {
#pragma unroll
            for (int e = 0; e < TG_KCH * BM / 2 / NT; ++e) {
                const int idx = tid + NT * e, k = idx / (BM / 2), r = 2 * (idx % (BM / 2));
                cp_async_16(&as[DmmaPolicy::a_index(k, r)], W + ((size_t)kb * BM + k0 + k) * ld + (size_t)I * BM + r);
            }
        } else {
            const double* const a_src = (mode == 1 ? Lmat : W) + (size_t)I * BM * ld + (size_t)kb * BM + k0;
#pragma unroll
            for (int e = 0; e < TG_KCH * BM / NT; ++e) {
                const int idx = tid + NT * e, r = idx / TG_KCH, k = idx % TG_KCH;
                cp_async_8(&as[DmmaPolicy::a_index(k, r)], a_src + (size_t)r * ld + k);
            }
        }
        const double* const b_src = mode == 2 ? W + ((size_t)J * BM + k0) * ld + (size_t)kb * BM
                                              : W + ((size_t)kb * BM + k0) * ld + (size_t)J * BM;
#pragma unroll
        for (int e = 0; e < TG_KCH * BN / 2 / NT; ++e) {
            const int idx = tid + NT * e, k = idx / (BN / 2), c = 2 * (idx % (BN / 2));
            cp_async_16(&bs[DmmaPolicy::b_index(k, c)], b_src + (size_t)k * ld + c);
        }
    };
    double acc[DmmaPolicy::RI][DmmaPolicy::CJ];
#pragma unroll
    for (int i = 0; i < DmmaPolicy::RI; ++i)
#pragma unroll
        for (int j = 0; j < DmmaPolicy::CJ; ++j) acc[i][j] = 0.0;
    const int total = (kb1 - kb0) * (BM / TG_KCH);
#pragma unroll
    for (int s = 0; s < TG_STAGES - 1; ++s) {
        if (s < total) issue(s, s);
        cp_async_commit();
    }
    for (int it = 0; it < total; ++it) {
        cp_async_wait<TG_STAGES - 2>();
        __syncthreads();   // chunk `it` has landed for every thread; everyone is done with the stage chunk it-1 used
        const int nx = it + TG_STAGES - 1;
        if (nx < total) issue(nx, nx % TG_STAGES);
        cp_async_commit();
        const double* const as = As + (size_t)(it % TG_STAGES) * TG_KCH * BM;
        const double* const bs = Bs + (size_t)(it % TG_STAGES) * TG_KCH * BN;
#pragma unroll
        for (int sub = 0; sub < TG_KCH / DmmaPolicy::KC; ++sub)
            pol.mma_tile<false>(acc, as + sub * DmmaPolicy::KC * BM, bs + sub * DmmaPolicy::KC * BN, -1);
    }
    cp_async_wait<0>();
    // mode 1 parks T_IJ in the upper block triangle, at block (J, I)
    double* C = mode == 3 ? out + (size_t)I * BM * ld + (size_t)J * BM
                          : (mode == 1 ? W + (size_t)J * BM * ld + (size_t)I * BM : W + (size_t)I * BM * ld + (size_t)J * BM);
    const double scale = mode == 2 ? -1.0 : 1.0;
#pragma unroll
    for (int i = 0; i < DmmaPolicy::RI; ++i) {
        const int r = pol.row_of(i);
#pragma unroll
        for (int jv = 0; jv < DmmaPolicy::CJ / 2; ++jv) {
            const int c = pol.cand_of(2 * jv);
            *reinterpret_cast<double2*>(&C[(size_t)r * ld + c]) =
                make_double2(scale * acc[i][2 * jv], scale * acc[i][2 * jv + 1]);
        }
    }
}

// gemm_nt_kernel (the Cholesky panel / trailing update: C_tile = beta C_tile + sign A_tile B_tile^T, modes as above) on the
// cp.async pipeline of tile_gemm_async_kernel: both operands are k-contiguous rows of a row-major matrix, scattered into the
// k-major fragment layout by 8-byte copies; rows / k beyond n arrive as zeros (cp.async with a zero source size).
__device__ __forceinline__ void cp_async_8_zfill(void* smem, const void* gmem, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem),
                 "r"(valid ? 8 : 0)
                 : "memory");
}

__global__ void __launch_bounds__(NT) gemm_nt_async_kernel(double* Amat, int n, int ld, int J, const double* Dinv, int mode) {
    extern __shared__ __align__(16) unsigned char tg_smem[];
    double* const As = reinterpret_cast<double*>(tg_smem);
    double* const Bs = As + (size_t)TG_STAGES * TG_KCH * BM;
    int I, K;
    if (mode == 0) {
        I = J + 1 + blockIdx.x;
        K = J;
    } else if (mode == 2) {
        I = J + 1 + blockIdx.x;
        K = J + 1;
        mode = 1;
    } else {
        const int K0 = mode == 3 ? J + 2 : J + 1;
        const int idx = blockIdx.x;
        int rrow = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while ((rrow + 1) * (rrow + 2) / 2 <= idx) ++rrow;
        while (rrow * (rrow + 1) / 2 > idx) --rrow;
        I = K0 + rrow;
        K = K0 + (idx - rrow * (rrow + 1) / 2);
        mode = 1;
    }
    const int tid = threadIdx.x;
    const DmmaPolicy pol(tid);
    const double* Ablk = Amat + (size_t)I * BM * ld + (size_t)J * BM;            // A[r][k]
    const double* Bblk = mode == 0 ? Dinv + (size_t)J * BM * BM : Amat + (size_t)K * BM * ld + (size_t)J * BM;
    const int ldb = mode == 0 ? BM : ld;
    const int rowsA = n - I * BM, rowsB = mode == 0 ? BM : n - K * BM, colsK = n - J * BM;   // valid extents
    auto issue = [&](int chunk, int stage) {
        const int k0 = chunk * TG_KCH;
        double* const as = As + (size_t)stage * TG_KCH * BM;
        double* const bs = Bs + (size_t)stage * TG_KCH * BN;
#pragma unroll
        for (int e = 0; e < TG_KCH * BM / NT; ++e) {
            const int idx = tid + NT * e, r = idx / TG_KCH, k = idx % TG_KCH;
            const bool va = r < rowsA && k0 + k < colsK, vb = r < rowsB && (mode == 0 || k0 + k < colsK);
            cp_async_8_zfill(&as[DmmaPolicy::a_index(k, r)], va ? Ablk + (size_t)r * ld + k0 + k : Ablk, va);
            cp_async_8_zfill(&bs[DmmaPolicy::b_index(k, r)], vb ? Bblk + (size_t)r * ldb + k0 + k : Bblk, vb);
        }
    };
    double acc[DmmaPolicy::RI][DmmaPolicy::CJ];
#pragma unroll
    for (int i = 0; i < DmmaPolicy::RI; ++i)
#pragma unroll
        for (int j = 0; j < DmmaPolicy::CJ; ++j) acc[i][j] = 0.0;
    constexpr int total = BM / TG_KCH;
#pragma unroll
    for (int s = 0; s < TG_STAGES - 1; ++s) {
        if (s < total) issue(s, s);
        cp_async_commit();
    }
    for (int it = 0; it < total; ++it) {
        cp_async_wait<TG_STAGES - 2>();
        __syncthreads();
        const int nx = it + TG_STAGES - 1;
        if (nx < total) issue(nx, nx % TG_STAGES);
        cp_async_commit();
        const double* const as = As + (size_t)(it % TG_STAGES) * TG_KCH * BM;
        const double* const bs = Bs + (size_t)(it % TG_STAGES) * TG_KCH * BN;
#pragma unroll
        for (int sub = 0; sub < TG_KCH / DmmaPolicy::KC; ++sub)
            pol.mma_tile<false>(acc, as + sub * DmmaPolicy::KC * BM, bs + sub * DmmaPolicy::KC * BN, -1);
    }
    cp_async_wait<0>();
    double* Cblk = Amat + (size_t)I * BM * ld + (size_t)K * BM;
    const int colsC = n - K * BM;
#pragma unroll
    for (int i = 0; i < DmmaPolicy::RI; ++i) {
        const int r = pol.row_of(i);
        if (r >= rowsA) continue;
#pragma unroll
        for (int j = 0; j < DmmaPolicy::CJ; ++j) {
            const int c = pol.cand_of(j);
            if (c >= colsC || c >= BM) continue;
            double* dst = &Cblk[(size_t)r * ld + c];
            if (mode == 0)
                *dst = acc[i][j];
            else if (I != K || c <= r)       // diagonal tiles: lower triangle only
                *dst = *dst - acc[i][j];
        }
    }
}

constexpr int LML_NG = MAX_D + 2;   // gradient slots: [amplitude, length scales..., noise]

// partial[b][g] = sum over this block's (i >= k) pairs of w_ik * dK_ik/dlog(theta_g), w = (2 - [i==k]) (a_i a_k - Kinv_ik)
template <int KIND>
__global__ void __launch_bounds__(256) lml_grad_kernel(const double* __restrict__ X, int n, int d, int n_ls, LsParam ls,
                                                       double amp, double noise, const double* __restrict__ alpha,
                                                       const double* __restrict__ Kinv, int ld, double* partial) {
    const int i = blockIdx.y * 16 + threadIdx.y;
    const int k = blockIdx.x * 16 + threadIdx.x;
    double g[LML_NG];
#pragma unroll
    for (int q = 0; q < LML_NG; ++q) g[q] = 0.0;
    if (blockIdx.x <= blockIdx.y && i < n && k <= i) {
        const double w = (i == k ? 1.0 : 2.0) * (alpha[i] * alpha[k] - Kinv[(size_t)i * ld + k]);
        if (i == k) {
            g[0] = w * amp;                 // d(c * 1)/dlog c
            g[1 + n_ls] = w * noise;        // d(noise * I)/dlog noise
        } else {
            double D[MAX_D];
            double dsum = 0.0;
            for (int q = 0; q < d; ++q) {
                const double df = (X[(size_t)i * d + q] - X[(size_t)k * d + q]) / ls.v[q];
                D[q] = df * df;
                dsum += D[q];
            }
            const double kv = amp * base_kernel<KIND>(dsum);
            g[0] = w * kv;
            double common;                  // dK/dlog l_q = common * D_q
            if (KIND == K_RBF) common = kv;
            else if (KIND == K_M12) common = dsum > 0.0 ? kv / sqrt(dsum) : 0.0;
            else if (KIND == K_M32) common = amp * 3.0 * exp(-sqrt(3.0 * dsum));
            else {
                const double tmp = sqrt(5.0 * dsum);
                common = amp * (5.0 / 3.0) * (tmp + 1.0) * exp(-tmp);
            }
            if (n_ls == 1) g[1] = w * common * dsum;
            else
                for (int q = 0; q < d; ++q) g[1 + q] = w * common * D[q];
        }
    }
    // block reduction (fixed order), one partial row per block
    __shared__ double red[8][LML_NG];
    const int tid = threadIdx.y * 16 + threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int q = 0; q < n_ls + 2; ++q) {
        double v = g[q];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (lane == 0) red[warp][q] = v;
    }
    __syncthreads();
    if (tid < n_ls + 2) {
        double t = 0.0;
        for (int w2 = 0; w2 < 8; ++w2) t += red[w2][tid];
        partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * LML_NG + tid] = t;
    }
}

// grad[g] = 0.5 * sum_b partial[b][g] over the lower-triangular blocks, fixed order
__global__ void __launch_bounds__(256) lml_grad_finalize_kernel(const double* __restrict__ partial, int gx, int gy, int ng,
                                                                double* grad) {
    __shared__ double red[8];
    for (int q = 0; q < ng; ++q) {
        double s = 0.0;
        for (int b = threadIdx.x; b < gx * gy; b += blockDim.x) {
            const int by = b / gx, bx = b - by * gx;
            if (bx <= by) s += partial[(size_t)b * LML_NG + q];
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += red[w];
            grad[q] = 0.5 * t;
        }
        __syncthreads();
    }
}

}  // namespace bopy
