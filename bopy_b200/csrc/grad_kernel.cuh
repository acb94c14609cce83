// Gradient of the acquisition with respect to the candidate, for small candidate sets (multi-start refinement,
// the step right after the sweep behind Optimizer._optimize, bopy/optimizer.py:65-67).
//
//   d mean / dx = y_std    * sum_i alpha_i dk_i/dx                       ($SK/_gpr.py:447-450 differentiated)
//   d var  / dx = -2 y_var * sum_i w_i     dk_i/dx,   w = K^-1 k* = L^-T v   (v = L^-1 k*, $SK/_gpr.py:460-466)
//   dk_i/dx_q   = kd_i (x_q - X_iq) / l_q^2,  kd from the closed forms of $SK/kernels.py:1569-1570, 1720-1729
//   d acq / dx  = (d acq / d mean) d mean/dx + (d acq / d var) d var/dx    (bopy/acquisition.py:83-85, 99-106, 123-128)
//
// The forward solve v is left in global memory by probe_kernel (keep_v).  This kernel runs the BACKWARD substitution
// the same way probe_kernel runs the forward one, mirrored: CTA (I, g) owns block COLUMN I of L,
//   S_I = V_I - sum_{J>I} L_JI^T W_J     W_J arrives from CTA (J, g) (release/acquire flag), J = n_blocks-1 .. I+1
//   W_I = inv(L_II)^T S_I
// reading the very same packed tiles transposed: tile c of block (J, I) holds [k][row] = -L[J*128+row][I*128+8c+k], so
// used transposed it feeds exactly ONE 8-row output atom (rows 8c..8c+7 of W_I) with the 128 rows as the contraction
// axis -- warp w, which owns atoms {w, 15-w}, multiplies only tiles w and 15-w of every block (same DMMA count as the
// forward pass).  After publishing W_I the CTA forms its share of the two sums above over its 128 training rows
// (kd recomputed from X/l, a 128-term dot product per (candidate, dimension) in a fixed order) and publishes it
// under a second flag; the CTA of block column 0 -- the end of the chain -- adds the shares in block order, applies
// the acquisition partials and writes the gradient.
#pragma once
#include "probe_kernel.cuh"

namespace bopy {

struct GradParams {
    const unsigned char* Lt;
    const double* Xt;          // [n_blocks][d+1][BM]
    const double* V;           // [nbatch][NA][n_pad][8] forward solve, every block row (probe_kernel, keep_v)
    double* W;                 // same layout: backward solve
    const double* Xs;          // candidates (m, d)
    long long m;
    int nbatch, groups;
    int n, n_blocks, d;
    double ls[MAX_D];
    double amp, y_std, y_var;
    int acq;
    double eta, kappa;
    const double* mean;        // (m,) de-normalised posterior moments of the forward pass
    const double* var;
    double* grad_out;          // (m, d)
    unsigned* flags;           // [nbatch][n_blocks]: W_I published
    unsigned* flags2;          // [nbatch][n_blocks]: gradient shares of block I published
    double* gpart;             // [nbatch][n_blocks][2][d][NC]
    unsigned* ticket;
    unsigned ticket_base, epoch;
    int solve_only;            // 1: just the backward solve W = L^-T V, every block stored (alpha_ = L^-T L^-1 y); no gradient
};

template <int NA> constexpr size_t grad_smem_bytes(int d) {
    return (size_t)probe_stages<NA>() * TILE_BYTES +
           ((size_t)3 * NA * 1024 + (size_t)(d + 1) * BM + (size_t)d * 8 * NA) * sizeof(double) +
           (2 * probe_stages<NA>() + 1) * sizeof(uint64_t) + 16;
}

// kd with dk/dx_q = kd * (x_q - X_q) / l_q^2 for the base kernel evaluated from the squared scaled distance
template <int KIND> __device__ __forceinline__ double base_kernel_grad_factor(double d2) {
    if (KIND == K_RBF) return -exp_nonpos(-0.5 * d2);
    const double r = sqrt(d2);
    if (KIND == K_M12) return r > 0.0 ? -exp_nonpos(-r) / r : 0.0;
    if (KIND == K_M32) return -3.0 * exp_nonpos(-r * 1.7320508075688772);
    const double k = r * 2.23606797749979;
    return -(5.0 / 3.0) * (1.0 + k) * exp_nonpos(-k);
}

// (d acq / d mean, d acq / d var); NaN unless sqrt(var) > 0 (the reference's scale > 0 rule)
__device__ __forceinline__ void acquisition_partials(int acq, double mean, double var, double eta, double kappa,
                                                     double& dm, double& dv) {
    const double sd = sqrt(var);
    if (!(sd > 0.0)) {
        dm = dv = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    if (acq == A_LCB) {
        dm = 1.0;
        dv = -kappa / (2.0 * sd);
        return;
    }
    const double z = (eta - mean) / sd;
    const double pdf = exp(-0.5 * z * z) / 2.5066282746310002;
    if (acq == A_EI) {   // a = -sd pdf(z) - (eta - mean) cdf(z)
        dm = ndtr_like_scipy(z);
        dv = -pdf / (2.0 * sd);
    } else {             // POI: a = 1 - cdf(z)
        dm = pdf / sd;
        dv = pdf * z / (2.0 * var);
    }
}

template <int NA, int KIND>
__global__ void __launch_bounds__(PROBE_NT, 1) grad_kernel(const GradParams p) {
    constexpr int STG = probe_stages<NA>();
    constexpr int NC = 8 * NA;
    constexpr int HC = 4 * NA;
    using E = EngineF64;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* const ring = smem_raw;
    double* const Wb = reinterpret_cast<double*>(smem_raw + STG * TILE_BYTES);   // [2][NA][128][8] W_J; later U/T [2][128][NC]
    double* const Rs = Wb + 2 * NA * 1024;                                       // [NA][128][8] S_I, then W_I
    double* const xrow = Rs + NA * 1024;                                         // [(d+1)][128] X/l block + alpha
    double* const xs_s = xrow + (p.d + 1) * BM;                                  // [d][NC] candidates / l
    uint64_t* const full = reinterpret_cast<uint64_t*>(xs_s + p.d * NC);
    uint64_t* const empty = full + STG;
    uint64_t* const xbar = empty + STG;
    int* const role_s = reinterpret_cast<int*>(xbar + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = p.n_blocks, n_pad = nb * BM;

    if (tid == 0) {
        *role_s = (int)(atomicAdd(p.ticket, 1u) - p.ticket_base);
        for (int s = 0; s < STG; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NT / 32);
        }
        mbar_init(xbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int role = *role_s;
    const int I = nb - 1 - role / p.groups, g = role % p.groups;   // the chain starts at the LAST block column

    if (warp == NT / 32) {
        // producer: blocks (J, I), J = nb-1 .. I+1, then the inverted diagonal block; 16 tiles each
        if (lane != 0) return;
        uint32_t gc = 0;
        for (int b = g; b < p.nbatch; b += p.groups)
            for (int J = nb - 1; J >= I; --J) {
                const unsigned char* const src = p.Lt + (E::row_base(J) + (long long)I * E::CHG) * TILE_BYTES;
                for (int c = 0; c < E::CHG; ++c, ++gc) {
                    const uint32_t stage = gc % STG;
                    mbar_wait(&empty[stage], ((gc / STG) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&full[stage], TILE_BYTES);
                    bulk_g2s(ring + stage * TILE_BYTES, src + (long long)c * TILE_BYTES, TILE_BYTES, &full[stage]);
                }
            }
        return;
    }

    if (tid == 0) {
        const uint32_t bytes = (uint32_t)(p.d + 1) * BM * sizeof(double);
        mbar_arrive_expect_tx(xbar, bytes);
        bulk_g2s(xrow, p.Xt + (long long)I * (p.d + 1) * BM, bytes, xbar);
    }
    const int kq = lane & 3, q8 = lane >> 2;
    const int atoms[2] = {warp, 15 - warp};
    int roff[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) roff[i] = (8 * atoms[i] + q8) * 8 + 2 * kq;
    // transposed A fragment of k-step s (contraction rows 4s..4s+3): tile[k = q8][row = 4s + kq]
    const int abase = q8 * BM + kq, aswz = 4 * (q8 & 3);
    uint32_t gcount = 0;
    bool first = true;

    for (int b = g; b < p.nbatch; b += p.groups) {
        const long long c0 = (long long)b * NC;
        for (int e = tid; e < NC * p.d; e += NT) {
            const int c = e / p.d, q = e - c * p.d;
            const long long gcand = c0 + c;
            const double v = gcand < p.m ? p.Xs[gcand * p.d + q] : 0.0;
            xs_s[q * NC + c] = __ddiv_rn(v, p.ls[q]);
        }
        // accumulators [atom][candidate atom][chain]{c0, c1}, seeded with V_I (written by the forward kernel)
        double acc[2][NA][2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const double2 v = *reinterpret_cast<const double2*>(
                    &p.V[(((long long)b * NA + a) * n_pad + (long long)I * BM) * 8 + roff[i]]);
                acc[i][a][0][0] = v.x;
                acc[i][a][0][1] = v.y;
                acc[i][a][1][0] = acc[i][a][1][1] = 0.0;
            }
        consumer_sync();
        if (first) {
            mbar_wait(xbar, 0);
            first = false;
        }

        // ---- S_I = V_I - sum_{J>I} L_JI^T W_J ---------------------------------------------------------------------
        for (int J = nb - 1; J > I; --J) {
            if (lane == 0) {
                const unsigned* const f = p.flags + (long long)b * nb + J;
                long long spins = 0;
                while (ld_acquire_gpu(f) != p.epoch)
                    if (++spins > PROBE_SPIN_LIMIT) __trap();
            }
            __syncwarp();
            double* const wb = Wb + (J & 1) * NA * 1024;
            for (int e = tid; e < NA * 512; e += NT) {
                const int a = e >> 9, o = e & 511;
                const double* const src = p.W + (((long long)b * NA + a) * n_pad + (long long)J * BM) * 8;
                *reinterpret_cast<double2*>(&wb[a * 1024 + 2 * o]) = ld_cg_v2(src + 2 * o);
            }
            consumer_sync();
            for (int c = 0; c < E::CHG; ++c, ++gcount) {
                const uint32_t stage = gcount % STG;
                mbar_wait(&full[stage], (gcount / STG) & 1u);
                const double* const As = reinterpret_cast<const double*>(ring + stage * TILE_BYTES);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (atoms[i] != c) continue;   // warp-uniform: tile c feeds output atom c only
#pragma unroll 4
                    for (int s = 0; s < BM / 4; s += 2) {
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const double av = As[abase + ((4 * (s + u)) ^ aswz)];
#pragma unroll
                            for (int a = 0; a < NA; ++a)
                                dmma_m8n8k4(acc[i][a][u][0], acc[i][a][u][1], av, wb[a * 1024 + (4 * (s + u) + kq) * 8 + q8]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int a = 0; a < NA; ++a)
                *reinterpret_cast<double2*>(&Rs[a * 1024 + roff[i]]) =
                    make_double2(acc[i][a][0][0] + acc[i][a][1][0], acc[i][a][0][1] + acc[i][a][1][1]);
        consumer_sync();

        // ---- W_I = inv(L_II)^T S_I: upper triangular, output atom kc needs the rows >= 8 kc -----------------------
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int a = 0; a < NA; ++a)
                acc[i][a][0][0] = acc[i][a][0][1] = acc[i][a][1][0] = acc[i][a][1][1] = 0.0;
        for (int kc = 0; kc < E::CHD; ++kc, ++gcount) {
            const uint32_t stage = gcount % STG;
            mbar_wait(&full[stage], (gcount / STG) & 1u);
            const double* const As = reinterpret_cast<const double*>(ring + stage * TILE_BYTES);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (atoms[i] != kc) continue;
                for (int s = 2 * kc; s < BM / 4; s += 2) {   // pairs: the chain index stays a compile-time constant
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const double av = As[abase + ((4 * (s + u)) ^ aswz)];
#pragma unroll
                        for (int a = 0; a < NA; ++a)
                            dmma_m8n8k4(acc[i][a][u][0], acc[i][a][u][1], av, Rs[a * 1024 + (4 * (s + u) + kq) * 8 + q8]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
        }
        consumer_sync();   // every warp is done reading S_I: Rs now receives W_I

        // ---- publish W_I --------------------------------------------------------------------------------------------
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const double2 w = make_double2(acc[i][a][0][0] + acc[i][a][1][0], acc[i][a][0][1] + acc[i][a][1][1]);
                *reinterpret_cast<double2*>(&Rs[a * 1024 + roff[i]]) = w;
                if (I > 0 || p.solve_only)
                    *reinterpret_cast<double2*>(&p.W[(((long long)b * NA + a) * n_pad + (long long)I * BM) * 8 + roff[i]]) = w;
            }
        consumer_sync();
        if (tid == 0 && I > 0) {
            __threadfence();
            st_release_gpu(p.flags + (long long)b * nb + I, p.epoch);
        }

        if (p.solve_only) {
            consumer_sync();
            continue;
        }
        // ---- this block's share of sum_i alpha_i dk_i/dx and sum_i w_i dk_i/dx (off the chain's critical path) -----
        double* const UT = Wb;   // [2][128][NC]: t = alpha kd, u = w kd  (W_J buffers are free after the chain step)
        {
            const int row = tid & (BM - 1), h = tid >> 7;
            double d2[HC];
#pragma unroll
            for (int j = 0; j < HC; ++j) d2[j] = 0.0;
            for (int q = 0; q < p.d; ++q) {
                const double xr = xrow[q * BM + row];
#pragma unroll
                for (int j = 0; j < HC; ++j) {
                    const double df = xs_s[q * NC + h * HC + j] - xr;
                    d2[j] = fma(df, df, d2[j]);
                }
            }
            const double amp_i = (I * BM + row < p.n) ? p.amp : 0.0;
            const double a_i = xrow[p.d * BM + row];
#pragma unroll
            for (int j = 0; j < HC; ++j) {
                const int c = h * HC + j;
                const double kd = amp_i * base_kernel_grad_factor<KIND>(d2[j]);
                UT[row * NC + c] = a_i * kd;
                UT[BM * NC + row * NC + c] = Rs[(c >> 3) * 1024 + row * 8 + (c & 7)] * kd;
            }
        }
        consumer_sync();
        double* const gdst = p.gpart + ((long long)b * nb + I) * 2 * p.d * NC;
        for (int o = tid; o < 2 * p.d * NC; o += NT) {
            const int which = o / (p.d * NC), rem = o - which * p.d * NC;
            const int q = rem / NC, c = rem - q * NC;
            const double xc = xs_s[q * NC + c];
            const double* const u = UT + which * BM * NC + c;
            double s0 = 0.0, s1 = 0.0;   // two chains, fixed order
            for (int r = 0; r < BM; r += 2) {
                s0 = fma(u[r * NC], xc - xrow[q * BM + r], s0);
                s1 = fma(u[(r + 1) * NC], xc - xrow[q * BM + r + 1], s1);
            }
            gdst[o] = s0 + s1;
        }
        consumer_sync();
        if (I > 0) {
            if (tid == 0) {
                __threadfence();
                st_release_gpu(p.flags2 + (long long)b * nb + I, p.epoch);
            }
        } else {
            // ---- end of the chain: shares in block order, acquisition partials, gradient -----------------------------
            if (warp == 0) {
                for (int J = 1 + lane; J < nb; J += 32) {
                    const unsigned* const f = p.flags2 + (long long)b * nb + J;
                    long long spins = 0;
                    while (ld_acquire_gpu(f) != p.epoch)
                        if (++spins > PROBE_SPIN_LIMIT) __trap();
                }
            }
            consumer_sync();
            for (int o = tid; o < p.d * NC; o += NT) {
                const int q = o / NC, c = o - q * NC;
                const long long gcand = c0 + c;
                if (gcand >= p.m) continue;
                double gm = 0.0, gv = 0.0;
                for (int J = 0; J < nb; ++J) {
                    const double* const src = p.gpart + ((long long)b * nb + J) * 2 * p.d * NC;
                    gm += ld_cg(src + o);
                    gv += ld_cg(src + p.d * NC + o);
                }
                double dm, dv;
                acquisition_partials(p.acq, p.mean[gcand], p.var[gcand], p.eta, p.kappa, dm, dv);
                p.grad_out[gcand * p.d + q] = (dm * (p.y_std * gm) + dv * (-2.0 * p.y_var * gv)) / p.ls[q];
            }
        }
        consumer_sync();   // xs_s, Rs and the U/T buffers are reused by the next batch
    }
}

}  // namespace bopy
