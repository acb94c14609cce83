// Latency path of the fused posterior -> acquisition -> arg-min for SMALL candidate sets.
//
// The reference's only production caller probes the acquisition one point per call (DIRECT,
// bopy/optimizer.py:95-107) and its plotting helpers / Kriging believer use a few hundred points
// (bopy/plotting.py:29-30, bopy/acquisition.py:189).  The throughput kernel (sweep_kernel) gives such a call to
// ONE CTA, which then walks all n^2/2 entries of L alone: 2.6 ms at n = 2048 whatever m <= 128 is.  Here the
// forward substitution of one small batch of candidates is spread over the block rows of L instead:
//
//   CTA (I, g) owns block row I of L for candidate group g.  For every batch b of its group (8*NA candidates):
//     P_I = K*_I - sum_{J<I-1} L_IJ V_J   V_J arrives from CTA (J, g) through global memory + a release/acquire flag
//     V_I = inv(L_II) P_I - M_I V_{I-1}   M_I = inv(L_II) L_{I,I-1} is precomputed, so that only ONE 128 x 128 x 8 product
//                                         separates the arrival of V_{I-1} from the publication of V_I (the diagonal
//                                         solve of P_I is done while waiting);  sum v^2 and K*_I.alpha go to a partials
//                                         array under a second flag, off the chain's critical path
//   the CTA of the last block row adds the partials in block-row order and runs the epilogue
//   (de-normalise, LCB / EI / POI, outputs, per-batch arg-min record).
//
// (A %globaltimer trace of the first version -- update by L_{I,I-1}, then the diagonal solve, then partials and V under
// one flag -- showed a 5.5 us hop = 0.4 flag + 2.3 load/update + 1.8 diagonal solve + 1.1 publish.)
// The chain V_0 -> V_1 -> ... is n/128 hops of a few microseconds; the 128 KB blocks of L a CTA needs stream from L2
// through a bulk-copy ring while it waits for the hop before.  Same arithmetic ($SK/_gpr.py:446-469), same packed
// operand tiles (swizzled [k][row], off-diagonal negated, diagonal blocks inverted), same DMMA fragments as
// sweep_kernel; only the order in which partial sums meet differs (results agree to rounding, not bit for bit).
//
// Roles are handed out by an atomic ticket, lowest block row first, so a CTA only ever waits for CTAs that started
// before it: no co-residency assumption, no cooperative launch.
#pragma once
#include "common.cuh"
#include "sweep_kernel.cuh"

namespace bopy {

constexpr int PROBE_NT = NT + 32;            // 8 compute warps + 1 producer warp
constexpr int PROBE_MAX_NC = 32;             // candidates per batch at NA = 4
constexpr long long PROBE_SPIN_LIMIT = 1LL << 24;   // ~10 s of polling: trap instead of hanging the GPU

struct ProbeParams {
    const unsigned char* Lt;   // packed factor tiles (EngineF64 layout)
    const unsigned char* Mt;   // [n_blocks][16] tiles of -inv(L_II) L_{I,I-1} (same layout): the last hop of a block row
    const double* Xt;          // [n_blocks][d+1][BM]
    double* V;                 // [nbatch][NA][n_pad][8]
    const double* Xs;          // candidates (m, d) row-major
    long long m;
    int nbatch, groups;
    int n, n_blocks, d;
    double ls[MAX_D];
    double amp, kss, y_mean, y_std, y_var;
    int acq;
    double eta, kappa;
    double* mean_out;
    double* var_out;
    double* acq_out;
    long long index_base;
    MinLoc* records;           // [nbatch] or nullptr
    unsigned* flags;           // [nbatch][n_blocks]: == epoch once V_I of the batch is published
    unsigned* flags2;          // [nbatch][n_blocks]: == epoch once the mean / sum v^2 partials of block row I are published
    double* part;              // [nbatch][n_blocks][2][NC]: mean / sum v^2 partials of block row I
    unsigned* ticket;          // role counter (monotonic over launches)
    unsigned ticket_base, epoch;
    int nan_skip;              // 1: NaN acquisition values never win the arg-min (np.nanargmin)
    int keep_v;                // 1: the last block row of V is stored too (the gradient's backward solve reads all of V)
    const double* rhs;         // optional (n,): solve L v = rhs for candidate 0 instead of a kernel column (alpha_ solve)
    long long* trace;          // optional [n_blocks][8] %globaltimer stamps of batch 0 (tools/probe_trace.py), or nullptr
};

__device__ __forceinline__ long long global_timer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define BOPY_TRACE(slot)                                                                          \
    do {                                                                                          \
        if (p.trace != nullptr && tid == 0 && b == g && g == 0) p.trace[I * 8 + (slot)] = global_timer_ns(); \
    } while (0)

template <int NA> __host__ __device__ constexpr int probe_stages() { return NA == 4 ? 8 : 16; }
template <int NA> constexpr size_t probe_smem_bytes(int d) {
    return (size_t)probe_stages<NA>() * TILE_BYTES +
           ((size_t)3 * NA * 1024 + (size_t)(d + 1) * BM + (size_t)d * 8 * NA + 8 * 4 * NA + 8 * 8 * NA) * sizeof(double) +
           (2 * probe_stages<NA>() + 1) * sizeof(uint64_t) + 16;
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// L2-coherent loads of data another SM published (never through this SM's L1)
__device__ __forceinline__ double2 ld_cg_v2(const double* p) {
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_cg(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

template <int NA, int KIND>
__global__ void __launch_bounds__(PROBE_NT, 1) probe_kernel(const ProbeParams p) {
    constexpr int STG = probe_stages<NA>();
    constexpr int NC = 8 * NA;      // candidates per batch
    constexpr int HC = 4 * NA;      // candidates per thread in the kernel-tile step
    using E = EngineF64;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* const ring = smem_raw;                                        // [STG] tiles of L_IJ / inv(L_II)
    double* const Vb = reinterpret_cast<double*>(smem_raw + STG * TILE_BYTES);   // [2][NA][128][8] V_J (double buffer)
    double* const Rs = Vb + 2 * NA * 1024;                                       // [NA][128][8] K*_I, then R_I
    double* const xrow = Rs + NA * 1024;                                         // [(d+1)][128] X/l block row + alpha
    double* const xs_s = xrow + (p.d + 1) * BM;                                  // [d][NC] candidates / l
    double* const partM = xs_s + p.d * NC;                                       // [8][HC]
    double* const partS = partM + 8 * HC;                                        // [8][NC]
    uint64_t* const full = reinterpret_cast<uint64_t*>(partS + 8 * NC);
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
    const int I = role / p.groups, g = role - I * p.groups;   // lowest block rows get the earliest tickets
    const int T_all = (I + 1) * E::CHG;                       // I off-diagonal blocks + the inverted diagonal block
    const unsigned char* const a_row = p.Lt + E::row_base(I) * TILE_BYTES;

    if (warp == NT / 32) {
        // =============================== bulk-copy producer (one lane) ====================================
        if (lane != 0) return;
        // tile order of a batch: blocks (I, 0 .. I-2), the inverted diagonal block, then M_I (stands for block (I, I-1))
        const unsigned char* const m_row = p.Mt + (long long)I * E::CHG * TILE_BYTES;
        const int T_early = (I > 0 ? I - 1 : 0) * E::CHG;
        uint32_t gc = 0;
        for (int b = g; b < p.nbatch; b += p.groups)
            for (int t = 0; t < T_all; ++t, ++gc) {
                const uint32_t stage = gc % STG;
                mbar_wait(&empty[stage], ((gc / STG) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&full[stage], TILE_BYTES);
                const unsigned char* src;
                if (t < T_early) src = a_row + (long long)t * TILE_BYTES;
                else if (t < T_early + E::CHD) src = a_row + (long long)(I * E::CHG + (t - T_early)) * TILE_BYTES;
                else src = m_row + (long long)(t - T_early - E::CHD) * TILE_BYTES;
                bulk_g2s(ring + stage * TILE_BYTES, src, TILE_BYTES, &full[stage]);
            }
        return;
    }

    // ========================================= compute warps ===============================================
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)(p.d + 1) * BM * sizeof(double);
        mbar_arrive_expect_tx(xbar, bytes);
        bulk_g2s(xrow, p.Xt + (long long)I * (p.d + 1) * BM, bytes, xbar);
    }
    const int kq = lane & 3, q8 = lane >> 2;
    // row atoms {w, 15-w}: in the triangular diagonal GEMM atom a needs a+1 k-tiles, so every warp does 17
    const int atoms[2] = {warp, 15 - warp};
    int aoff[2], roff[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        aoff[i] = DmmaPolicy::a_index(kq, 8 * atoms[i] + q8);
        roff[i] = (8 * atoms[i] + q8) * 8 + 2 * kq;   // accumulator pair inside a [128][8] block
    }
    const bool last = (I == nb - 1);
    uint32_t gcount = 0;
    bool first = true;

    for (int b = g; b < p.nbatch; b += p.groups) {
        const long long c0 = (long long)b * NC;
        BOPY_TRACE(0);
        for (int e = tid; e < NC * p.d; e += NT) {
            const int c = e / p.d, q = e - c * p.d;
            const long long gcand = c0 + c;
            const double v = gcand < p.m ? p.Xs[gcand * p.d + q] : 0.0;
            xs_s[q * NC + c] = __ddiv_rn(v, p.ls[q]);
        }
        consumer_sync();
        if (first) {
            mbar_wait(xbar, 0);
            first = false;
        }

        // ---- kernel tile K*[block row I, batch] and its share of the mean: thread = (row, half of the candidates) ----
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
                    d2[j] = fma(df, df, d2[j]);   // cdist's summation order over the dimensions
                }
            }
            const double amp_i = (I * BM + row < p.n) ? p.amp : 0.0;   // identity padding beyond n
            const double a_i = xrow[p.d * BM + row];
            double mp[HC];
#pragma unroll
            for (int j = 0; j < HC; ++j) {
                const int c = h * HC + j;
                double kv = __dmul_rn(amp_i, base_kernel<KIND>(d2[j]));
                if (p.rhs != nullptr)   // right-hand-side mode: column 0 of the batch is the given vector, the rest 0
                    kv = (c == 0 && I * BM + row < p.n) ? p.rhs[I * BM + row] : 0.0;
                Rs[(c >> 3) * 1024 + row * 8 + (c & 7)] = kv;
                mp[j] = kv * a_i;
            }
#pragma unroll
            for (int j = 0; j < HC; ++j) {
#pragma unroll
                for (int mask = 16; mask > 0; mask >>= 1) mp[j] += __shfl_xor_sync(0xffffffffu, mp[j], mask);
            }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < HC; ++j) partM[warp * HC + j] = mp[j];
            }
        }
        consumer_sync();
        double mean_part = 0.0, ss_part = 0.0;   // threads 0..NC-1: this block row's share for candidate tid
        if (tid < NC) {
            const int h = tid / HC, j = tid - h * HC;
            mean_part = ((partM[(4 * h) * HC + j] + partM[(4 * h + 1) * HC + j]) + partM[(4 * h + 2) * HC + j]) +
                        partM[(4 * h + 3) * HC + j];
        }
        // accumulators [atom][candidate atom][k-step chain]{c0, c1}, seeded with K*
        double acc[2][NA][2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const double2 v = *reinterpret_cast<const double2*>(&Rs[a * 1024 + roff[i]]);
                acc[i][a][0][0] = v.x;
                acc[i][a][0][1] = v.y;
                acc[i][a][1][0] = acc[i][a][1][1] = 0.0;
            }
        consumer_sync();   // K* consumed: Rs is free for the residual
        BOPY_TRACE(1);

        // one block of the solve: wait for V_J, stage it, acc += (16 tiles from the ring) . V_J
        auto block_update = [&](int J) {
            if (lane == 0) {
                const unsigned* const f = p.flags + (long long)b * nb + J;
                long long spins = 0;
                while (ld_acquire_gpu(f) != p.epoch)
                    if (++spins > PROBE_SPIN_LIMIT) __trap();
            }
            __syncwarp();
            if (J == I - 1) BOPY_TRACE(2);
            double* const vb = Vb + (J & 1) * NA * 1024;
            for (int e = tid; e < NA * 512; e += NT) {
                const int a = e >> 9, o = e & 511;
                const double* const src = p.V + (((long long)b * NA + a) * n_pad + (long long)J * BM) * 8;
                *reinterpret_cast<double2*>(&vb[a * 1024 + 2 * o]) = ld_cg_v2(src + 2 * o);
            }
            consumer_sync();
            for (int c = 0; c < E::CHG; ++c, ++gcount) {
                const uint32_t stage = gcount % STG;
                mbar_wait(&full[stage], (gcount / STG) & 1u);
                const double* const As = reinterpret_cast<const double*>(ring + stage * TILE_BYTES);
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const double a0 = As[aoff[0] + s * 4 * BM], a1 = As[aoff[1] + s * 4 * BM];
#pragma unroll
                    for (int a = 0; a < NA; ++a) {
                        const double bv = vb[a * 1024 + (c * 8 + 4 * s + kq) * 8 + q8];
                        dmma_m8n8k4(acc[0][a][s][0], acc[0][a][s][1], a0, bv);
                        dmma_m8n8k4(acc[1][a][s][0], acc[1][a][s][1], a1, bv);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
            }
        };

        // ---- P_I = K*_I - sum_{J < I-1} L_IJ V_J: everything that does not need the previous hop ------------------
        for (int J = 0; J + 1 < I; ++J) block_update(J);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int a = 0; a < NA; ++a)
                *reinterpret_cast<double2*>(&Rs[a * 1024 + roff[i]]) =
                    make_double2(acc[i][a][0][0] + acc[i][a][1][0], acc[i][a][0][1] + acc[i][a][1][1]);
        consumer_sync();
        BOPY_TRACE(3);

        // ---- inv(L_II) P_I (lower triangular: row atom a needs the k tiles kc <= a) --------------------------------
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
            for (int s = 0; s < 2; ++s) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (atoms[i] < kc) continue;   // warp-uniform
                    const double av = As[aoff[i] + s * 4 * BM];
#pragma unroll
                    for (int a = 0; a < NA; ++a)
                        dmma_m8n8k4(acc[i][a][s][0], acc[i][a][s][1], av, Rs[a * 1024 + (kc * 8 + 4 * s + kq) * 8 + q8]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
        }
        BOPY_TRACE(6);

        // ---- the hop: V_I = inv(L_II) P_I - M_I V_{I-1}, the M_I tiles already sit in the ring ----------------------
        if (I > 0) block_update(I - 1);

        BOPY_TRACE(4);
        // ---- publish V_I, fold into sum v^2 -----------------------------------------------------------------------
        {
            double sq[NA][2];
#pragma unroll
            for (int a = 0; a < NA; ++a) sq[a][0] = sq[a][1] = 0.0;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int a = 0; a < NA; ++a) {
                    const double v0 = acc[i][a][0][0] + acc[i][a][1][0], v1 = acc[i][a][0][1] + acc[i][a][1][1];
                    if (!last || p.keep_v)
                        *reinterpret_cast<double2*>(
                            &p.V[(((long long)b * NA + a) * n_pad + (long long)I * BM) * 8 + roff[i]]) = make_double2(v0, v1);
                    sq[a][0] = fma(v0, v0, sq[a][0]);
                    sq[a][1] = fma(v1, v1, sq[a][1]);
                }
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    double v = sq[a][u];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (q8 == 0) partS[warp * NC + a * 8 + 2 * kq + u] = v;
                }
        }
        consumer_sync();   // every V_I store of this CTA is ordered before the release below
        if (!last && tid == 0) {
            __threadfence();
            st_release_gpu(p.flags + (long long)b * nb + I, p.epoch);   // the chain moves on
        }
        BOPY_TRACE(5);
        if (tid < NC) {
            ss_part = ((partS[tid] + partS[NC + tid]) + (partS[2 * NC + tid] + partS[3 * NC + tid])) +
                      ((partS[4 * NC + tid] + partS[5 * NC + tid]) + (partS[6 * NC + tid] + partS[7 * NC + tid]));
        }
        if (!last) {
            if (tid < NC) {
                double* const dst = p.part + (((long long)b * nb + I) * 2) * NC;
                dst[tid] = mean_part;
                dst[NC + tid] = ss_part;
            }
            consumer_sync();
            if (tid == 0) {
                __threadfence();
                st_release_gpu(p.flags2 + (long long)b * nb + I, p.epoch);
            }
        } else {
            // ---- epilogue (CTA of the last block row): partials in block-row order, acquisition, arg-min ----------
            MinLoc mine;
            mine.val = 0.0;
            mine.idx = -1;
            if (warp == 0) {   // the partials of every earlier block row, published under their second flag
                for (int J = lane; J < I; J += 32) {
                    const unsigned* const f = p.flags2 + (long long)b * nb + J;
                    long long spins = 0;
                    while (ld_acquire_gpu(f) != p.epoch)
                        if (++spins > PROBE_SPIN_LIMIT) __trap();
                }
            }
            consumer_sync();
            if (tid < NC) {
                double mean_c = 0.0, ss_c = 0.0;
                for (int J = 0; J < I; ++J) {
                    const double* const src = p.part + (((long long)b * nb + J) * 2) * NC;
                    mean_c += ld_cg(src + tid);
                    ss_c += ld_cg(src + NC + tid);
                }
                mean_c += mean_part;
                ss_c += ss_part;
                const long long gcand = c0 + tid;
                if (gcand < p.m) {
                    const double mean = __dadd_rn(__dmul_rn(p.y_std, mean_c), p.y_mean);
                    const double var = __dmul_rn(__dadd_rn(p.kss, -ss_c), p.y_var);
                    if (p.mean_out) p.mean_out[gcand] = mean;
                    if (p.var_out) p.var_out[gcand] = var;
                    if (p.acq != A_NONE) {
                        const double av = acquisition_value(p.acq, mean, var, p.eta, p.kappa);
                        if (p.acq_out) p.acq_out[gcand] = av;
                        if (!(p.nan_skip && av != av)) {
                            mine.val = av;
                            mine.idx = p.index_base + gcand;
                        }
                    }
                }
            }
            if (p.records != nullptr && warp == 0) {   // NC <= 32: the batch lives in warp 0
                mine = minloc_warp_reduce(mine);
                if (lane == 0) p.records[b] = mine;
            }
        }
        consumer_sync();   // xs_s, Rs and the partial buffers are reused by the next batch
    }
}

}  // namespace bopy
