// EXPERIMENT, off by default (BOPY_B200_WARP_KERNEL=1 selects it): a warp-autonomous variant of the fused hot path for
// n <= 256 (one or two block rows of L; BASELINE config C3).  Correct (parity tests), but measured SLOWER than sweep_kernel at
// C3: 4.30 ms against 3.33 ms per 2^20 candidates.  ncu (profiles/r02/ncu_sweep_warp_c3_summary.json): DMMA 47.5 % + FP64
// 11.8 % of cycles against 64.2 % + 13.0 %; a third of the warps' time goes to the kernel-tile step, which runs 3.7x slower
// than in sweep_kernel -- DFMA and DMMA share the FP64 pipe, so evaluating exponentials beside another warp's DMMAs is
// zero-sum for the pipe and only adds queueing, and the 4000-instruction unrolled step no longer fits the instruction cache
// (stall_no_inst is its top stall reason).  Kept for the record and for A/B runs; the idea it tests:
//
// sweep_kernel splits a 128 x 128 (rows x candidates) tile over 8 warps as 4 row groups x 2 candidate halves, so every
// phase of a block row (kernel tile, residual, diagonal solve, publish) ends in a CTA-wide barrier and V_0 travels through the
// global workspace.  With two block rows those fixed costs are a quarter of the time (ncu at C3: FP64 pipe 77 % busy).
// Here a warp owns 16 candidates and ALL 128 rows of a block row:
//   * every dependency of the blocked solve stays inside the warp: the residual tile R_I goes to the warp's own 16 columns
//     of a shared-memory tile (only __syncwarp), V_I = inv(L_II) R_I overwrites R_I there row atom by row atom (atom a of R
//     is dead once step kc = a of the triangular product has read it) and is the B operand of the next block row's product:
//     no workspace, no V traffic at all; the row reductions (mean, sum v^2) are warp shuffles;
//   * the only shared resource is the ring of L tiles (8 KB each, one bulk async copy, `full` mbarriers), so warps drift
//     apart by up to the ring depth instead of meeting at barriers.  There is no producer warp: the packed factor of a
//     handle with <= 2 block rows is one contiguous run of 16 or 48 tiles that every 64-candidate pass walks in order, so
//     the warp that releases a ring stage LAST (a shared-memory counter per stage) issues the copy of the tile that stage
//     holds next.  (A fifth warp would cap the kernel at 168 registers -- two CTAs x six warp slots -- and ptxas then keeps
//     ONE operand register, every pair of DMMAs waiting for a shared-memory load; with four warps it has 255.)
//   * a CTA is 4 warps (64 candidates at a time, a 128-candidate tile as two halves) and < 113 KB of shared memory: TWO
//     CTAs per SM, which are out of phase with each other -- one fills the FP64 pipe while the other evaluates
//     exponentials or runs its epilogue.
// Same packed factor, fragment maps, k order and epilogue as sweep_kernel (V is bit-identical); the mean and sum v^2 add
// the rows of a block row in another order (one chain over the 16 row atoms + 3 shuffles), a difference of rounding.
//
// Reference arithmetic: see sweep_kernel.cuh.
#pragma once
#include "sweep_kernel.cuh"

namespace bopy {

constexpr int WK_WARPS = 4;                    // compute warps per CTA
constexpr int WK_NT = WK_WARPS * 32;           // compute threads
constexpr int WK_HALF = 16 * WK_WARPS;         // candidates per pass (half of a 128-candidate tile)
constexpr int WK_MAX_BLOCKS = 2;
constexpr size_t WK_SMEM_LIMIT = 113 * 1024;   // two CTAs per SM

__host__ __device__ constexpr size_t wk_smem_bytes(int d, int n_blocks, int stages) {
    return (size_t)stages * TILE_BYTES + (size_t)BM * WK_HALF * sizeof(double) +            // ring, residual / V tile
           (size_t)n_blocks * (d + 1) * BM * sizeof(double) + (size_t)d * WK_HALF * sizeof(double) +   // X/l + alpha, candidates
           (size_t)2 * stages * sizeof(uint64_t) + 64 + 3 * WK_WARPS * sizeof(MinLoc) + WK_WARPS * 32 * sizeof(double);
}
// ring depth for (d, n_blocks): as deep as two CTAs per SM allow, 0 = does not fit (the blocked kernel serves the handle)
inline int wk_stages(int d, int n_blocks) {
    for (int s = 6; s >= 3; --s)
        if (wk_smem_bytes(d, n_blocks, s) <= WK_SMEM_LIMIT) return s;
    return 0;
}

#ifdef __CUDACC__
// one k4-step of the triangular product for row atoms A0 .. 15: acc[a] += Ltile[atom a] (x) b
// A fragment of row atom a (k = kq): element offset kq * BM + ((8 a + q8) ^ (4 kq)) = base[a & 1] + 8 a, because the XOR only
// flips bit 2 (inside q8) and bit 3 (the parity of a): two registers instead of sixteen
__device__ __forceinline__ double wk_lds(const double* p) {   // volatile: keeps its place ahead of the DMMAs it is issued before
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(smem_u32(p)));
    return v;
}
// (the A load of atom a + 1 is issued before the products of atom a: with 64 accumulators in 168 registers the compiler keeps
// a single operand register otherwise and every pair of DMMAs waits for a shared-memory load)
__device__ __forceinline__ void wk_full_step(double (&acc)[16][4], const double* __restrict__ As, const int (&abase)[2], int s,
                                             double b0, double b1) {
    double av = wk_lds(As + abase[0] + s * 4 * BM);
#pragma unroll
    for (int a = 0; a < 16; ++a) {
        const double avn = a < 15 ? wk_lds(As + abase[(a + 1) & 1] + 8 * (a + 1) + s * 4 * BM) : 0.0;
        dmma_m8n8k4(acc[a][0], acc[a][1], av, b0);
        dmma_m8n8k4(acc[a][2], acc[a][3], av, b1);
        av = avn;
    }
}

// One k4-step of tile kc of inv(L_II) (lower triangular: row atoms a >= kc only) against rows 8 kc + 4 s .. of the residual.
// The atoms are ONE unrolled sequence entered at atom kc (a switch with fall-through: a jump table into straight-line code),
// not sixteen specialised copies: the sixteen-way unrolled triangle was 40 KB of instructions, past the instruction cache.
#define BOPY_WK_ATOM(A)                                                                                  \
    case A: {                                                                                            \
        const double avn = (A) < 15 ? wk_lds(As + abase[((A) + 1) & 1] + 8 * ((A) + 1) + s * 4 * BM) : 0.0; \
        dmma_m8n8k4(acc[A][0], acc[A][1], av, b0);                                                       \
        dmma_m8n8k4(acc[A][2], acc[A][3], av, b1);                                                       \
        av = avn;                                                                                        \
    }
__device__ __forceinline__ void wk_tri_step(double (&acc)[16][4], const double* __restrict__ As, const int (&abase)[2], int s,
                                            double b0, double b1, int kc) {
    double av = wk_lds(As + abase[kc & 1] + 8 * kc + s * 4 * BM);
    switch (kc) {
        BOPY_WK_ATOM(0) BOPY_WK_ATOM(1) BOPY_WK_ATOM(2) BOPY_WK_ATOM(3) BOPY_WK_ATOM(4) BOPY_WK_ATOM(5) BOPY_WK_ATOM(6)
        BOPY_WK_ATOM(7) BOPY_WK_ATOM(8) BOPY_WK_ATOM(9) BOPY_WK_ATOM(10) BOPY_WK_ATOM(11) BOPY_WK_ATOM(12) BOPY_WK_ATOM(13)
        BOPY_WK_ATOM(14) BOPY_WK_ATOM(15)
    }
}
#undef BOPY_WK_ATOM

// atom kc of V is final after tile kc: it replaces the residual rows it was computed from (row 8 kc + q8, candidates
// 2 kq + {0,1} and + 8: the B layout of the next block row's product)
#define BOPY_WK_STORE(A)                                                                                          \
    case A:                                                                                                       \
        *reinterpret_cast<double2*>(dst) = make_double2(acc[A][0], acc[A][1]);                                    \
        *reinterpret_cast<double2*>(dst + 8 - 16 * ((voff >> 3) & 1)) = make_double2(acc[A][2], acc[A][3]);       \
        break;
__device__ __forceinline__ void wk_store_atom(const double (&acc)[16][4], double* __restrict__ Rw, int voff, int kc) {
    double* const dst = Rw + kc * 8 * WK_HALF + voff;
    switch (kc) {
        BOPY_WK_STORE(0) BOPY_WK_STORE(1) BOPY_WK_STORE(2) BOPY_WK_STORE(3) BOPY_WK_STORE(4) BOPY_WK_STORE(5) BOPY_WK_STORE(6)
        BOPY_WK_STORE(7) BOPY_WK_STORE(8) BOPY_WK_STORE(9) BOPY_WK_STORE(10) BOPY_WK_STORE(11) BOPY_WK_STORE(12)
        BOPY_WK_STORE(13) BOPY_WK_STORE(14) BOPY_WK_STORE(15)
    }
}
#undef BOPY_WK_STORE

// The warp that released a stage last (the counter read WK_WARPS - 1 before its increment) copies the tile the stage holds
// next.  The counter's old value is looked at one stage later, so that nobody waits for the shared-memory atomic.
__device__ __forceinline__ void wk_refill(unsigned* released, uint64_t* full, unsigned char* ring, const unsigned char* Lt,
                                          int stages, uint32_t T_pass, uint32_t total, int lane, unsigned& pend_tok,
                                          uint32_t pend_g) {
    if (lane == 0 && pend_tok == WK_WARPS - 1) {
        const uint32_t stage = pend_g % stages, gn = pend_g + stages;
        released[stage] = 0;
        if (gn < total) {
            mbar_arrive_expect_tx(&full[stage], TILE_BYTES);
            bulk_g2s(ring + stage * TILE_BYTES, Lt + (size_t)(gn % T_pass) * TILE_BYTES, TILE_BYTES, &full[stage]);
        }
    }
    pend_tok = 0xffffffffu;
}
__device__ __forceinline__ void wk_release(unsigned* released, uint64_t* full, unsigned char* ring, const unsigned char* Lt,
                                           uint32_t stage, uint32_t g, int stages, uint32_t T_pass, uint32_t total, int lane,
                                           unsigned& pend_tok, uint32_t& pend_g) {
    wk_refill(released, full, ring, Lt, stages, T_pass, total, lane, pend_tok, pend_g);
    __syncwarp();                       // every lane's reads of the stage are done (their DMMAs were issued)
    if (lane == 0) {
        __threadfence_block();
        pend_tok = atomicAdd(&released[stage], 1u);
    }
    pend_g = g;
}

template <int KIND>
__global__ void __launch_bounds__(WK_NT, 2) sweep_warp_kernel(const SweepParams p, const int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* const ring = smem_raw;                                                    // [stages] L tiles
    double* const Rs = reinterpret_cast<double*>(smem_raw + (size_t)stages * TILE_BYTES);     // [BM][WK_HALF]
    double* const xrows = Rs + (size_t)BM * WK_HALF;                                          // [n_blocks][d+1][BM]
    double* const xs_s = xrows + (size_t)p.n_blocks * (p.d + 1) * BM;                         // [d][WK_HALF] candidates / l
    uint64_t* const full = reinterpret_cast<uint64_t*>(xs_s + (size_t)p.d * WK_HALF);
    unsigned* const released = reinterpret_cast<unsigned*>(full + stages);   // [stages] warps done with the stage's current tile
    MinLoc* const red = reinterpret_cast<MinLoc*>(reinterpret_cast<unsigned char*>(full + 2 * stages) + 64);   // [WK_WARPS]
    MinLoc* const bests = red + WK_WARPS;                    // [2][WK_WARPS]: per-warp best of the tile / of the whole sweep
    double* const sums = reinterpret_cast<double*>(bests + 2 * WK_WARPS);   // [WK_WARPS][8][4] mean / sum v^2 of the warp's candidates

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = p.n_blocks;
    const unsigned char* const Lt = reinterpret_cast<const unsigned char*>(p.Lt);

    // one pass (64 candidates) walks the T_pass tiles of the packed factor in storage order: block row 0 = 16 tiles of
    // inv(L_00); block row 1 = 16 tiles of -L_10, then 16 of inv(L_11)
    const uint32_t T_pass = R == 1 ? 16u : 48u;
    uint32_t total = 0;                      // tiles this CTA consumes in all
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) total += ((tile * BN + WK_HALF < p.m) ? 2u : 1u) * T_pass;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            released[s] = 0;
        }
        fence_mbar_init();
        for (uint32_t g = 0; g < (uint32_t)stages && g < total; ++g) {
            mbar_arrive_expect_tx(&full[g], TILE_BYTES);
            bulk_g2s(ring + g * TILE_BYTES, Lt + (size_t)(g % T_pass) * TILE_BYTES, TILE_BYTES, &full[g]);
        }
    }
    // the (at most two) block rows of X/l and alpha stay resident
    for (int e = tid; e < R * (p.d + 1) * BM; e += WK_NT) xrows[e] = p.Xt[e];
    __syncthreads();

    // ===================================== compute warps (autonomous) =================================================
    const int kq = lane & 3, q8 = lane >> 2;
    double* const Rw = Rs + 16 * warp;               // this warp's 16 columns of the residual / V tile
    double* const xw = xs_s + 16 * warp;             // ... and of the staged candidates
    int abase[2], boff[2];
    abase[0] = DmmaPolicy::a_index(kq, q8);            // even atoms: a_index(kq, 8 a + q8) - 8 a
    abase[1] = DmmaPolicy::a_index(kq, 8 + q8) - 8;    // odd atoms
    // B fragment (k = kq, column 8 jj + q8) in the [row][WK_HALF] tile; the XOR swizzle of the 128-wide tiles keyed by row & 3
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) boff[jj] = kq * WK_HALF + ((8 * jj + q8) ^ (4 * kq));
    // where this lane's accumulator pair (row q8 of an atom, candidates 2 kq + {0,1}) goes as a B element: row & 3 == q8 & 3
    const int voff = q8 * WK_HALF + ((2 * kq) ^ (4 * (q8 & 3)));
    uint32_t gcount = 0;
    unsigned pend_tok = 0xffffffffu;   // lane 0: the release counter's value before this warp's last release ...
    uint32_t pend_g = 0;               // ... and the sequence number of the tile it released
    // rarely touched state lives in shared memory: 64 accumulators leave ~40 registers for everything else
    double* const wsum = sums + warp * 32;           // [j < 4: mean, 4 + j: sum v^2][kq], owned by the lanes with q8 == 0
    MinLoc* const tbest = bests + warp;
    MinLoc* const best = bests + WK_WARPS + warp;    // both owned by lane 0
    if (lane == 0) {
        best->val = 0.0;
        best->idx = -1;
    }

    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        if (lane == 0) {
            tbest->val = 0.0;
            tbest->idx = -1;
        }
        const int halves = (tile * BN + WK_HALF < p.m) ? 2 : 1;
        for (int h = 0; h < halves; ++h) {
            const long long c0 = tile * BN + h * WK_HALF + 16 * warp;     // this warp's first candidate
            __syncwarp();
            for (int e = lane; e < 16 * p.d; e += 32) {
                const int c = e / p.d, q = e - c * p.d;
                const double v = c0 + c < p.m ? p.Xs[(c0 + c) * p.d + q] : 0.0;
                xw[q * WK_HALF + c] = __ddiv_rn(v, p.ls[q]);
            }
            __syncwarp();
            if (q8 == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) wsum[j * 4 + kq] = 0.0;
            }

            for (int I = 0; I < R; ++I) {
                const double* const xrow = xrows + (size_t)I * (p.d + 1) * BM;
                double acc[16][4];
                // ---- kernel tile K*[block row I, the warp's 16 candidates] and its share of the mean -----------------
                {
                    double mp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {      // 4 row atoms = 16 kernel evaluations in flight per thread
                        double d2[4][4];
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int j = 0; j < 4; ++j) d2[a][j] = 0.0;
                        for (int q = 0; q < p.d; ++q) {
                            double xc[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) xc[j] = xw[q * WK_HALF + 8 * (j >> 1) + 2 * kq + (j & 1)];
#pragma unroll
                            for (int a = 0; a < 4; ++a) {
                                const double xr = xrow[q * BM + 8 * (ch * 4 + a) + q8];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const double df = xc[j] - xr;
                                    d2[a][j] = fma(df, df, d2[a][j]);     // cdist's summation order over the dimensions
                                }
                            }
                        }
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            const int row = 8 * (ch * 4 + a) + q8;
                            const double amp_i = (I * BM + row < p.n) ? p.amp : 0.0;   // padding rows: amplitude 0, no branch
                            const double a_i = xrow[p.d * BM + row];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const double kv = __dmul_rn(amp_i, base_kernel<KIND>(d2[a][j]));
                                acc[ch * 4 + a][j] = kv;
                                mp[j] = fma(kv, a_i, mp[j]);
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double v = mp[j];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        if (q8 == 0) wsum[j * 4 + kq] += v;
                    }
                }

                // ---- R_I = K*_I - sum_{J<I} L_IJ V_J: V_J is the warp's own columns of Rs (n_blocks <= 2: J = 0 only) --------
                for (int t = 0; t < I * 16; ++t, ++gcount) {
                    const uint32_t stage = gcount % stages;
                    mbar_wait(&full[stage], (gcount / stages) & 1u);
                    const double* const As = reinterpret_cast<const double*>(ring + stage * TILE_BYTES);
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const double b0 = Rw[(t * 8 + s * 4) * WK_HALF + boff[0]];
                        const double b1 = Rw[(t * 8 + s * 4) * WK_HALF + boff[1]];
                        wk_full_step(acc, As, abase, s, b0, b1);
                    }
                    wk_release(released, full, ring, Lt, stage, gcount, stages, T_pass, total, lane, pend_tok, pend_g);
                }
                // ---- the residual becomes the B operand of the triangular product (all reads of V_{I-1} are done: DMMA is
                //      warp-synchronous) ---------------------------------------------------------------------------------------
                __syncwarp();
#pragma unroll
                for (int a = 0; a < 16; ++a) {
                    double* const dst = Rw + a * 8 * WK_HALF + voff;
                    *reinterpret_cast<double2*>(dst) = make_double2(acc[a][0], acc[a][1]);
                    *reinterpret_cast<double2*>(dst + 8 - 16 * ((voff >> 3) & 1)) = make_double2(acc[a][2], acc[a][3]);
                }
                __syncwarp();

                // ---- V_I = inv(L_II) R_I, lower triangular: row atom a needs the tiles kc <= a -------------------------------
#pragma unroll
                for (int a = 0; a < 16; ++a)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[a][j] = 0.0;
                const bool keep = I + 1 < R;
                for (int kc = 0; kc < 16; ++kc, ++gcount) {
                    const uint32_t stage = gcount % stages;
                    mbar_wait(&full[stage], (gcount / stages) & 1u);
                    const double* const As = reinterpret_cast<const double*>(ring + stage * TILE_BYTES);
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const double b0 = Rw[(kc * 8 + s * 4) * WK_HALF + boff[0]];
                        const double b1 = Rw[(kc * 8 + s * 4) * WK_HALF + boff[1]];
                        wk_tri_step(acc, As, abase, s, b0, b1, kc);
                    }
                    if (keep) wk_store_atom(acc, Rw, voff, kc);
                    wk_release(released, full, ring, Lt, stage, gcount, stages, T_pass, total, lane, pend_tok, pend_g);
                }
                wk_refill(released, full, ring, Lt, stages, T_pass, total, lane, pend_tok, pend_g);   // before the long kernel-tile step
                // ---- sum v^2 of the block row -----------------------------------------------------------------------------
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double v = 0.0;
#pragma unroll
                    for (int a = 0; a < 16; ++a) v = fma(acc[a][j], acc[a][j], v);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (q8 == 0) wsum[(4 + j) * 4 + kq] += v;
                }
            }

            // ---- epilogue: lanes with q8 < 4 take candidate 8 (q8 >> 1) + 2 kq + (q8 & 1) (accumulator column j = q8) ----------
            MinLoc mine;
            mine.val = 0.0;
            mine.idx = -1;
            __syncwarp();
            {
                const int j = q8 & 3;
                const double m_sel = wsum[j * 4 + kq], s_sel = wsum[(4 + j) * 4 + kq];
                const long long gc = c0 + 8 * (j >> 1) + 2 * kq + (j & 1);
                if (q8 < 4 && gc < p.m) {
                    const double mean = __dadd_rn(__dmul_rn(p.y_std, m_sel), p.y_mean);
                    const double var = __dmul_rn(__dadd_rn(p.kss, -s_sel), p.y_var);
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
                mine = minloc_warp_reduce(mine);
                if (lane == 0 && minloc_better(mine, *tbest)) *tbest = mine;
            }
        }
        if (lane == 0 && minloc_better(*tbest, *best)) *best = *tbest;
        if (p.tile_records != nullptr) {          // per-128-candidate records (segmented arg-min): the one place the warps meet
            if (lane == 0) red[warp] = *tbest;
            asm volatile("bar.sync 1, %0;" ::"n"(WK_NT) : "memory");
            if (tid == 0) {
                MinLoc t = red[0];
                for (int w = 1; w < WK_WARPS; ++w)
                    if (minloc_better(red[w], t)) t = red[w];
                p.tile_records[tile] = t;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(WK_NT) : "memory");
        }
    }
    if (p.partials != nullptr) {
        if (lane == 0) red[warp] = *best;
        asm volatile("bar.sync 1, %0;" ::"n"(WK_NT) : "memory");
        if (tid == 0) {
            MinLoc t = red[0];
            for (int w = 1; w < WK_WARPS; ++w)
                if (minloc_better(red[w], t)) t = red[w];
            p.partials[blockIdx.x] = t;
        }
    }
}
#endif  // __CUDACC__

}  // namespace bopy
