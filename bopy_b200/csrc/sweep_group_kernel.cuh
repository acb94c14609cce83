// Group mode of the fused sweep: G CTAs share ONE candidate tile, so that the solve workspace in flight fits the L2.
//
// sweep_kernel gives every CTA its own tile: 148 tiles in flight, 148 x n_pad x 128 x 8 B of V (310 MB at n = 2048) against
// a 126 MB L2 -- every V slice is re-read (n/128)/2 times and half of those reads go to DRAM (ncu: 218 GB per 2^21
// candidates for 100 MB of algorithmic input).  Here the unit of work is a JOB = (tile, block row I):
//     R_I = K*_I - sum_{J<I} L_IJ V_J,   V_I = inv(L_II) R_I,   partial mean / sum v^2 of the row
// exactly the arithmetic of one block-row iteration of sweep_kernel (same policies, same packed factor, same fragment
// order), but the V_J it consumes were published by OTHER CTAs of its group, through the group's workspace slot in global
// memory and release/acquire flags (the protocol of probe_kernel.cuh).  Jobs of a group are claimed from an atomic
// counter in a fixed order (rows ascending within a tile; the first `lead` rows of the next tile interleaved with the last
// `lead` rows of the current one, so that the dependency chain V_0 -> V_1 -> ... of a tile never idles the group); the CTA
// that runs the last row of a tile adds the rows' partial sums in block-row order (the very order sweep_kernel uses: the
// results are bit-identical) and runs the epilogue.  Tiles in flight: ceil(148 / G) x slots instead of 148.
//
// No co-residency assumption: groups are formed from an atomic ticket, a job is only ever claimed by a running CTA, and a
// job waits only for jobs claimed before it.  Waits are bounded (trap after ~20 s instead of hanging the GPU).
#pragma once
#include "probe_kernel.cuh"   // ld_acquire_gpu / st_release_gpu / ld_cg
#include "sweep_kernel.cuh"

namespace bopy {

// control block (unsigned words, zeroed before every launch): ticket, next tile, job counter per group,
// (sequence, tile) of the tile in every slot, V_I-published sequence numbers per (slot, block row)
struct GroupCtl {
    __host__ __device__ static size_t job_off() { return 2; }
    __host__ __device__ static size_t tile_off(int ng) { return (size_t)((2 + ng + 1) & ~1); }
    __host__ __device__ static size_t ready_off(int ng, int S) { return tile_off(ng) + (size_t)2 * ng * S; }
    __host__ __device__ static size_t words(int ng, int S, int R) { return ready_off(ng, S) + (size_t)ng * S * R; }
};

// the group's j-th job -> (k-th tile of the group, block row I).  lead = c: rows c .. R-c-1 of tile k, then
// (k, R-c), (k+1, 0), (k, R-c+1), (k+1, 1), ...; every (k, I) appears once and after (k, I-1) and (k-2, R-1).
__host__ __device__ __forceinline__ void group_job(unsigned j, int R, int c, int& k, int& I) {
    if ((int)j < c) {
        k = 0;
        I = (int)j;
        return;
    }
    j -= (unsigned)c;
    k = (int)(j / (unsigned)R);
    const int r = (int)(j % (unsigned)R), body = R - 2 * c;
    if (r < body) {
        I = c + r;
    } else {
        const int q = r - body;
        if (q & 1) {
            k += 1;
            I = q >> 1;
        } else {
            I = R - c + (q >> 1);
        }
    }
}

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// wait until *f >= want (sequence numbers only grow within a launch)
__device__ __forceinline__ void group_wait_ge(const unsigned* f, unsigned want) {
    if (ld_acquire_gpu(f) >= want) return;
    const long long t0 = global_timer_ns();
    unsigned spins = 0;
    while (ld_acquire_gpu(f) < want) {
        __nanosleep(32);
        if ((++spins & 0xffffu) == 0 && global_timer_ns() - t0 > 20000000000LL) __trap();
    }
}

// wait until the slot holds the group's seq-th tile (1-based); returns its tile number, -1 = past the end
__device__ __forceinline__ long long group_wait_tile(const unsigned long long* f, unsigned seq) {
    unsigned long long v = ld_acquire_gpu_u64(f);
    if ((unsigned)(v >> 32) != seq) {
        const long long t0 = global_timer_ns();
        unsigned spins = 0;
        while ((unsigned)((v = ld_acquire_gpu_u64(f)) >> 32) != seq) {
            __nanosleep(32);
            if ((++spins & 0xffffu) == 0 && global_timer_ns() - t0 > 20000000000LL) __trap();
        }
    }
    return (long long)(unsigned)v - 1;
}

template <class E, int KIND>
__global__ void __launch_bounds__(NT_ALL, 1) sweep_group_kernel(const SweepParams p) {
    static_assert(!E::kMixed, "group mode: one element type for both GEMMs");
    using PG = typename E::PG;
    using PD = typename E::PD;
    using TG = typename E::TG;
    using TD = typename E::TD;
    constexpr int CHG = E::CHG, CHD = E::CHD;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* const stA = smem_raw;
    unsigned char* const stB = smem_raw + STAGES * TILE_BYTES;
    unsigned char* const rs_raw = smem_raw + 2 * STAGES * TILE_BYTES;
    TD* const Rs = reinterpret_cast<TD*>(rs_raw);
    double* const xrow_alias = reinterpret_cast<double*>(rs_raw);
    double* const partM = reinterpret_cast<double*>(rs_raw + 48 * 1024);
    double* const partS = reinterpret_cast<double*>(rs_raw + 52 * 1024);
    double* const xs_s = reinterpret_cast<double*>(rs_raw + (size_t)BM * BN * sizeof(TD));
    unsigned char* const tail = reinterpret_cast<unsigned char*>(xs_s + (size_t)p.d * BN);
    uint64_t* const full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* const empty = full + STAGES;
    uint64_t* const xbar = empty + STAGES;
    uint64_t* const jobbar = xbar + 1;                                   // a job descriptor was posted
    volatile int* const jobq = reinterpret_cast<volatile int*>(tail + 80);   // [2][4]: tile, I, slot, k | staged << 30
    volatile unsigned* const ticket_s = reinterpret_cast<volatile unsigned*>(tail + 112);
    MinLoc* const red = reinterpret_cast<MinLoc*>(tail + 128);
    double* const xrow = p.xrow_separate ? reinterpret_cast<double*>(tail + 192) : xrow_alias;
    // p.xs_stage: the producer copies a job's raw candidate rows ([BN][d], one bulk copy) to shared memory ahead of the job
    uint64_t* const cbar = reinterpret_cast<uint64_t*>(tail + 120);
    const double* const xs_raw = reinterpret_cast<const double*>(tail + 192 + (size_t)(p.d + 1) * BM * sizeof(double));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = p.n_blocks, S = p.group_slots;
    const int ngroups = ((int)gridDim.x + p.group_size - 1) / p.group_size;
    const unsigned char* const Lt = reinterpret_cast<const unsigned char*>(p.Lt);
    const long long slot_bytes = (long long)R * BM * BN * sizeof(TG);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NT / 32);
        }
        mbar_init(xbar, 1);
        mbar_init(jobbar, 1);
        mbar_init(cbar, 1);
        fence_mbar_init();
        *ticket_s = atomicAdd(p.gctl, 1u);
    }
    __syncthreads();
    const int group = (int)(*ticket_s) / p.group_size;
    unsigned* const job_ctr = p.gctl + GroupCtl::job_off() + group;
    unsigned long long* const tile_of =
        reinterpret_cast<unsigned long long*>(p.gctl + GroupCtl::tile_off(ngroups)) + (long long)group * S;
    unsigned* const ready = p.gctl + GroupCtl::ready_off(ngroups, S) + (long long)group * S * R;
    unsigned char* const Vgroup = reinterpret_cast<unsigned char*>(p.Vws) + (long long)group * S * slot_bytes;
    double* const part_group = p.gpart + (long long)group * S * R * 2 * BN;

    if (warp >= NT / 32) {
        // =============================== job claims + TMA producer (one lane) =============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
        if (warp != NT / 32 || lane != 0) return;
        uint32_t g = 0, dn = 0;
        long long staged_tile = -1;
        for (;;) {
            const unsigned j = atomicAdd(job_ctr, 1u);
            int k, I;
            group_job(j, R, p.group_lead, k, I);
            const int slot = k % S;
            long long tile;
            if (I == 0) {
                // tile numbers are handed out in the order of the group's tiles (so that the tiles past the end are a suffix
                // of every group's sequence): wait until the previous tile of the group has its number
                if (k > 0) group_wait_tile(tile_of + (k - 1) % S, (unsigned)k);
                const unsigned t = atomicAdd(p.gctl + 1, 1u);
                tile = (long long)t < p.ntiles ? (long long)t : -1;
                // the slot is free once the last row of the tile that used it before has finished
                if (tile >= 0 && k >= S) group_wait_ge(ready + (long long)slot * R + (R - 1), (unsigned)(k - S + 1));
                st_release_gpu_u64(tile_of + slot, ((unsigned long long)(unsigned)(k + 1) << 32) | (unsigned)(tile + 1));
            } else {
                tile = group_wait_tile(tile_of + slot, (unsigned)(k + 1));
            }
            if (tile < 0 && I < p.group_lead) continue;   // a head row of a tile past the end: rows of the last tile may follow
            // candidates of a new, full tile: staged by one bulk copy (the consumers of the previous job are past their
            // kernel-tile step, the only reader of the staging buffer)
            const bool stage_xs = p.xs_stage && tile >= 0 && tile != staged_tile && (tile + 1) * BN <= p.m;
            volatile int* const q = jobq + 4 * (dn & 1u);
            q[0] = (int)tile;
            q[1] = I;
            q[2] = slot;
            q[3] = k | (stage_xs ? (1 << 30) : 0);
            mbar_arrive(jobbar);
            ++dn;
            if (tile < 0) break;                          // body or tail row of a tile past the end: nothing follows
            staged_tile = tile;
            if (stage_xs) {
                const uint32_t bytes = (uint32_t)BN * p.d * sizeof(double);
                mbar_arrive_expect_tx(cbar, bytes);
                bulk_g2s(const_cast<double*>(xs_raw), p.Xs + tile * BN * p.d, bytes, cbar);
            }
            if (p.xrow_separate) {
                const uint32_t bytes = (uint32_t)(p.d + 1) * BM * sizeof(double);
                mbar_arrive_expect_tx(xbar, bytes);
                bulk_g2s(xrow, p.Xt + (long long)I * (p.d + 1) * BM, bytes, xbar);
            }
            const unsigned char* const Vt = Vgroup + (long long)slot * slot_bytes;
            const unsigned* const rdy = ready + (long long)slot * R;
            const int T_gemm = I * CHG, T_all = T_gemm + CHD;
            const unsigned char* const a_row = Lt + E::row_base(I) * TILE_BYTES;
            for (int t = 0; t < T_all; ++t, ++g) {
                const uint32_t stage = g % STAGES;
                if (t < T_gemm && t % CHG == 0) {
                    group_wait_ge(rdy + t / CHG, (unsigned)(k + 1));   // V_J published by its owner
                    fence_proxy_async();                               // ... before this lane's bulk loads of it
                }
                mbar_wait(&empty[stage], ((g / STAGES) & 1u) ^ 1u);
                if (t < T_gemm) {
                    mbar_arrive_expect_tx(&full[stage], 2 * TILE_BYTES);
                    bulk_g2s(stA + stage * TILE_BYTES, a_row + (long long)t * TILE_BYTES, TILE_BYTES, &full[stage]);
                    bulk_g2s(stB + stage * TILE_BYTES, Vt + (long long)t * TILE_BYTES, TILE_BYTES, &full[stage]);
                } else {
                    mbar_arrive_expect_tx(&full[stage], TILE_BYTES);
                    bulk_g2s(stA + stage * TILE_BYTES, a_row + (long long)t * TILE_BYTES, TILE_BYTES, &full[stage]);
                }
            }
        }
        return;
    }

    // ===================================== compute warps ==================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_COMPUTE));
    const PG pg(tid);
    const PD pd(tid);
    uint32_t gcount = 0, xphase = 0, cphase = 0;
    long long cur_tile = -1;
    MinLoc best;
    best.val = 0.0;
    best.idx = -1;

    for (uint32_t dn = 0;; ++dn) {
        mbar_wait(jobbar, dn & 1u);
        const long long tile = jobq[4 * (dn & 1u) + 0];
        if (tile < 0) break;
        const int I = jobq[4 * (dn & 1u) + 1], slot = jobq[4 * (dn & 1u) + 2], kf = jobq[4 * (dn & 1u) + 3];
        const int k = kf & ~(1 << 30);
        const bool staged = (kf >> 30) & 1;
        const long long c0 = tile * BN;
        TG* const Vt = reinterpret_cast<TG*>(Vgroup + (long long)slot * slot_bytes);
        const int T_gemm = I * CHG;

        if (!p.xrow_separate && tid == 0) {   // lands in the Rs region, free since the previous job's last barrier
            const uint32_t bytes = (uint32_t)(p.d + 1) * BM * sizeof(double);
            mbar_arrive_expect_tx(xbar, bytes);
            bulk_g2s(xrow, p.Xt + (long long)I * (p.d + 1) * BM, bytes, xbar);
        }
        if (tile != cur_tile) {
            if (staged) {
                mbar_wait(cbar, cphase);
                cphase ^= 1;
            }
            for (int e = tid; e < BN * p.d; e += NT) {
                const int c = e / p.d, q = e - c * p.d;
                const long long gc = c0 + c;
                const double v = staged ? xs_raw[e] : (gc < p.m ? p.Xs[gc * p.d + q] : 0.0);
                xs_s[q * BN + c] = __ddiv_rn(v, p.ls[q]);
            }
            cur_tile = tile;
            consumer_sync();
        }

        // ---- kernel tile K*[block row I, this tile's candidates] and the row's share of the mean ---------------
        TG acc[PG::RI][PG::CJ];
        double mean_part = 0.0;
        mbar_wait(xbar, xphase);
        xphase ^= 1;
        {
            constexpr int RH = PG::RI / 2;
            double mp[PG::CJ];
#pragma unroll
            for (int j = 0; j < PG::CJ; ++j) mp[j] = 0.0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double d2[RH][PG::CJ];
#pragma unroll
                for (int i = 0; i < RH; ++i)
#pragma unroll
                    for (int j = 0; j < PG::CJ; ++j) d2[i][j] = 0.0;
                for (int q = 0; q < p.d; ++q) {
                    double xr[RH], xc[PG::CJ];
#pragma unroll
                    for (int i = 0; i < RH; ++i) xr[i] = xrow[q * BM + pg.row_of(h * RH + i)];
#pragma unroll
                    for (int j = 0; j < PG::CJ; ++j) xc[j] = xs_s[q * BN + pg.cand_of(j)];
#pragma unroll
                    for (int i = 0; i < RH; ++i)
#pragma unroll
                        for (int j = 0; j < PG::CJ; ++j) {
                            const double df = xc[j] - xr[i];
                            d2[i][j] = fma(df, df, d2[i][j]);
                        }
                }
#pragma unroll
                for (int i = 0; i < RH; ++i) {
                    const int row = pg.row_of(h * RH + i);
                    const double amp_i = (I * BM + row < p.n) ? p.amp : 0.0;
                    const double a_i = xrow[p.d * BM + row];
#pragma unroll
                    for (int j = 0; j < PG::CJ; ++j) {
                        const double kv = __dmul_rn(amp_i, base_kernel<KIND>(d2[i][j]));
                        acc[h * RH + i][j] = static_cast<TG>(kv);
                        mp[j] = fma(kv, a_i, mp[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < PG::CJ; ++j) mp[j] = pg.reduce_rows(mp[j]);
            if (pg.leader) {
#pragma unroll
                for (int j = 0; j < PG::CJ; ++j) partM[pg.part * BN + pg.cand_of(j)] = mp[j];
            }
            consumer_sync();
            if (tid < BN) mean_part = ((partM[tid] + partM[BN + tid]) + partM[2 * BN + tid]) + partM[3 * BN + tid];
            consumer_sync();  // xrow / partM consumed: Rs may be overwritten from here on
        }

        // ---- R_I = K*_I - sum_J L_IJ V_J -------------------------------------------------------------------------
        for (int t = 0; t < T_gemm; ++t, ++gcount) {
            const uint32_t stage = gcount % STAGES;
            mbar_wait(&full[stage], (gcount / STAGES) & 1u);
            pg.template mma_tile<false>(acc, reinterpret_cast<const TG*>(stA + stage * TILE_BYTES),
                                        reinterpret_cast<const TG*>(stB + stage * TILE_BYTES), -1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
        }
#pragma unroll
        for (int i = 0; i < PG::RI; ++i) {
            const int row = pg.row_of(i);
#pragma unroll
            for (int jv = 0; jv < PG::CJ / PG::CV; ++jv)
                store_cands(reinterpret_cast<TG*>(Rs) + PD::b_index(row, pg.cand_of(jv * PG::CV)), &acc[i][jv * PG::CV], 0);
        }
        consumer_sync();

        // ---- V_I = inv(L_II) R_I ---------------------------------------------------------------------------------
        TD accd[PD::RI][PD::CJ];
#pragma unroll
        for (int i = 0; i < PD::RI; ++i)
#pragma unroll
            for (int j = 0; j < PD::CJ; ++j) accd[i][j] = static_cast<TD>(0);
        for (int kc = 0; kc < CHD; ++kc, ++gcount) {
            const uint32_t stage = gcount % STAGES;
            mbar_wait(&full[stage], (gcount / STAGES) & 1u);
            pd.template mma_tile<true>(accd, reinterpret_cast<const TD*>(stA + stage * TILE_BYTES), Rs + kc * PD::KC * BN, kc);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
        }

        // ---- V_I to the group's slot, sum v^2 of the row -----------------------------------------------------------
        const bool last_row = I + 1 == R;
        {
            TG* const Vrow = Vt + (long long)I * BM * BN;
            double sq[PD::CJ];
#pragma unroll
            for (int j = 0; j < PD::CJ; ++j) sq[j] = 0.0;
#pragma unroll
            for (int i = 0; i < PD::RI; ++i) {
                if (!last_row) {
                    const int row = pd.row_of(i);
#pragma unroll
                    for (int jv = 0; jv < PD::CJ / PD::CV; ++jv)
                        store_cands(&Vrow[PG::b_index(row, pd.cand_of(jv * PD::CV))], &accd[i][jv * PD::CV], 0);
                }
#pragma unroll
                for (int j = 0; j < PD::CJ; ++j) {
                    const double v = static_cast<double>(accd[i][j]);
                    sq[j] = fma(v, v, sq[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < PD::CJ; ++j) sq[j] = pd.reduce_rows(sq[j]);
            if (pd.leader) {
#pragma unroll
                for (int j = 0; j < PD::CJ; ++j) partS[pd.part * BN + pd.cand_of(j)] = sq[j];
            }
            fence_proxy_async();   // generic stores (V, Rs reads) before later bulk copies: other CTAs' loads of V_I, the next X/l row
        }
        consumer_sync();
        double ss_part = 0.0;
        if (tid < BN) ss_part = ((partS[tid] + partS[BN + tid]) + partS[2 * BN + tid]) + partS[3 * BN + tid];
        double* const part_row = part_group + ((long long)slot * R + I) * 2 * BN;
        unsigned* const rdy = ready + (long long)slot * R;

        if (!last_row) {
            if (tid < BN) {
                __stcg(part_row + tid, mean_part);
                __stcg(part_row + BN + tid, ss_part);
            }
            consumer_sync();   // every V_I / partial store of this CTA is ordered before the release below
            if (tid == 0) {
                __threadfence();
                st_release_gpu(rdy + I, (unsigned)(k + 1));
            }
            continue;
        }

        // ---- last row of the tile: totals in block-row order, epilogue, arg-min ------------------------------------
        MinLoc mine;
        mine.val = 0.0;
        mine.idx = -1;
        if (tid < BN) {
            double mean_c = 0.0, ss_c = 0.0;
            const double* const prow = part_group + (long long)slot * R * 2 * BN;
            // every row's flag was acquired by the producer lane before it loaded that row's V, and those loads were
            // consumed through the ring's barriers; one more acquire of the last flag here, then L2-coherent loads
            if (R > 1) group_wait_ge(rdy + (R - 2), (unsigned)(k + 1));
            for (int J = 0; J + 1 < R; ++J) {
                mean_c += ld_cg(prow + (long long)J * 2 * BN + tid);
                ss_c += ld_cg(prow + (long long)J * 2 * BN + BN + tid);
            }
            mean_c += mean_part;
            ss_c += ss_part;
            const long long gc = c0 + tid;
            if (gc < p.m) {
                const double mean = __dadd_rn(__dmul_rn(p.y_std, mean_c), p.y_mean);
                const double var = __dmul_rn(__dadd_rn(p.kss, -ss_c), p.y_var);
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
            if (warp < BN / 32) {
                mine = minloc_warp_reduce(mine);
                if (lane == 0) red[warp] = mine;
            }
            consumer_sync();
            if (tid == 0) {
                MinLoc tbest = red[0];
                for (int w = 1; w < BN / 32; ++w)
                    if (minloc_better(red[w], tbest)) tbest = red[w];
                if (p.tile_records != nullptr) p.tile_records[tile] = tbest;
                if (minloc_better(tbest, best)) best = tbest;
            }
        }
        consumer_sync();   // partial sums of the slot read, red reusable
        if (tid == 0) {
            __threadfence();
            st_release_gpu(rdy + I, (unsigned)(k + 1));   // the slot may take its next tile
        }
    }
    if (tid == 0 && p.partials != nullptr) p.partials[blockIdx.x] = best;
}
#endif  // __CUDACC__

}  // namespace bopy
