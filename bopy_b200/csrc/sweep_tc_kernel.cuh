// fp32 mode of the fused sweep on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the accumulators in
// tensor memory (TMEM), operands staged in shared memory by bulk async copies (TMA).
//
// Same algorithm and reference arithmetic as sweep_kernel.cuh (K* tile -> blocked forward substitution -> diagonal
// variance -> LCB / EI / POI -> arg-min; $SK/_gpr.py:446-469, bopy/acquisition.py:83-128).  What differs is the
// n^2/2 off-diagonal work  R_I = K*_I - sum_{J<I} L_IJ V_J :
//   * L_IJ and V_J are stored as TF32 PAIRS  x ~ hi + lo  (both rounded to nearest when they are packed / published),
//     in the tensor core's canonical K-major operand layout (tc_common.cuh); the product is the 3xTF32 expansion
//     lo.hi + hi.lo + hi.hi, three tcgen05.mma M=128 N=128 K=8 per 8 columns of L, issued by ONE thread.
//   * The tensor core adds into its fp32 accumulator with TRUNCATION and no guard bits (tools/tcgen05_probe.cu): a long
//     accumulation chain drifts towards zero.  The hi.hi products therefore accumulate over only FOLD_TILES * 8 columns
//     in one of two ping-pong TMEM accumulators; the compute warps read each finished partial sum with tcgen05.ld (in
//     the m16n8 fragment shape) and join it to an fp32 running sum by round-to-nearest adds; the running sum is folded
//     into the fp64 residual tile every FLUSH_BLOCKS 128-column blocks.  The lo.hi / hi.lo products are ~2^-11 of the
//     result, so their own chain (a third accumulator, one per block row) may be as long as the row.
//   * K*, the mean, the diagonal solve V_I = inv(L_II) R_I (DMMA) and sum v^2 stay fp64, as in the mixed engine.
// Warp roles: warps 0-7 compute, warp 8 = TMA producer of the (L_IJ, V_J) ring, warp 9 = MMA issuer + TMEM owner,
// warp 10 = TMA producer of the inv(L_II) ring.
#pragma once
#include "sweep_kernel.cuh"
#include "tc_common.cuh"

namespace bopy {

// element ownership of the compute warps = the tensor-memory fragment a warp may load: lanes (rows) 32 (warp % 4) ..,
// columns (candidates) 64 (warp / 4) ..
struct TcPolicy {
    using Elem = float;
    static constexpr int VEC = 4, KC = 8, CH = 16;          // 8 columns of L per operand tile, 16 tiles per 128 columns
    static constexpr int RI = 4, CJ = 16, CV = 2;
    static constexpr bool kSwizzled = false;
    int lane, rg, cg, part;
    bool leader;
    __device__ explicit TcPolicy(int tid) {
        lane = tid & 31;
        const int warp = tid >> 5;
        rg = warp & 3;
        cg = warp >> 2;
        part = rg;
        leader = (lane >> 2) == 0;
    }
    __device__ __forceinline__ int row_of(int i) const { return 32 * rg + 16 * (i >> 1) + 8 * (i & 1) + (lane >> 2); }
    __device__ __forceinline__ int cand_of(int j) const { return 64 * cg + 8 * (j >> 1) + 2 * (lane & 3) + (j & 1); }
    __device__ __forceinline__ double reduce_rows(double v) const {
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        return v;
    }
    // float offset of element (k = row of the 128-row block, c = candidate) inside a block row of the V workspace:
    // 16 tiles of 8 KB, each [hi 4 KB | lo 4 KB] in the operand layout; `lo` adds TF32_TILE_FLOATS
    __host__ __device__ static __forceinline__ int v_index(int k, int c) { return (k >> 3) * 2 * tc::TF32_TILE_FLOATS + tc::tile_index(k & 7, c); }
};

struct TcTag {};   // Engine<TcPolicy, DmmaPolicy>: same tile counts per block row as the fp64 engine (CHG = CHD = 16)
using EngineTc = Engine<TcPolicy, DmmaPolicy>;

// V[k][c] of one workspace slot as a double, whatever the engine stored
template <class E> __device__ __forceinline__ double v_value(const typename E::TG* Vslot, int k, int c) {
    if constexpr (std::is_same<typename E::PG, TcPolicy>::value) {
        const float* blk = Vslot + (long long)(k / BM) * BM * BN * 2;
        const int o = TcPolicy::v_index(k % BM, c);
        return static_cast<double>(blk[o]) + static_cast<double>(blk[o + tc::TF32_TILE_FLOATS]);
    } else {
        return static_cast<double>(Vslot[E::PG::b_index(k, c)]);
    }
}
template <class E> __host__ __device__ constexpr int v_elems_per_entry() { return std::is_same<typename E::PG, TcPolicy>::value ? 2 : 1; }

constexpr int TC_PERM_STAGES = 4;     // (L_IJ, V_J) ring slots that are always there: 8 KB + 8 KB each
constexpr int TC_BORROWED = 8;        // + the 128 KB of the fp64 residual tile, lent to the ring while a block row's GEMM runs
constexpr int TC_DSTAGES = 2;         // inv(L_II) ring: 8 KB per stage
constexpr int TC_FOLD_TILES = 4;      // default: operand tiles (x 8 columns of L) per hi.hi accumulation chain in tensor memory
constexpr int TC_NT_ALL = NT + 128;
constexpr int TC_TAIL_BYTES = 512;    // mbarriers, tensor-memory address, arg-min scratch

// shared memory: [perm ring][Rs = borrowed ring slots][inv(L_II) ring][candidates / l][partial sums 4 KB][tail][X/l block row]
__host__ __device__ constexpr size_t tc_smem_bytes(int d, int perm, int dstages) {
    return (size_t)perm * 2 * TILE_BYTES + (size_t)BM * BN * sizeof(double) + (size_t)dstages * TILE_BYTES +
           (size_t)d * BN * sizeof(double) + 4096 + TC_TAIL_BYTES + (size_t)(d + 1) * BM * sizeof(double);
}
// the deepest rings that fit next to the d-dependent buffers
inline void tc_ring_depths(int d, int* perm, int* dstages) {
    const int opts[5][2] = {{4, 2}, {3, 2}, {2, 2}, {2, 1}, {1, 1}};
    for (const auto& o : opts)
        if (tc_smem_bytes(d, o[0], o[1]) <= SMEM_LIMIT) {
            *perm = o[0];
            *dstages = o[1];
            return;
        }
    *perm = 1;
    *dstages = 1;
}

template <int KIND, int FOLD>
__global__ void __launch_bounds__(TC_NT_ALL, 1) sweep_tc_kernel(const SweepParams p) {
    using PG = TcPolicy;
    using PD = DmmaPolicy;
    using E = EngineTc;
    constexpr int CHG = E::CHG, CHD = E::CHD;
    constexpr int CHUNKS_PER_BLOCK = CHG / FOLD;      // FOLD = operand tiles per hi.hi chain
    constexpr int MAX_SLOTS = TC_PERM_STAGES + TC_BORROWED;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t NP = (uint32_t)p.tc_stages;                                     // permanent ring slots
    const uint32_t NSLOT = NP + TC_BORROWED;                                       // + the slots borrowed from Rs
    const uint32_t DS = (uint32_t)p.tc_dstages;                                    // depth of the inv(L_II) ring
    unsigned char* const ring = smem_raw;                                          // slot s at ring + s * 16 KB: [A 8 KB | B 8 KB]
    unsigned char* const rs_raw = smem_raw + (size_t)NP * 2 * TILE_BYTES;          // = ring slot NP: the borrowed slots start here
    double* const Rs = reinterpret_cast<double*>(rs_raw);                          // [BM][BN] fp64 residual (B operand of the diagonal GEMM)
    unsigned char* const dring = rs_raw + (size_t)BM * BN * sizeof(double);        // [DS] inv(L_II) tiles
    double* const xs_s = reinterpret_cast<double*>(dring + (size_t)DS * TILE_BYTES);   // [d][BN] candidates / l
    double* const part = xs_s + (size_t)p.d * BN;                                  // [4][BN] partial sums (mean, then sum v^2)
    unsigned char* const tail = reinterpret_cast<unsigned char*>(part + 4 * BN);
    uint64_t* const full = reinterpret_cast<uint64_t*>(tail);      // [MAX_SLOTS] producer -> MMA issuer
    uint64_t* const empty = full + MAX_SLOTS;                      // [MAX_SLOTS] tcgen05.commit -> producer
    uint64_t* const dfull = empty + MAX_SLOTS;                     // [<= 4]
    uint64_t* const dempty = dfull + 4;                            // [<= 4]
    uint64_t* const accfull = dempty + 4;                          // [2] hi.hi chain finished -> compute warps
    uint64_t* const accempty = accfull + 2;                        // [2] compute warps read it -> MMA issuer
    uint64_t* const lofull = accempty + 2;                         // [2] cross-term accumulator of a block row finished
    uint64_t* const loempty = lofull + 2;                          // [2]
    uint64_t* const xbar = loempty + 2;
    uint64_t* const vbar = xbar + 1;
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(vbar + 1);
    MinLoc* const red = reinterpret_cast<MinLoc*>(tail + 384);     // [4]
    double* const xrow = reinterpret_cast<double*>(tail + TC_TAIL_BYTES);          // [(d+1)][BM] X/l block row + alpha

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_pad = p.n_blocks * BM;
    const unsigned char* const Lt = reinterpret_cast<const unsigned char*>(p.Lt);
    const long long slot_bytes = (long long)n_pad * BN * 2 * sizeof(float);

    if (tid == 0) {
        for (uint32_t s = 0; s < NSLOT; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (uint32_t s = 0; s < DS; ++s) {
            mbar_init(&dfull[s], 1);
            mbar_init(&dempty[s], NT / 32);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&accfull[s], 1);
            mbar_init(&accempty[s], NT / 32);
            mbar_init(&lofull[s], 1);
            mbar_init(&loempty[s], NT / 32);
        }
        mbar_init(xbar, 1);
        mbar_init(vbar, 1);
        fence_mbar_init();
    }
    if (warp == NT / 32 + 1) tc::tmem_alloc(tmem_slot, 512);   // the whole tensor memory of the SM: one CTA per SM
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = *tmem_slot;
    // tensor-memory columns: [0,128) and [128,256) hi.hi ping-pong, [256,384) and [384,512) cross terms by block-row parity

    // Ring discipline.  Stage t of a block row's GEMM uses slot t mod NSLOT, every block row starting again at slot 0.
    // Slots 0 .. NP-1 are permanent; slots NP .. NSLOT-1 ARE the residual tile Rs: the compute warps own that memory from
    // the end of the row's GEMM (all stages consumed by then) to the end of its diagonal solve (vbar), the ring owns it in
    // between.  So the producer may fill the first NP stages of the next row at any time, and waits for vbar before the
    // first borrowed slot.  Each slot's mbarrier parity is tracked in a bit mask (a slot is used a varying number of times
    // per row).
    if (warp >= NT / 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
        if (warp == NT / 32) {
            // =============================== TMA producer of the (L_IJ, V_J) ring =================================
            // (the whole warp runs the loop so that counters and addresses stay warp-uniform; one elected lane issues)
            uint32_t emask = 0, vphase = 0;
            // operand-tile index (within the packed block row / the V slot) of stage t of block row I: zig-zag over the
            // older V slices (the ones read last by the previous row are still in L2), V_{I-1} last
            auto tile_of = [](int I, int t) -> long long {
                const int jpos = t / CHG, c = t - jpos * CHG;
                const int J = (I & 1) ? (jpos < I - 1 ? I - 2 - jpos : I - 1) : jpos;
                return (long long)J * CHG + c;
            };
            for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const unsigned char* const Vt =
                    reinterpret_cast<const unsigned char*>(p.Vws) + (p.slot_per_tile ? tile : (long long)blockIdx.x) * slot_bytes;
                for (int I = 1; I < p.n_blocks; ++I) {
                    const unsigned char* const a_row = Lt + E::row_base(I) * TILE_BYTES;
                    const int T_gemm = I * CHG;
                    // V_{I-1} is published, and Rs released, when the compute warps arrive on vbar: needed before the first
                    // borrowed slot and before the first tile of V_{I-1}, whichever comes first
                    const int t_vbar = min((int)NP, T_gemm - CHG);
                    uint32_t slot = 0;
                    for (int t = 0; t < T_gemm; ++t) {
                        if (t == t_vbar) {
                            mbar_wait(vbar, vphase);
                            vphase ^= 1u;
                        }
                        mbar_wait(&empty[slot], ((emask >> slot) & 1u) ^ 1u);
                        emask ^= 1u << slot;
                        const long long tt = tile_of(I, t);
                        unsigned char* const st = ring + (size_t)slot * 2 * TILE_BYTES;
                        if (tc::elect_one()) {
                            mbar_arrive_expect_tx(&full[slot], 2 * TILE_BYTES);
                            bulk_g2s(st, a_row + tt * TILE_BYTES, TILE_BYTES, &full[slot]);
                            bulk_g2s(st + TILE_BYTES, Vt + tt * TILE_BYTES, TILE_BYTES, &full[slot]);
                        }
                        __syncwarp();
                        if (++slot == NSLOT) slot = 0;
                    }
                }
            }
        } else if (warp == NT / 32 + 2 && lane == 0) {
            // =============================== TMA producer of the inv(L_II) ring ===================================
            uint32_t stage = 0, parity = 1;
            for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x)
                for (int I = 0; I < p.n_blocks; ++I) {
                    const unsigned char* const d_row = Lt + (E::row_base(I) + (long long)I * CHG) * TILE_BYTES;
                    for (int t = 0; t < CHD; ++t) {
                        mbar_wait(&dempty[stage], parity);
                        mbar_arrive_expect_tx(&dfull[stage], TILE_BYTES);
                        bulk_g2s(dring + (size_t)stage * TILE_BYTES, d_row + (long long)t * TILE_BYTES, TILE_BYTES, &dfull[stage]);
                        if (++stage == DS) {
                            stage = 0;
                            parity ^= 1u;
                        }
                    }
                }
        } else if (warp == NT / 32 + 1) {
            // =============================== MMA issuer + tensor-memory owner =====================================
            {   // the whole warp runs the loop (warp-uniform counters); one elected lane issues
                constexpr uint32_t idesc = tc::idesc_tf32(BM, BN, false, false);
                // descriptor of the A_hi tile of slot 0; the other tiles / slots differ in the address field only
                // (16-byte units: +256 per 4 KB tile, +1024 per slot)
                const uint64_t desc0 = tc::smem_desc(smem_u32(ring), tc::TILE_LBO, tc::TILE_SBO);
                uint32_t fmask = 0, chunk = 0, rows = 0;
                long long w_full = 0, w_acc = 0, w_lo = 0, t_all0 = clock64();
                const bool prof = p.tc_prof != nullptr;
                for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x)
                    for (int I = 1; I < p.n_blocks; ++I, ++rows) {
                        const uint32_t lb = rows & 1u;
                        long long c0_ = prof ? clock64() : 0;
                        mbar_wait(&loempty[lb], ((rows >> 1) & 1u) ^ 1u);
                        if (prof) w_lo += clock64() - c0_;
                        const uint32_t d_lo = tmem + 256u + 128u * lb;
                        const int nchunks = I * CHUNKS_PER_BLOCK;
                        uint32_t slot = 0;
                        for (int c = 0; c < nchunks; ++c, ++chunk) {
                            const uint32_t b = chunk & 1u;
                            long long c1_ = prof ? clock64() : 0;
                            mbar_wait(&accempty[b], ((chunk >> 1) & 1u) ^ 1u);
                            if (prof) w_acc += clock64() - c1_;
                            const uint32_t d_hi = tmem + 128u * b;
#pragma unroll
                            for (int f = 0; f < FOLD; ++f) {
                                long long c2_ = prof ? clock64() : 0;
                                mbar_wait(&full[slot], (fmask >> slot) & 1u);
                                if (prof) w_full += clock64() - c2_;
                                fmask ^= 1u << slot;
                                tc::fence_after_thread_sync();
                                if (tc::elect_one()) {
                                    const uint64_t a_hi = desc0 + (uint64_t)(slot * 1024u);
                                    const uint64_t a_lo = a_hi + 256u, b_hi = a_hi + 512u, b_lo = a_hi + 768u;
                                    tc::mma_tf32(d_lo, a_lo, b_hi, idesc, (c | f) != 0);
                                    tc::mma_tf32(d_lo, a_hi, b_lo, idesc, 1);
                                    tc::mma_tf32(d_hi, a_hi, b_hi, idesc, f != 0);
                                    tc::mma_commit(&empty[slot]);                  // the slot is free once these MMAs have read it
                                    if (f == FOLD - 1) tc::mma_commit(&accfull[b]);   // partial sum complete
                                }
                                __syncwarp();
                                if (++slot == NSLOT) slot = 0;
                            }
                        }
                        if (tc::elect_one()) tc::mma_commit(&lofull[lb]);
                        __syncwarp();
                    }
                if (prof && lane == 0) {
                    long long* o = p.tc_prof + (size_t)blockIdx.x * 16;
                    o[8] = w_full;
                    o[9] = w_acc;
                    o[10] = w_lo;
                    o[11] = clock64() - t_all0;
                }
            }
            __syncwarp();
            asm volatile("bar.sync 2, %0;" ::"n"(NT + 32) : "memory");   // every compute warp has read its last fragment
            tc::fence_after_thread_sync();
            tc::tmem_dealloc(tmem, 512);
        }
        return;
    }

    // ===================================== compute warps ==================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_COMPUTE));
    const PG pg(tid);
    const PD pd(tid);
    uint32_t dstage = 0, dparity = 0;   // position in the inv(L_II) ring
    uint32_t chunk = 0;    // hi.hi partial sums consumed so far
    uint32_t rows = 0;     // block rows with off-diagonal work finished so far
    uint32_t xphase = 0;
    MinLoc best;
    best.val = 0.0;
    best.idx = -1;
    const bool prof = p.tc_prof != nullptr && tid == 0;
    long long pk = 0, pgw = 0, pgf = 0, pdg = 0, ppub = 0, pstamp = 0, pall = prof ? clock64() : 0;
#define TC_PROF_MARK(acc_)                   \
    if (prof) {                              \
        const long long now_ = clock64();    \
        acc_ += now_ - pstamp;               \
        pstamp = now_;                       \
    }

    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const long long c0 = tile * BN;
        float* const Vt = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(p.Vws) +
                                                   (p.slot_per_tile ? tile : (long long)blockIdx.x) * slot_bytes);

        auto issue_xrow = [&](int I) {
            const uint32_t bytes = (uint32_t)(p.d + 1) * BM * sizeof(double);
            mbar_arrive_expect_tx(xbar, bytes);
            bulk_g2s(xrow, p.Xt + (long long)I * (p.d + 1) * BM, bytes, xbar);
        };

        if (tid == 0 && tile == blockIdx.x) issue_xrow(0);   // (later tiles: prefetched by the previous tile's last row)
        for (int e = tid; e < BN * p.d; e += NT) {
            const int c = e / p.d, q = e - c * p.d;
            const long long gc = c0 + c;
            const double v = gc < p.m ? p.Xs[gc * p.d + q] : 0.0;
            xs_s[q * BN + c] = __ddiv_rn(v, p.ls[q]);
        }
        double mean_c = 0.0, ss_c = 0.0;
        consumer_sync();

        for (int I = 0; I < p.n_blocks; ++I) {
            if (prof) pstamp = clock64();
            // ---- S = sum_J (-L_IJ) V_J from the tensor core: short hi.hi chains + the row's cross terms, summed in fp32
            //      (round to nearest) in registers -----------------------------------------------------------------------
            float acc[PG::RI][PG::CJ];
#pragma unroll
            for (int i = 0; i < PG::RI; ++i)
#pragma unroll
                for (int j = 0; j < PG::CJ; ++j) acc[i][j] = 0.f;
            if (I > 0) {
                auto add_fragment = [&](uint32_t col) {   // both 16-row halves are requested before the one wait
                    uint32_t r0[32], r1[32];
                    tc::tmem_ld_16x256b_x8(tmem + ((uint32_t)(32 * pg.rg) << 16) + col + 64u * pg.cg, r0);
                    tc::tmem_ld_16x256b_x8(tmem + ((uint32_t)(32 * pg.rg + 16) << 16) + col + 64u * pg.cg, r1);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                acc[h][2 * nt + e] += __uint_as_float(r0[4 * nt + 2 * h + e]);
                                acc[2 + h][2 * nt + e] += __uint_as_float(r1[4 * nt + 2 * h + e]);
                            }
                };
                const int nchunks = I * CHUNKS_PER_BLOCK;
                for (int c = 0; c < nchunks; ++c, ++chunk) {
                    const uint32_t b = chunk & 1u;
                    mbar_wait(&accfull[b], (chunk >> 1) & 1u);
                    TC_PROF_MARK(pgw)
                    tc::fence_after_thread_sync();
                    add_fragment(128u * b);
                    tc::fence_before_thread_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&accempty[b]);
                    TC_PROF_MARK(pgf)
                }
                {   // the cross terms of the whole block row
                    const uint32_t lb = rows & 1u;
                    mbar_wait(&lofull[lb], (rows >> 1) & 1u);
                    tc::fence_after_thread_sync();
                    add_fragment(256u + 128u * lb);
                    tc::fence_before_thread_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&loempty[lb]);
                    ++rows;
                }
                TC_PROF_MARK(pgf)
                // every MMA of this row has completed (the last chain's commit covers them all): the borrowed ring slots
                // are drained and Rs belongs to the compute warps again
            }

            // ---- kernel tile K*[block row I, this tile's candidates] (fp64), its share of the mean, and the residual
            //      R_I = K*_I + S, written straight into Rs (the diagonal policy's B layout).  (Working the kernel tile off
            //      while the warps wait for the tensor core was tried -- K* parked in an L2 scratch: its shared-memory loads
            //      then compete with the MMAs' operand reads and the tile takes twice as long; net loss.) ---------------------
            mbar_wait(xbar, xphase);
            xphase ^= 1;
            {
                double mp[PG::CJ];
#pragma unroll
                for (int j = 0; j < PG::CJ; ++j) mp[j] = 0.0;
#pragma unroll
                for (int i = 0; i < PG::RI; ++i) {
                    const int row = pg.row_of(i);
                    const double amp_i = (I * BM + row < p.n) ? p.amp : 0.0;   // rows beyond n (identity padding): amplitude 0, no branch
                    const double a_i = xrow[p.d * BM + row];
                    double d2[PG::CJ];
#pragma unroll
                    for (int j = 0; j < PG::CJ; ++j) d2[j] = 0.0;
                    for (int q = 0; q < p.d; ++q) {
                        const double xr = xrow[q * BM + row];
#pragma unroll
                        for (int j = 0; j < PG::CJ; ++j) {
                            const double df = xs_s[q * BN + pg.cand_of(j)] - xr;
                            d2[j] = fma(df, df, d2[j]);   // cdist's summation order over the dimensions
                        }
                    }
#pragma unroll
                    for (int jv = 0; jv < PG::CJ / 2; ++jv) {
                        const int j = 2 * jv;
                        const double k0 = __dmul_rn(amp_i, base_kernel<KIND>(d2[j]));
                        const double k1 = __dmul_rn(amp_i, base_kernel<KIND>(d2[j + 1]));
                        mp[j] = fma(k0, a_i, mp[j]);
                        mp[j + 1] = fma(k1, a_i, mp[j + 1]);
                        *reinterpret_cast<double2*>(&Rs[PD::b_index(row, pg.cand_of(j))]) =
                            make_double2(k0 + static_cast<double>(acc[i][j]), k1 + static_cast<double>(acc[i][j + 1]));
                    }
                }
#pragma unroll
                for (int j = 0; j < PG::CJ; ++j) mp[j] = pg.reduce_rows(mp[j]);
                if (pg.leader) {
#pragma unroll
                    for (int j = 0; j < PG::CJ; ++j) part[pg.part * BN + pg.cand_of(j)] = mp[j];
                }
            }
            consumer_sync();   // Rs complete, X/l block row consumed
            if (tid < BN) mean_c += ((part[tid] + part[BN + tid]) + part[2 * BN + tid]) + part[3 * BN + tid];
            if (tid == 0) {   // prefetch the next block row (or the next tile's first) of X/l
                if (I + 1 < p.n_blocks) issue_xrow(I + 1);
                else if (tile + gridDim.x < p.ntiles) issue_xrow(0);
            }
            TC_PROF_MARK(pk)

            // ---- V_I = inv(L_II) R_I (fp64 DMMA, operands from the inv(L_II) ring) -----------------------------------
            double accd[PD::RI][PD::CJ];
#pragma unroll
            for (int i = 0; i < PD::RI; ++i)
#pragma unroll
                for (int j = 0; j < PD::CJ; ++j) accd[i][j] = 0.0;
            for (int kc = 0; kc < CHD; ++kc) {
                mbar_wait(&dfull[dstage], dparity);
                pd.template mma_tile<true>(accd, reinterpret_cast<const double*>(dring + (size_t)dstage * TILE_BYTES),
                                           Rs + kc * PD::KC * BN, kc);
                __syncwarp();
                if (lane == 0) mbar_arrive(&dempty[dstage]);
                if (++dstage == DS) {
                    dstage = 0;
                    dparity ^= 1u;
                }
            }
            TC_PROF_MARK(pdg)

            // ---- V_I: publish as TF32 pairs in the operand layout, fold into sum v^2 ---------------------------------
            consumer_sync();   // the mean's partial sums have been read: `part` may take the sum v^2 shares
            {
                const bool publish = (I + 1 < p.n_blocks) || p.slot_per_tile;
                float* const Vrow = Vt + (long long)I * BM * BN * 2;
                double sq[PD::CJ];
#pragma unroll
                for (int j = 0; j < PD::CJ; ++j) sq[j] = 0.0;
#pragma unroll
                for (int i = 0; i < PD::RI; ++i) {
                    const int row = pd.row_of(i);
#pragma unroll
                    for (int j = 0; j < PD::CJ; ++j) {
                        const double v = accd[i][j];
                        if (publish) {
                            float hi, lo;
                            tc::tf32_pair(v, hi, lo);
                            const int o = PG::v_index(row, pd.cand_of(j));
                            Vrow[o] = hi;
                            Vrow[o + tc::TF32_TILE_FLOATS] = lo;
                        }
                        sq[j] = fma(v, v, sq[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < PD::CJ; ++j) sq[j] = pd.reduce_rows(sq[j]);
                if (pd.leader) {
#pragma unroll
                    for (int j = 0; j < PD::CJ; ++j) part[pd.part * BN + pd.cand_of(j)] = sq[j];
                }
                fence_proxy_async();  // V stores and the last generic accesses of Rs, before the producer's bulk copies
            }
            consumer_sync();   // every warp finished the diagonal GEMM (Rs reads) and published its part of V_I
            if (tid == 0 && I + 1 < p.n_blocks) mbar_arrive(vbar);   // V_I may be loaded; Rs may be lent to the ring again
            if (tid < BN) ss_c += ((part[tid] + part[BN + tid]) + part[2 * BN + tid]) + part[3 * BN + tid];
            TC_PROF_MARK(ppub)
            // (the next write to `part` is the next row's mean share, behind that row's GEMM folds or, for the next
            //  tile's row 0, behind the epilogue's barriers)
        }

        // ---- epilogue: de-normalise, acquisition, arg-min (identical to sweep_kernel) ----------------------------------
        MinLoc mine;
        mine.val = 0.0;
        mine.idx = -1;
        if (tid < BN) {
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
        consumer_sync();  // xs_s, part and red are reused by the next tile
    }
    if (tid == 0 && p.partials != nullptr) p.partials[blockIdx.x] = best;
    if (prof) {
        long long* o = p.tc_prof + (size_t)blockIdx.x * 16;
        o[0] = pk;
        o[1] = pgw;
        o[2] = pgf;
        o[4] = pdg;
        o[5] = ppub;
        o[6] = clock64() - pall;
    }
#undef TC_PROF_MARK
    tc::fence_before_thread_sync();
    asm volatile("bar.arrive 2, %0;" ::"n"(NT + 32) : "memory");
}

}  // namespace bopy
