// C ABI of the B200-native bopy hot path (see include/bopy_b200.h for the contract).
#include "bopy_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "common.cuh"
#include "fit_kernels.cuh"
#include "grad_kernel.cuh"
#include "minloc_comm.cuh"
#include "probe_inv_kernel.cuh"
#include "probe_kernel.cuh"
#include "prune_kernels.cuh"
#include "small_kernel.cuh"
#include "sweep_group_kernel.cuh"
#include "sweep_kernel.cuh"
#include "sweep_tc_kernel.cuh"
#include "sweep_warp_kernel.cuh"

using namespace bopy;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t err__ = (expr);                                                                \
        if (err__ != cudaSuccess) {                                                                \
            cudaGetLastError(); /* reported here: do not leave it for the next call to trip over */ \
            return fail(BOPY_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__),  \
                        __FILE__, __LINE__);                                                       \
        }                                                                                          \
    } while (0)


}  // namespace

struct bopy_gp {
    int device = 0, dtype = BOPY_F64, kernel = BOPY_KERNEL_RBF;
    long long n = 0;
    int d = 0, n_blocks = 0, n_pad = 0, sm_count = 0;
    void* Lt = nullptr;        // packed factor tiles
    double* Xt = nullptr;      // [n_blocks][d+1][BM]
    double* Dinv = nullptr;    // [n_blocks][BM][BM] fp64 inverted diagonal blocks
    void* Vws = nullptr;       // [sm_count][n_pad][BN]
    MinLoc* partials = nullptr;
    double ls[MAX_D];
    double amp = 1.0, noise = 0.0, y_mean = 0.0, y_std = 1.0;
    bool ready = false;
    bool fma64 = false;        // fp64 solve with the register-tiled FMA engine instead of DMMA (BOPY_B200_F64_ENGINE=fma)
    int f32_engine = 0;        // fp32 solve: 0 = tcgen05.mma kind::tf32 + TMEM (default), 1 = warp-level mma.sync 3xTF32
                               // (BOPY_B200_F32_ENGINE=mma_sync), 2 = register-tiled FFMA (BOPY_B200_F32_ENGINE=fma)
    bool small_n = true;       // n <= 32 on fp64 handles: thread-per-candidate kernel (BOPY_B200_SMALL_N=0 switches it off)
    int nan_skip = 0;          // arg-min policy for NaN acquisition values: 0 = np.argmin (first NaN wins), 1 = np.nanargmin
    // latency path (probe_kernel): used for m <= probe_max_m on fp64 handles whose block rows fit one wave of CTAs
    bool probe_capable = false;
    long long probe_max_m = 0;
    int probe_max_batch = 0;
    MinLoc* probe_records = nullptr;   // [probe_max_batch]
    unsigned* probe_flags = nullptr;   // the role ticket, then [2][probe_max_batch][n_blocks]: V_I published / partials published
    void* Mt = nullptr;                // [n_blocks][16] tiles of -inv(L_II) L_{I,I-1} (latency path's last hop)
    double* probe_part = nullptr;      // [probe_max_batch][n_blocks][2][PROBE_MAX_NC]
    unsigned probe_ticket_base = 0, probe_epoch = 0;
    long long* probe_trace = nullptr;  // [n_blocks][8] time stamps of the next latency-path launch (bopy_gp_probe_trace)
    // acquisition gradient (grad_kernel), allocated on first use: chunks of up to sm_count batches
    unsigned* grad_flags = nullptr;    // [2][sm_count][n_blocks]: W_I published / gradient shares published
    double* grad_part = nullptr;       // [sm_count][n_blocks][2][d][PROBE_MAX_NC]
    double* grad_mv = nullptr;         // [2][sm_count * PROBE_MAX_NC]: mean / var scratch of a chunk
    unsigned grad_epoch = 0;
    // the factor itself (row-major, leading dimension n_pad), kept by bopy_gp_fit so that bopy_gp_append can grow it
    double* Lfull = nullptr;
    bool Lfull_valid = false;
    double alpha_reg = 0.0;            // the jitter the kept factor was built with
    bool Lfull_factor = false;         // Lfull holds the CURRENT factor (kept by the fit, or copied by bopy_gp_set_state)
    // inverse path (probe_inv_kernel.cuh): W = L^-1 for calls of <= inv_max_m candidates on a state that is probed often
    int inv_mode = -1;                 // -1 auto (W is built at the inv_auto_calls()-th small call on a state), 0 off, 1 on
    long long inv_max_m = 0;           // 0: this handle has no inverse path
    double* Winv = nullptr;            // [n_pad][n_pad]
    bool Winv_valid = false;
    int inv_small_calls = 0;           // small calls since the state last changed
    double* inv_part = nullptr;        // [sm_count][INV_MAX_NC]
    unsigned* inv_ticket = nullptr;
    unsigned inv_ticket_base = 0;
    double* inv_host_out = nullptr;    // [3][INV_MAX_NC] mapped pinned host memory: acq / mean / var of a host-buffer call of <= 8 candidates
    // staging of the host-buffer entry point (small calls: one point per DIRECT probe)
    double* host_x = nullptr;          // [HOST_CALL_MAX_M][d]
    double* host_out = nullptr;        // [3][HOST_CALL_MAX_M]: acq / mean / var
    // group mode of the fp64 sweep (sweep_group_kernel.cuh): group_size CTAs share a candidate tile, so that the V
    // workspace in flight fits the L2; 1 = off.  Buffers are allocated on first use.
    int warp_stages = 0;               // n <= 256 on fp64 handles: ring depth of the warp-autonomous kernel (sweep_warp_kernel.cuh), 0 = off
    int group_size = 1, group_slots = 2, group_lead = 0;
    unsigned* gctl = nullptr;
    double* gpart = nullptr;
};

namespace {


template <class E, int KIND> int launch_sweep_t(SweepParams p, int grid, cudaStream_t st) {
    const size_t smem = sweep_smem_bytes<E>(p.d);
    p.xrow_separate = sweep_xrow_separate<E>(p.d) ? 1 : 0;
    CUDA_TRY(cudaFuncSetAttribute(sweep_kernel<E, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sweep_kernel<E, KIND><<<grid, NT_ALL, smem, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

template <int KIND> int launch_sweep_group_t(SweepParams p, int grid, cudaStream_t st) {
    using E = EngineF64;
    size_t smem = sweep_smem_bytes<E>(p.d);
    p.xrow_separate = sweep_xrow_separate<E>(p.d) ? 1 : 0;
    // staging buffer for a tile's raw candidate rows behind the X/l row buffer, when it fits and the rows are 16-byte aligned
    const size_t stage_bytes = (size_t)BN * p.d * sizeof(double);
    p.xs_stage = (p.xrow_separate && smem + stage_bytes <= SMEM_LIMIT && reinterpret_cast<uintptr_t>(p.Xs) % 16 == 0) ? 1 : 0;
    if (p.xs_stage) smem += stage_bytes;
    CUDA_TRY(cudaFuncSetAttribute(sweep_group_kernel<E, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sweep_group_kernel<E, KIND><<<grid, NT_ALL, smem, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

template <int KIND> int launch_sweep_warp_t(const SweepParams& p, int grid, int stages, cudaStream_t st) {
    const size_t smem = wk_smem_bytes(p.d, p.n_blocks, stages);
    CUDA_TRY(cudaFuncSetAttribute(sweep_warp_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sweep_warp_kernel<KIND><<<grid, WK_NT, smem, st>>>(p, stages);
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

int launch_sweep_warp(int kernel, const SweepParams& p, int grid, int stages, cudaStream_t st) {
    switch (kernel) {
        case BOPY_KERNEL_RBF: return launch_sweep_warp_t<K_RBF>(p, grid, stages, st);
        case BOPY_KERNEL_MATERN12: return launch_sweep_warp_t<K_M12>(p, grid, stages, st);
        case BOPY_KERNEL_MATERN32: return launch_sweep_warp_t<K_M32>(p, grid, stages, st);
        case BOPY_KERNEL_MATERN52: return launch_sweep_warp_t<K_M52>(p, grid, stages, st);
    }
    return fail(BOPY_ERR_BAD_ARG, "unknown kernel id %d", kernel);
}

int launch_sweep_group(int kernel, const SweepParams& p, int grid, cudaStream_t st) {
    switch (kernel) {
        case BOPY_KERNEL_RBF: return launch_sweep_group_t<K_RBF>(p, grid, st);
        case BOPY_KERNEL_MATERN12: return launch_sweep_group_t<K_M12>(p, grid, st);
        case BOPY_KERNEL_MATERN32: return launch_sweep_group_t<K_M32>(p, grid, st);
        case BOPY_KERNEL_MATERN52: return launch_sweep_group_t<K_M52>(p, grid, st);
    }
    return fail(BOPY_ERR_BAD_ARG, "unknown kernel id %d", kernel);
}

template <int KIND, int FOLD> int launch_sweep_tc_f(SweepParams p, int grid, cudaStream_t st);

template <int KIND> int launch_sweep_tc_t(SweepParams p, int grid, cudaStream_t st) {
    tc_ring_depths(p.d, &p.tc_stages, &p.tc_dstages);
    p.tc_fold = TC_FOLD_TILES;
    {   // development knobs (defaults: sweep_tc_kernel.cuh)
        const char* ns = std::getenv("BOPY_B200_TC_STAGES");
        if (ns && std::atoi(ns) >= 1 && std::atoi(ns) <= p.tc_stages) p.tc_stages = std::atoi(ns);
        const char* ds = std::getenv("BOPY_B200_TC_DSTAGES");
        if (ds && std::atoi(ds) >= 1 && std::atoi(ds) <= 4 && tc_smem_bytes(p.d, p.tc_stages, std::atoi(ds)) <= SMEM_LIMIT)
            p.tc_dstages = std::atoi(ds);
    }
    return launch_sweep_tc_f<KIND, TC_FOLD_TILES>(p, grid, st);
}

template <int KIND, int FOLD> int launch_sweep_tc_f(SweepParams p, int grid, cudaStream_t st) {
    p.xrow_separate = 1;
    const size_t smem = tc_smem_bytes(p.d, p.tc_stages, p.tc_dstages);
    CUDA_TRY(cudaFuncSetAttribute(sweep_tc_kernel<KIND, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool prof = std::getenv("BOPY_B200_TC_PROF") != nullptr;   // development aid: per-phase cycle counters to stderr
    if (prof) {
        CUDA_TRY(cudaMalloc(&p.tc_prof, (size_t)grid * 16 * sizeof(long long)));
        CUDA_TRY(cudaMemset(p.tc_prof, 0, (size_t)grid * 16 * sizeof(long long)));
    }
    sweep_tc_kernel<KIND, FOLD><<<grid, TC_NT_ALL, smem, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    if (prof) {
        std::vector<long long> h((size_t)grid * 16);
        CUDA_TRY(cudaMemcpy(h.data(), p.tc_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(p.tc_prof);
        double a[16] = {0};
        for (int b = 0; b < grid; ++b)
            for (int k = 0; k < 16; ++k) a[k] += (double)h[(size_t)b * 16 + k] / grid;
        std::fprintf(stderr,
                     "[tc_prof] cycles per CTA: compute tid0: K*+R %.0f | wait accfull %.0f | fold %.0f | diag %.0f | publish %.0f | "
                     "total %.0f ;  MMA warp: wait full %.0f | wait accempty %.0f | wait loempty %.0f | total %.0f  (tiles/CTA %.1f)\n",
                     a[0], a[1], a[2], a[4], a[5], a[6], a[8], a[9], a[10], a[11], (double)p.ntiles / grid);
    }
    return BOPY_OK;
}

template <class E> int launch_sweep_k(int kernel, const SweepParams& p, int grid, cudaStream_t st) {
    switch (kernel) {
        case BOPY_KERNEL_RBF: return launch_sweep_t<E, K_RBF>(p, grid, st);
        case BOPY_KERNEL_MATERN12: return launch_sweep_t<E, K_M12>(p, grid, st);
        case BOPY_KERNEL_MATERN32: return launch_sweep_t<E, K_M32>(p, grid, st);
        case BOPY_KERNEL_MATERN52: return launch_sweep_t<E, K_M52>(p, grid, st);
    }
    return fail(BOPY_ERR_BAD_ARG, "unknown kernel id %d", kernel);
}

template <> int launch_sweep_k<EngineTc>(int kernel, const SweepParams& p, int grid, cudaStream_t st) {
    switch (kernel) {
        case BOPY_KERNEL_RBF: return launch_sweep_tc_t<K_RBF>(p, grid, st);
        case BOPY_KERNEL_MATERN12: return launch_sweep_tc_t<K_M12>(p, grid, st);
        case BOPY_KERNEL_MATERN32: return launch_sweep_tc_t<K_M32>(p, grid, st);
        case BOPY_KERNEL_MATERN52: return launch_sweep_tc_t<K_M52>(p, grid, st);
    }
    return fail(BOPY_ERR_BAD_ARG, "unknown kernel id %d", kernel);
}

template <int NA, int KIND> int launch_probe_t(const ProbeParams& p, int grid, cudaStream_t st) {
    const size_t smem = probe_smem_bytes<NA>(p.d);
    CUDA_TRY(cudaFuncSetAttribute(probe_kernel<NA, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<NA, KIND><<<grid, PROBE_NT, smem, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

template <int NA> int launch_probe_k(int kernel, const ProbeParams& p, int grid, cudaStream_t st) {
    switch (kernel) {
        case BOPY_KERNEL_RBF: return launch_probe_t<NA, K_RBF>(p, grid, st);
        case BOPY_KERNEL_MATERN12: return launch_probe_t<NA, K_M12>(p, grid, st);
        case BOPY_KERNEL_MATERN32: return launch_probe_t<NA, K_M32>(p, grid, st);
        case BOPY_KERNEL_MATERN52: return launch_probe_t<NA, K_M52>(p, grid, st);
    }
    return fail(BOPY_ERR_BAD_ARG, "unknown kernel id %d", kernel);
}

// candidates per batch (8, 16 or 32) and the launch shape of the latency path for m candidates
struct ProbePlan {
    int na, nbatch, groups, grid;
};
ProbePlan probe_plan(const bopy_gp* gp, long long m) {
    ProbePlan pl;
    const int gmax = std::max(1, gp->sm_count / gp->n_blocks);
    pl.na = (m + 7) / 8 <= gmax ? 1 : ((m + 15) / 16 <= gmax ? 2 : 4);
    pl.nbatch = (int)((m + 8 * pl.na - 1) / (8 * pl.na));
    pl.groups = std::min(pl.nbatch, gmax);
    pl.grid = pl.groups * gp->n_blocks;
    return pl;
}

// most candidates one latency-path launch can hold: V lives in the sweep workspace, records / flags per batch
long long probe_capacity(const bopy_gp* gp) {
    return std::min<long long>((long long)gp->sm_count * BN, (long long)gp->probe_max_batch * PROBE_MAX_NC);
}

bool probe_applies(const bopy_gp* gp, long long m, int slot_per_tile, const MinLoc* tile_records) {
    return gp->probe_max_m > 0 && m <= gp->probe_max_m && slot_per_tile == 0 && tile_records == nullptr;
}

template <class P> int launch_cov_k(const bopy_gp* gp, const void* Vws, const double* Xs, long long m,
                                    const LsParam& ls, double* cov, cudaStream_t st) {   // P: an Engine
    dim3 block(16, 16), grid((unsigned)((m + 15) / 16), (unsigned)((m + 15) / 16));
    const double kss = gp->amp + gp->noise, yv = gp->y_std * gp->y_std;
    const typename P::TG* V = reinterpret_cast<const typename P::TG*>(Vws);   // (TG = float for every fp32 engine)
    switch (gp->kernel) {
        case BOPY_KERNEL_RBF:
            cov_kernel<P, K_RBF><<<grid, block, 0, st>>>(V, gp->n_pad, (int)gp->n, Xs, m, gp->d, ls, gp->amp, kss, yv, cov);
            break;
        case BOPY_KERNEL_MATERN12:
            cov_kernel<P, K_M12><<<grid, block, 0, st>>>(V, gp->n_pad, (int)gp->n, Xs, m, gp->d, ls, gp->amp, kss, yv, cov);
            break;
        case BOPY_KERNEL_MATERN32:
            cov_kernel<P, K_M32><<<grid, block, 0, st>>>(V, gp->n_pad, (int)gp->n, Xs, m, gp->d, ls, gp->amp, kss, yv, cov);
            break;
        default:
            cov_kernel<P, K_M52><<<grid, block, 0, st>>>(V, gp->n_pad, (int)gp->n, Xs, m, gp->d, ls, gp->amp, kss, yv, cov);
            break;
    }
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

// engine selection: f64 -> DMMA warp tiles (FMA thread tiles on request, for A/B runs); f32 -> mixed engine
template <class F> int dispatch_engine(const bopy_gp* gp, F&& f) {
    if (gp->dtype == BOPY_F32)
        return gp->f32_engine == 2 ? f(EngineMixedFma()) : (gp->f32_engine == 1 ? f(EngineMixed()) : f(EngineTc()));
    if (gp->fma64) return f(EngineF64Fma());
    return f(EngineF64());
}

// bytes of one V entry in the solve workspace: fp64, an fp32, or a TF32 (hi, lo) pair
size_t v_entry_bytes(const bopy_gp* gp) {
    if (gp->dtype == BOPY_F64) return sizeof(double);
    return gp->f32_engine == 0 ? 2 * sizeof(float) : sizeof(float);
}

long long packed_tiles(const bopy_gp* gp) {
    long long n = 0;
    dispatch_engine(gp, [&](auto e) { n = decltype(e)::total_tiles(gp->n_blocks); return 0; });
    return n;
}

// Group mode of the fp64 sweep (sweep_group_kernel.cuh): CTAs per candidate tile, workspace slots per group, rows of the next
// tile interleaved with the current one.  Requests < 0 mean "choose":
//   group size: 0 (one tile per CTA) when that kernel's workspace, sm_count x n_pad x 128 x 8 B, fits GROUP_L2_BUDGET anyway
//               or the handle has fewer than 4 block rows; else the smallest size whose tiles in flight (groups x slots x
//               n_pad x 128 x 8 B) fit the budget, at most 5/8 of the block rows (a tile's rows form a dependency chain);
//   slots 3, lead = half the block rows: two tiles of a group run half a tile apart, the third slot decouples them
//   (measured at C4: 8 CTAs x 3 slots x lead 8 = the speed of the one-tile-per-CTA kernel, 2 slots x lead 2 = -0.8 %).
// group size 1 = the one-tile-per-CTA kernel with the zig-zag V order of round 1 (A/B runs), 0 = the same kernel with
// group mode's ascending order (bit-identical to group mode).  BOPY_B200_SWEEP_GROUP / _SLOTS / _LEAD preset the requests.
constexpr size_t GROUP_L2_BUDGET = (size_t)96 << 20;   // of the 126 MB L2; the packed factor and the candidate stream live there too
void configure_group_mode(bopy_gp* gp, int G, int S, int lead) {
    const int R = gp->n_blocks, sm = gp->sm_count;
    const size_t slot_bytes = (size_t)gp->n_pad * BN * sizeof(double);
    const bool capable = R >= 4 && gp->dtype == BOPY_F64 && !gp->fma64;
    if (S < 0) S = 3;
    S = std::max(2, std::min(S, 4));
    if (G < 0) {
        if (!capable || (size_t)sm * slot_bytes <= GROUP_L2_BUDGET) {
            G = 0;
        } else {
            const int cap = std::max(2, 5 * R / 8);
            G = 2;
            while (G < cap && (size_t)((sm + G - 1) / G) * S * slot_bytes > GROUP_L2_BUDGET) ++G;
        }
    }
    G = std::max(0, std::min(G, sm));
    if (!capable && G > 1) G = 0;
    while (S > 2 && G > 1 && ((sm + G - 1) / G) * S > sm) --S;       // the slots live in the handle's [sm_count] workspace
    if (G > 1 && ((sm + G - 1) / G) * S > sm) G = 0;
    if (lead < 0) lead = S >= 3 ? R / 2 : std::min(2, R / 2);
    lead = std::max(0, std::min(lead, R / 2));
    if (S < 3) lead = std::min(lead, std::max(0, R / 2 - 1));        // a full interleave needs the third slot (sweep_group_kernel.cuh)
    gp->group_size = G;
    gp->group_slots = S;
    gp->group_lead = lead;
}

int env_int(const char* name, int fallback) {
    const char* v = std::getenv(name);
    return v != nullptr ? std::atoi(v) : fallback;
}

int check_ready(const bopy_gp* gp) {
    if (gp == nullptr) return fail(BOPY_ERR_BAD_ARG, "gp handle is NULL");
    if (!gp->ready) return fail(BOPY_ERR_NOT_READY, "bopy_gp_set_state has not been called on this handle");
    return BOPY_OK;
}

// n <= 32 on an fp64 handle (and no V export): the thread-per-candidate kernel serves every m
bool small_applies(const bopy_gp* gp, int slot_per_tile) {
    return gp->small_n && gp->dtype == BOPY_F64 && !gp->fma64 && gp->n <= SMALL_N_MAX && slot_per_tile == 0;
}

template <int NP> int launch_small_k(int kernel, const SmallParams& q, int grid, cudaStream_t st) {
    switch (kernel) {
        case BOPY_KERNEL_RBF: small_n_kernel<NP, K_RBF><<<grid, SMALL_NT, 0, st>>>(q); break;
        case BOPY_KERNEL_MATERN12: small_n_kernel<NP, K_M12><<<grid, SMALL_NT, 0, st>>>(q); break;
        case BOPY_KERNEL_MATERN32: small_n_kernel<NP, K_M32><<<grid, SMALL_NT, 0, st>>>(q); break;
        case BOPY_KERNEL_MATERN52: small_n_kernel<NP, K_M52><<<grid, SMALL_NT, 0, st>>>(q); break;
        default: return fail(BOPY_ERR_BAD_ARG, "unknown kernel id %d", kernel);
    }
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

int small_grid(const bopy_gp* gp, long long m) {
    return (int)std::min<long long>((m + SMALL_NT - 1) / SMALL_NT, (long long)gp->sm_count * SMALL_CTAS_PER_SM);
}

// one launch of the latency path (probe_kernel) over m candidates laid out by `pl`
int launch_probe(bopy_gp* gp, const ProbePlan& pl, const double* Xs, long long m, int acq, double eta, double kappa,
                 double* mean_out, double* var_out, double* acq_out, long long index_base, MinLoc* records, int keep_v,
                 cudaStream_t st, const double* rhs = nullptr) {
    ProbeParams q;
    std::memset(&q, 0, sizeof(q));
    q.Lt = reinterpret_cast<const unsigned char*>(gp->Lt);
    q.Xt = gp->Xt;
    q.V = reinterpret_cast<double*>(gp->Vws);
    q.Xs = Xs;
    q.m = m;
    q.nbatch = pl.nbatch;
    q.groups = pl.groups;
    q.n = (int)gp->n;
    q.n_blocks = gp->n_blocks;
    q.d = gp->d;
    for (int k = 0; k < gp->d; ++k) q.ls[k] = gp->ls[k];
    q.amp = gp->amp;
    q.kss = gp->amp + gp->noise;
    q.y_mean = gp->y_mean;
    q.y_std = gp->y_std;
    q.y_var = gp->y_std * gp->y_std;
    q.acq = acq;
    q.eta = eta;
    q.kappa = kappa;
    q.mean_out = mean_out;
    q.var_out = var_out;
    q.acq_out = acq_out;
    q.index_base = index_base;
    q.nan_skip = gp->nan_skip;
    q.records = records;
    q.Mt = reinterpret_cast<const unsigned char*>(gp->Mt);
    q.flags = gp->probe_flags + 1;
    q.flags2 = gp->probe_flags + 1 + (size_t)gp->probe_max_batch * gp->n_blocks;
    q.part = gp->probe_part;
    q.ticket = gp->probe_flags;
    q.ticket_base = gp->probe_ticket_base;
    if (++gp->probe_epoch == 0) {   // 2^32 launches: flags of batches not used for a whole cycle must not match again
        CUDA_TRY(cudaMemsetAsync(gp->probe_flags + 1, 0, (size_t)2 * gp->probe_max_batch * gp->n_blocks * sizeof(unsigned), st));
        gp->probe_epoch = 1;
    }
    q.epoch = gp->probe_epoch;
    q.keep_v = keep_v;
    q.rhs = rhs;
    q.trace = gp->probe_trace;
    int rc = pl.na == 1 ? launch_probe_k<1>(gp->kernel, q, pl.grid, st)
                        : (pl.na == 2 ? launch_probe_k<2>(gp->kernel, q, pl.grid, st)
                                      : launch_probe_k<4>(gp->kernel, q, pl.grid, st));
    if (rc == BOPY_OK) gp->probe_ticket_base += (unsigned)pl.grid;
    return rc;
}

// W = L^-1 (lower triangular, row-major, leading dimension ld = nb * 128; diagonal blocks = Dinv).  Default: the recursive
// 2 x 2 block inversion on tile_gemm_async_kernel (two launches per level, log2(nb) levels; the strict upper block triangle
// of W is its scratch).  BOPY_B200_TRTRI=diagonal: one block diagonal at a time on tile_gemm_kernel (round-2's first
// version, for A/B runs: 3.2 ms at n = 2048 and 48 ms at n = 8192 against 0.64 / 7.8 ms, profiles/r02/trtri_bench.log).
bool trtri_by_diagonals() {
    const char* v = std::getenv("BOPY_B200_TRTRI");
    return v != nullptr && std::strcmp(v, "diagonal") == 0;
}
int launch_trtri(const double* L, double* W, const double* Dinv, int nb, int ld, cudaStream_t st) {
    copy_diag_blocks_kernel<<<nb, 256, 0, st>>>(Dinv, W, ld);
    if (trtri_by_diagonals()) {
        for (int delta = 1; delta < nb; ++delta) {
            tile_gemm_kernel<<<nb - delta, NT, 0, st>>>(0, nb, delta, L, W, Dinv, nullptr, ld);
            tile_gemm_kernel<<<nb - delta, NT, 0, st>>>(1, nb, delta, L, W, Dinv, nullptr, ld);
        }
    } else {
        CUDA_TRY(cudaFuncSetAttribute(tile_gemm_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM_BYTES));
        for (int h = 1; h < nb; h *= 2) {
            const int grid = ((nb + 2 * h - 1) / (2 * h)) * h * h;
            tile_gemm_async_kernel<<<grid, NT, TG_SMEM_BYTES, st>>>(1, nb, h, L, W, nullptr, ld);
            tile_gemm_async_kernel<<<grid, NT, TG_SMEM_BYTES, st>>>(2, nb, h, L, W, nullptr, ld);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

// ---- inverse path (probe_inv_kernel.cuh): calls of a handful of candidates on a state that is probed often -----------
// auto mode: W = L^-1 is built at the 16th small call on one state.  The build costs what 6 chained calls cost at n = 2048
// (0.64 ms against 0.105 ms) and 24 at n = 8192 (7.9 ms against 0.33 ms; profiles/r02/trtri_bench.log), so a state that is
// probed less often than that never pays for it and one that is probed more often loses at most the price of a build
int inv_auto_calls(const bopy_gp*) { return 16; }
constexpr int INV_MAX_NPAD = 8192;                  // W and the kept factor are n_pad^2 doubles each (512 MB at 8192)
constexpr size_t INV_SMEM_BUDGET = (size_t)200 << 10;

// most candidates one inverse-path call can hold on this handle: K* of the call lives in shared memory ([nc][n_pad])
long long inv_capacity(const bopy_gp* gp) {
    // fp64 DMMA handles, and fp32 handles too: the path only needs the fp64 factor, inv(L_II), X / l and alpha_, which every
    // handle keeps in fp64 -- an fp32-mode surrogate probed by DIRECT gets fp64 answers in 0.03 ms instead of a 0.57 ms sweep
    const bool served = gp->probe_capable || (gp->dtype == BOPY_F32 && gp->n_blocks <= gp->sm_count);
    if (!served || gp->n_pad > INV_MAX_NPAD) return 0;
    if (gp->dtype == BOPY_F64 && gp->small_n && gp->n_pad == BM && gp->n <= SMALL_N_MAX) return 0;   // served by small_n_kernel
    for (int nc = INV_MAX_NC; nc >= 1; nc >>= 1)
        if (inv_smem_bytes(nc, gp->n_pad, gp->d) <= INV_SMEM_BUDGET) return nc;
    return 0;
}

bool inv_applies(const bopy_gp* gp, long long m, int slot_per_tile, const MinLoc* tile_records) {
    if (gp->inv_mode == 0 || m > gp->inv_max_m || !gp->Lfull_factor || slot_per_tile != 0 || tile_records != nullptr) return false;
    // fp64 handles: a sub-path of the latency path (bopy_gp_set_latency_path(gp, 0) switches both off)
    return gp->dtype == BOPY_F32 || probe_applies(gp, m, slot_per_tile, tile_records);
}

// every change of the fitted state: W no longer matches, the count of small calls starts again
void state_changed(bopy_gp* gp) {
    gp->Winv_valid = false;
    gp->inv_small_calls = 0;
}

// W = L^-1 by blocked triangular inversion of the kept factor (launch_trtri)
int build_linv(bopy_gp* gp, cudaStream_t st) {
    const int nb = gp->n_blocks, np = gp->n_pad;
    if (gp->inv_ticket == nullptr) {
        if (gp->Winv == nullptr) {
            CUDA_TRY(cudaMalloc(&gp->Winv, (size_t)np * np * sizeof(double)));
            CUDA_TRY(cudaMemsetAsync(gp->Winv, 0, (size_t)np * np * sizeof(double), st));   // the upper blocks stay 0
        }
        if (gp->inv_part == nullptr) CUDA_TRY(cudaMalloc(&gp->inv_part, (size_t)gp->sm_count * INV_MAX_NC * sizeof(double)));
        CUDA_TRY(cudaMalloc(&gp->inv_ticket, sizeof(unsigned)));
        CUDA_TRY(cudaMemsetAsync(gp->inv_ticket, 0, sizeof(unsigned), st));
        gp->inv_ticket_base = 0;
    }
    const int rc = launch_trtri(gp->Lfull, gp->Winv, gp->Dinv, nb, np, st);
    if (rc != BOPY_OK) return rc;
    gp->Winv_valid = true;
    return BOPY_OK;
}

// a call the inverse path could serve (inv_applies): is W there, or is it time to build it?  Counts the call otherwise.
int inv_prepare(bopy_gp* gp, cudaStream_t st, bool* use) {
    *use = gp->Winv_valid;
    if (!*use && (gp->inv_mode == 1 || ++gp->inv_small_calls >= inv_auto_calls(gp))) {
        const int rc = build_linv(gp, st);
        if (rc != BOPY_OK) return rc;
        *use = true;
    }
    return BOPY_OK;
}

template <int NC, int KIND> int launch_inv_t(const InvParams& q, int grid, size_t smem, cudaStream_t st) {
    CUDA_TRY(cudaFuncSetAttribute(probe_inv_kernel<NC, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_inv_kernel<NC, KIND><<<grid, INV_NT, smem, st>>>(q);
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

template <int NC> int launch_inv_k(int kernel, const InvParams& q, int grid, size_t smem, cudaStream_t st) {
    switch (kernel) {
        case BOPY_KERNEL_RBF: return launch_inv_t<NC, K_RBF>(q, grid, smem, st);
        case BOPY_KERNEL_MATERN12: return launch_inv_t<NC, K_M12>(q, grid, smem, st);
        case BOPY_KERNEL_MATERN32: return launch_inv_t<NC, K_M32>(q, grid, smem, st);
        case BOPY_KERNEL_MATERN52: return launch_inv_t<NC, K_M52>(q, grid, smem, st);
    }
    return fail(BOPY_ERR_BAD_ARG, "unknown kernel id %d", kernel);
}

int inv_grid(const bopy_gp* gp) {
    return (int)std::min<long long>(gp->sm_count, (gp->n + INV_WARPS - 1) / INV_WARPS);
}

// one launch of the inverse path over m <= inv_max_m candidates (W must be valid)
int launch_inv(bopy_gp* gp, const double* Xs, long long m, int acq, double eta, double kappa, double* mean_out,
               double* var_out, double* acq_out, long long index_base, double* min_val, long long* min_idx,
               cudaStream_t st, const double* xs_inline_host = nullptr) {
    InvParams q;
    std::memset(&q, 0, sizeof(q));
    q.W = gp->Winv;
    q.Xt = gp->Xt;
    q.Xs = Xs;
    if (xs_inline_host != nullptr) {   // host-buffer entry: the candidates are kernel parameters
        q.Xs = nullptr;
        std::memcpy(q.xs_inline, xs_inline_host, (size_t)m * gp->d * sizeof(double));
    }
    q.m = (int)m;
    q.n = (int)gp->n;
    q.n_blocks = gp->n_blocks;
    q.d = gp->d;
    for (int k = 0; k < gp->d; ++k) q.ls[k] = gp->ls[k];
    q.amp = gp->amp;
    q.kss = gp->amp + gp->noise;
    q.y_mean = gp->y_mean;
    q.y_std = gp->y_std;
    q.y_var = gp->y_std * gp->y_std;
    q.acq = acq;
    q.eta = eta;
    q.kappa = kappa;
    q.mean_out = mean_out;
    q.var_out = var_out;
    q.acq_out = acq_out;
    q.index_base = index_base;
    q.nan_skip = gp->nan_skip;
    q.min_val = min_val;
    q.min_idx = min_idx;
    q.part = gp->inv_part;
    q.ticket = gp->inv_ticket;
    q.ticket_base = gp->inv_ticket_base;
    const int grid = inv_grid(gp);
    const int nc = m <= 1 ? 1 : (m <= 2 ? 2 : (m <= 4 ? 4 : 8));
    const size_t smem = inv_smem_bytes(nc, gp->n_pad, gp->d);
    const int rc = nc == 1 ? launch_inv_k<1>(gp->kernel, q, grid, smem, st)
                           : (nc == 2 ? launch_inv_k<2>(gp->kernel, q, grid, smem, st)
                                      : (nc == 4 ? launch_inv_k<4>(gp->kernel, q, grid, smem, st)
                                                 : launch_inv_k<8>(gp->kernel, q, grid, smem, st)));
    if (rc == BOPY_OK) gp->inv_ticket_base += (unsigned)grid;
    return rc;
}

template <int NA, int KIND> int launch_grad_t(const GradParams& p, int grid, cudaStream_t st) {
    const size_t smem = grad_smem_bytes<NA>(p.d);
    CUDA_TRY(cudaFuncSetAttribute(grad_kernel<NA, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    grad_kernel<NA, KIND><<<grid, PROBE_NT, smem, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

template <int NA> int launch_grad_k(int kernel, const GradParams& p, int grid, cudaStream_t st) {
    switch (kernel) {
        case BOPY_KERNEL_RBF: return launch_grad_t<NA, K_RBF>(p, grid, st);
        case BOPY_KERNEL_MATERN12: return launch_grad_t<NA, K_M12>(p, grid, st);
        case BOPY_KERNEL_MATERN32: return launch_grad_t<NA, K_M32>(p, grid, st);
        case BOPY_KERNEL_MATERN52: return launch_grad_t<NA, K_M52>(p, grid, st);
    }
    return fail(BOPY_ERR_BAD_ARG, "unknown kernel id %d", kernel);
}

constexpr long long HOST_CALL_MAX_M = 4096;

template <int KIND> void launch_mean_bound_k(const bopy_gp* gp, const double* Xs, long long m, const LsParam& ls, double sd_max,
                                             int acq, double eta, double kappa, double* mean_out, double* bound_out,
                                             cudaStream_t st) {
    const unsigned grid = (unsigned)((m + PRUNE_NT - 1) / PRUNE_NT);
    const size_t smem = (size_t)(gp->d + 1) * BM * sizeof(double);
#define BOPY_MB(DP) \
    mean_bound_kernel<KIND, DP><<<grid, PRUNE_NT, smem, st>>>(gp->Xt, Xs, m, gp->n_blocks, gp->d, ls, gp->amp, gp->y_mean, \
                                                              gp->y_std, sd_max, acq, eta, kappa, mean_out, bound_out)
    if (gp->d <= 2) BOPY_MB(2);
    else if (gp->d <= 4) BOPY_MB(4);
    else if (gp->d <= 8) BOPY_MB(8);
    else if (gp->d <= 16) BOPY_MB(16);
    else BOPY_MB(32);
#undef BOPY_MB
}

void launch_mean_bound(const bopy_gp* gp, const double* Xs, long long m, double sd_max, int acq, double eta, double kappa,
                       double* mean_out, double* bound_out, cudaStream_t st) {
    LsParam ls;
    for (int q = 0; q < MAX_D; ++q) ls.v[q] = q < gp->d ? gp->ls[q] : 1.0;
    switch (gp->kernel) {
        case BOPY_KERNEL_RBF: launch_mean_bound_k<K_RBF>(gp, Xs, m, ls, sd_max, acq, eta, kappa, mean_out, bound_out, st); break;
        case BOPY_KERNEL_MATERN12: launch_mean_bound_k<K_M12>(gp, Xs, m, ls, sd_max, acq, eta, kappa, mean_out, bound_out, st); break;
        case BOPY_KERNEL_MATERN32: launch_mean_bound_k<K_M32>(gp, Xs, m, ls, sd_max, acq, eta, kappa, mean_out, bound_out, st); break;
        default: launch_mean_bound_k<K_M52>(gp, Xs, m, ls, sd_max, acq, eta, kappa, mean_out, bound_out, st); break;
    }
}

// flags / partial sums / moment scratch of grad_kernel, allocated on first use (chunks of up to sm_count batches)
int ensure_grad_buffers(bopy_gp* gp) {
    if (gp->grad_mv != nullptr) return BOPY_OK;   // grad_mv is allocated last: its presence means all three are there
    const int gb = gp->sm_count, nb = gp->n_blocks;
    const size_t nflags = (size_t)2 * gb * nb;
    if (gp->grad_flags == nullptr) {
        CUDA_TRY(cudaMalloc(&gp->grad_flags, nflags * sizeof(unsigned)));
        CUDA_TRY(cudaMemset(gp->grad_flags, 0, nflags * sizeof(unsigned)));
    }
    if (gp->grad_part == nullptr)
        CUDA_TRY(cudaMalloc(&gp->grad_part, (size_t)gb * nb * 2 * gp->d * PROBE_MAX_NC * sizeof(double)));
    CUDA_TRY(cudaMalloc(&gp->grad_mv, (size_t)2 * gb * PROBE_MAX_NC * sizeof(double)));
    return BOPY_OK;
}

// alpha_ = L^-T (L^-1 yn) on the PACKED factor, spread over the block rows: the latency path's forward solve with yn as
// right-hand side (probe_kernel, rhs mode), then its mirror image (grad_kernel, solve-only mode).  ~n/128 hops of a few
// microseconds each way instead of one thread block walking the whole factor twice (solve_alpha_kernel).
int solve_alpha_chain(bopy_gp* gp, const double* yn_dev, double* alpha_dev, cudaStream_t st) {
    int rc = ensure_grad_buffers(gp);
    if (rc != BOPY_OK) return rc;
    const int gb = gp->sm_count, nb = gp->n_blocks;
    const long long chunk = (long long)gb * PROBE_MAX_NC;
    const ProbePlan pl = probe_plan(gp, 1);
    rc = launch_probe(gp, pl, gp->Xt /* any d doubles: the candidate is not used */, 1, BOPY_ACQ_NONE, 0.0, 0.0, nullptr,
                      nullptr, nullptr, 0, nullptr, 1, st, yn_dev);
    if (rc != BOPY_OK) return rc;
    GradParams q;
    std::memset(&q, 0, sizeof(q));
    q.Lt = reinterpret_cast<const unsigned char*>(gp->Lt);
    q.Xt = gp->Xt;
    q.V = reinterpret_cast<const double*>(gp->Vws);
    q.W = reinterpret_cast<double*>(gp->Vws) + chunk * gp->n_pad;
    q.Xs = gp->Xt;
    q.m = 1;
    q.nbatch = pl.nbatch;
    q.groups = pl.groups;
    q.n = (int)gp->n;
    q.n_blocks = nb;
    q.d = gp->d;
    for (int k = 0; k < gp->d; ++k) q.ls[k] = gp->ls[k];
    q.amp = gp->amp;
    q.y_std = gp->y_std;
    q.y_var = gp->y_std * gp->y_std;
    q.acq = A_LCB;
    q.flags = gp->grad_flags;
    q.flags2 = gp->grad_flags + (size_t)gb * nb;
    q.gpart = gp->grad_part;
    q.mean = gp->grad_mv;
    q.var = gp->grad_mv;
    q.grad_out = gp->grad_part;
    q.ticket = gp->probe_flags;
    q.ticket_base = gp->probe_ticket_base;
    if (++gp->grad_epoch == 0) {
        CUDA_TRY(cudaMemsetAsync(gp->grad_flags, 0, (size_t)2 * gb * nb * sizeof(unsigned), st));
        gp->grad_epoch = 1;
    }
    q.epoch = gp->grad_epoch;
    q.solve_only = 1;
    rc = launch_grad_k<1>(gp->kernel, q, pl.grid, st);
    if (rc != BOPY_OK) return rc;
    gp->probe_ticket_base += (unsigned)pl.grid;
    extract_column_kernel<<<(unsigned)((gp->n + 255) / 256), 256, 0, st>>>(q.W, (int)gp->n, alpha_dev);
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

// group mode serves the fp64 DMMA sweep on the handle's own workspace (never the V-exporting sweep of predict_cov)
bool group_applies(const bopy_gp* gp, long long ntiles, int slot_per_tile, const void* Vws) {
    return gp->group_size > 1 && gp->dtype == BOPY_F64 && !gp->fma64 && slot_per_tile == 0 && Vws == gp->Vws &&
           ntiles < (1LL << 31);
}

// n <= 256 (one or two block rows) on an fp64 DMMA handle, no V export: the warp-autonomous kernel (sweep_warp_kernel.cuh)
bool warp_applies(const bopy_gp* gp, int slot_per_tile) {
    return gp->warp_stages > 0 && slot_per_tile == 0;
}

// launch shape of group mode for a sweep of ntiles tiles.  Few tiles: larger groups, so that every SM works (a tile
// cannot use more thread blocks than it has block rows)
struct GroupPlan {
    int G, grid, ngroups;
};
GroupPlan group_plan(const bopy_gp* gp, long long ntiles) {
    GroupPlan pl;
    pl.G = gp->group_size;
    if (ntiles * pl.G < gp->sm_count)
        pl.G = (int)std::max<long long>(pl.G, std::min<long long>(gp->n_blocks, gp->sm_count / ntiles));
    pl.grid = (int)std::min<long long>(ntiles * pl.G, gp->sm_count);
    pl.ngroups = (pl.grid + pl.G - 1) / pl.G;
    return pl;
}

// the one place the sweep is launched from
int run_sweep(bopy_gp* gp, const double* Xs, long long m, int acq, double eta, double kappa, double* mean_out,
              double* var_out, double* acq_out, long long index_base, double* min_val, long long* min_idx,
              void* Vws, int slot_per_tile, cudaStream_t st, MinLoc* tile_records = nullptr, bool allow_probe = true,
              bool allow_inv = true, const double* xs_inline_host = nullptr) {   // xs_inline_host: small_n handles, m <= 8 only
    SweepParams p;
    std::memset(&p, 0, sizeof(p));
    p.Lt = gp->Lt;
    p.Xt = gp->Xt;
    p.Vws = Vws;
    p.Xs = Xs;
    p.m = m;
    p.ntiles = (m + BN - 1) / BN;
    p.n = (int)gp->n;
    p.n_blocks = gp->n_blocks;
    p.d = gp->d;
    p.slot_per_tile = slot_per_tile;
    for (int q = 0; q < gp->d; ++q) p.ls[q] = gp->ls[q];
    p.amp = gp->amp;
    p.kss = gp->amp + gp->noise;
    p.y_mean = gp->y_mean;
    p.y_std = gp->y_std;
    p.y_var = gp->y_std * gp->y_std;
    p.acq = acq;
    p.eta = eta;
    p.kappa = kappa;
    p.mean_out = mean_out;
    p.var_out = var_out;
    p.acq_out = acq_out;
    p.index_base = index_base;
    p.nan_skip = gp->nan_skip;
    const bool want_min = (min_val != nullptr || min_idx != nullptr);
    if (small_applies(gp, slot_per_tile)) {
        SmallParams q;
        std::memset(&q, 0, sizeof(q));
        q.Xt = gp->Xt;
        q.Dinv = gp->Dinv;
        q.Xs = Xs;
        if (xs_inline_host != nullptr) {
            q.Xs = nullptr;
            std::memcpy(q.xs_inline, xs_inline_host, (size_t)m * gp->d * sizeof(double));
        }
        q.m = m;
        q.ntiles = (m + SMALL_NT - 1) / SMALL_NT;
        q.n = (int)gp->n;
        q.d = gp->d;
        for (int k = 0; k < gp->d; ++k) q.ls[k] = gp->ls[k];
        q.amp = gp->amp;
        q.kss = p.kss;
        q.y_mean = gp->y_mean;
        q.y_std = gp->y_std;
        q.y_var = p.y_var;
        q.acq = acq;
        q.eta = eta;
        q.kappa = kappa;
        q.mean_out = mean_out;
        q.var_out = var_out;
        q.acq_out = acq_out;
        q.index_base = index_base;
        q.partials = want_min ? gp->partials : nullptr;
        q.tile_records = tile_records;
        q.nan_skip = gp->nan_skip;
        const int grid = small_grid(gp, m);
        int rc = gp->n <= 8 ? launch_small_k<8>(gp->kernel, q, grid, st)
                            : (gp->n <= 16 ? launch_small_k<16>(gp->kernel, q, grid, st) : launch_small_k<32>(gp->kernel, q, grid, st));
        if (rc != BOPY_OK) return rc;
        if (want_min) {
            minloc_finalize_kernel<<<1, 256, 0, st>>>(gp->partials, grid, min_val, min_idx);
            CUDA_TRY(cudaGetLastError());
        }
        return BOPY_OK;
    }
    if (allow_probe && allow_inv && inv_applies(gp, m, slot_per_tile, tile_records)) {
        // a handful of candidates on a state that is probed often: one matrix-vector product with W = L^-1
        bool use = false;
        const int rc = inv_prepare(gp, st, &use);
        if (rc != BOPY_OK) return rc;
        if (use)
            return launch_inv(gp, Xs, m, acq, eta, kappa, mean_out, var_out, acq_out, index_base, min_val, min_idx, st);
    }
    if (allow_probe && probe_applies(gp, m, slot_per_tile, tile_records)) {
        // small m: latency path, the forward substitution spread over the block rows of L (probe_kernel.cuh)
        const ProbePlan pl = probe_plan(gp, m);
        int rc = launch_probe(gp, pl, Xs, m, acq, eta, kappa, mean_out, var_out, acq_out, index_base,
                              want_min ? gp->probe_records : nullptr, 0, st);
        if (rc != BOPY_OK) return rc;
        if (want_min) {
            minloc_finalize_kernel<<<1, 256, 0, st>>>(gp->probe_records, pl.nbatch, min_val, min_idx);
            CUDA_TRY(cudaGetLastError());
        }
        return BOPY_OK;
    }
    p.partials = want_min ? gp->partials : nullptr;
    p.tile_records = tile_records;
    p.zigzag = gp->group_size == 1 ? 1 : 0;
    int grid = (int)std::min<long long>(p.ntiles, gp->sm_count);
    int rc;
    if (warp_applies(gp, slot_per_tile)) {
        // n <= 256: warp-autonomous kernel, two thread blocks per SM, no workspace
        grid = (int)std::min<long long>(p.ntiles, 2LL * gp->sm_count);
        rc = launch_sweep_warp(gp->kernel, p, grid, gp->warp_stages, st);
    } else if (group_applies(gp, p.ntiles, slot_per_tile, Vws)) {
        // group mode: G thread blocks per candidate tile, ceil(grid / G) x slots tiles of V in flight
        const GroupPlan pl = group_plan(gp, p.ntiles);
        const int S = gp->group_slots, ng_max = (gp->sm_count + gp->group_size - 1) / gp->group_size;
        const size_t ctl_bytes = GroupCtl::words(ng_max, S, gp->n_blocks) * sizeof(unsigned);
        if (gp->gctl == nullptr) {
            CUDA_TRY(cudaMalloc(&gp->gctl, ctl_bytes));
            CUDA_TRY(cudaMalloc(&gp->gpart, (size_t)ng_max * S * gp->n_blocks * 2 * BN * sizeof(double)));
        }
        grid = pl.grid;
        p.group_size = pl.G;
        p.group_slots = S;
        p.group_lead = gp->group_lead;
        p.gctl = gp->gctl;
        p.gpart = gp->gpart;
        CUDA_TRY(cudaMemsetAsync(gp->gctl, 0, ctl_bytes, st));
        rc = launch_sweep_group(gp->kernel, p, grid, st);
    } else {
        rc = dispatch_engine(gp, [&](auto pol) { return launch_sweep_k<decltype(pol)>(gp->kernel, p, grid, st); });
    }
    if (rc != BOPY_OK) return rc;
    if (want_min) {
        minloc_finalize_kernel<<<1, 256, 0, st>>>(gp->partials, grid, min_val, min_idx);
        CUDA_TRY(cudaGetLastError());
    }
    return BOPY_OK;
}

}  // namespace

extern "C" {

int bopy_abi_version(void) { return BOPY_B200_ABI_VERSION; }

const char* bopy_last_error(void) { return g_last_error.c_str(); }

int bopy_gp_create(bopy_gp** out, int device, int dtype, int kernel, int64_t n, int d) {
    if (out == nullptr) return fail(BOPY_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (dtype != BOPY_F64 && dtype != BOPY_F32) return fail(BOPY_ERR_BAD_ARG, "dtype must be BOPY_F64 or BOPY_F32");
    if (kernel < BOPY_KERNEL_RBF || kernel > BOPY_KERNEL_MATERN52) return fail(BOPY_ERR_BAD_ARG, "unknown kernel id %d", kernel);
    if (n < 1) return fail(BOPY_ERR_BAD_ARG, "n must be >= 1 (got %lld)", (long long)n);
    if (d < 1) return fail(BOPY_ERR_BAD_ARG, "d must be >= 1 (got %d)", d);
    if (d > MAX_D) return fail(BOPY_ERR_UNSUPPORTED, "d = %d exceeds the supported maximum of %d", d, MAX_D);
    if (n > (1 << 20)) return fail(BOPY_ERR_UNSUPPORTED, "n = %lld exceeds the supported maximum", (long long)n);
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(BOPY_ERR_BAD_ARG, "device %d out of range (%d visible)", device, count);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(BOPY_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    bopy_gp* gp = new (std::nothrow) bopy_gp();
    if (gp == nullptr) return fail(BOPY_ERR_BAD_ARG, "out of host memory");
    gp->device = device;
    gp->dtype = dtype;
    gp->kernel = kernel;
    gp->n = n;
    gp->d = d;
    gp->n_blocks = (int)((n + BM - 1) / BM);
    gp->n_pad = gp->n_blocks * BM;
    gp->sm_count = prop.multiProcessorCount;
    {   // keep stream-ordered scratch (fit / LML / segment records) in the pool instead of returning it at every sync
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    const char* engine = std::getenv("BOPY_B200_F64_ENGINE");
    gp->fma64 = engine != nullptr && std::strcmp(engine, "fma") == 0;
    const char* engine32 = std::getenv("BOPY_B200_F32_ENGINE");
    gp->f32_engine = engine32 == nullptr ? 0 : (std::strcmp(engine32, "fma") == 0 ? 2 : (std::strcmp(engine32, "mma_sync") == 0 ? 1 : 0));
    configure_group_mode(gp, env_int("BOPY_B200_SWEEP_GROUP", -1), env_int("BOPY_B200_SWEEP_SLOTS", -1),
                         env_int("BOPY_B200_SWEEP_LEAD", -1));
    // opt-in experiment (BOPY_B200_WARP_KERNEL=1): measured slower than the blocked kernel at C3 (4.30 vs 3.33 ms), DESIGN.md section 4
    if (dtype == BOPY_F64 && !gp->fma64 && gp->n_blocks <= WK_MAX_BLOCKS && env_int("BOPY_B200_WARP_KERNEL", 0) != 0)
        gp->warp_stages = wk_stages(d, gp->n_blocks);
    const size_t es = v_entry_bytes(gp);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&gp->Lt, (size_t)packed_tiles(gp) * TILE_BYTES);
    if (e == cudaSuccess) e = cudaMalloc(&gp->Xt, (size_t)gp->n_blocks * (d + 1) * BM * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&gp->Dinv, (size_t)gp->n_blocks * BM * BM * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&gp->Vws, (size_t)gp->sm_count * gp->n_pad * BN * es);
    if (e == cudaSuccess) e = cudaMalloc(&gp->partials, (size_t)gp->sm_count * SMALL_CTAS_PER_SM * sizeof(MinLoc));
    {
        const char* sn = std::getenv("BOPY_B200_SMALL_N");
        gp->small_n = !(sn != nullptr && std::strcmp(sn, "0") == 0);
    }
    // latency path: fp64 DMMA handles whose block rows fit one wave of CTAs (the V chain needs them all in flight)
    gp->probe_capable = dtype == BOPY_F64 && !gp->fma64 && gp->n_blocks <= gp->sm_count;
    if (gp->probe_capable) {
        gp->probe_max_batch = 4 * gp->sm_count;   // sm_count*128 candidates in batches of 32; >= one batch of 8 per CTA group
        const size_t nflags = (size_t)2 * gp->probe_max_batch * gp->n_blocks + 1;
        if (e == cudaSuccess) e = cudaMalloc(&gp->Mt, (size_t)gp->n_blocks * EngineF64::CHG * TILE_BYTES);
        if (e == cudaSuccess) e = cudaMalloc(&gp->probe_records, (size_t)gp->probe_max_batch * sizeof(MinLoc));
        if (e == cudaSuccess) e = cudaMalloc(&gp->probe_flags, nflags * sizeof(unsigned));
        if (e == cudaSuccess) e = cudaMemset(gp->probe_flags, 0, nflags * sizeof(unsigned));
        if (e == cudaSuccess)
            e = cudaMalloc(&gp->probe_part, (size_t)gp->probe_max_batch * gp->n_blocks * 2 * PROBE_MAX_NC * sizeof(double));
        // default switch-over: 4096 = the measured crossover against the one-tile-per-block kernel at n = 2048.  In group mode
        // the throughput kernel spreads a few tiles over all SMs (its time for small m is one dependency chain of n/128 hops,
        // ~31 us each), while the latency path serves 32 candidates per 4.4 us hop and group of thread blocks: they meet at
        // m ~ 225 x (148 / block rows): ~2000 at n = 2048, ~450 at n = 8192 (profiles/r02/latency_small_m.log)
        long long max_m = 4096;
        if (gp->group_size >= 2) max_m = std::min<long long>(max_m, 225LL * std::max(1, gp->sm_count / gp->n_blocks));
        if (const char* v = std::getenv("BOPY_B200_PROBE_MAX_M")) max_m = std::atoll(v);
        gp->probe_max_m = std::max(0LL, std::min<long long>(max_m, probe_capacity(gp)));
    }
    gp->inv_max_m = inv_capacity(gp);
    gp->inv_mode = std::max(-1, std::min(1, env_int("BOPY_B200_INVERSE_PATH", -1)));
    if (e != cudaSuccess) {
        bopy_gp_destroy(gp);
        return fail(BOPY_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    }
    *out = gp;
    return BOPY_OK;
}

void bopy_gp_destroy(bopy_gp* gp) {
    if (gp == nullptr) return;
    cudaSetDevice(gp->device);
    cudaFree(gp->Lt);
    cudaFree(gp->Xt);
    cudaFree(gp->Dinv);
    cudaFree(gp->Vws);
    cudaFree(gp->partials);
    cudaFree(gp->probe_records);
    cudaFree(gp->probe_flags);
    cudaFree(gp->probe_part);
    cudaFree(gp->Mt);
    cudaFree(gp->probe_trace);
    cudaFree(gp->grad_flags);
    cudaFree(gp->grad_part);
    cudaFree(gp->grad_mv);
    cudaFree(gp->host_x);
    cudaFree(gp->host_out);
    cudaFree(gp->Lfull);
    cudaFree(gp->Winv);
    cudaFree(gp->inv_part);
    cudaFree(gp->inv_ticket);
    if (gp->inv_host_out != nullptr) cudaFreeHost(gp->inv_host_out);
    cudaFree(gp->gctl);
    cudaFree(gp->gpart);
    delete gp;
}

}  // extern "C"

namespace {

int check_hyper(const bopy_gp* gp, const double* length_scale_host, int n_ls, double amplitude, double noise_level) {
    if (length_scale_host == nullptr) return fail(BOPY_ERR_BAD_ARG, "length_scale_host is NULL");
    if (n_ls != 1 && n_ls != gp->d) return fail(BOPY_ERR_BAD_ARG, "n_ls must be 1 or d = %d (got %d)", gp->d, n_ls);
    for (int q = 0; q < n_ls; ++q)
        if (!(length_scale_host[q] > 0.0)) return fail(BOPY_ERR_BAD_ARG, "length_scale[%d] must be positive", q);
    if (!(amplitude > 0.0)) return fail(BOPY_ERR_BAD_ARG, "amplitude must be positive");
    if (!(noise_level >= 0.0)) return fail(BOPY_ERR_BAD_ARG, "noise_level must be non-negative");
    return BOPY_OK;
}

LsParam store_hyper(bopy_gp* gp, const double* length_scale_host, int n_ls, double amplitude, double noise_level,
                    double y_mean, double y_std) {
    LsParam ls;
    for (int q = 0; q < MAX_D; ++q) ls.v[q] = 1.0;
    for (int q = 0; q < gp->d; ++q) {
        gp->ls[q] = length_scale_host[n_ls == 1 ? 0 : q];
        ls.v[q] = gp->ls[q];
    }
    gp->amp = amplitude;
    gp->noise = noise_level;
    gp->y_mean = y_mean;
    gp->y_std = y_std;
    return ls;
}

// pack (L, Dinv, X, alpha) into the sweep layout; gp->Dinv must already hold the inverted diagonal blocks
int pack_factor(bopy_gp* gp, const double* L_dev, int ld, cudaStream_t st);
int pack_x(bopy_gp* gp, const double* X_dev, const double* alpha_dev, const LsParam& ls, cudaStream_t st);

int pack_state(bopy_gp* gp, const double* X_dev, const double* L_dev, int ld, const double* alpha_dev, const LsParam& ls,
               cudaStream_t st) {
    int rc = pack_factor(gp, L_dev, ld, st);
    return rc != BOPY_OK ? rc : pack_x(gp, X_dev, alpha_dev, ls, st);
}

int pack_factor(bopy_gp* gp, const double* L_dev, int ld, cudaStream_t st) {
    const int n = (int)gp->n;
    dispatch_engine(gp, [&](auto e) {
        using E = decltype(e);
        dim3 grid((gp->n_blocks - 1) * E::CHG + E::CHD, gp->n_blocks);
        pack_tiles_kernel<E><<<grid, 256, 0, st>>>(L_dev, n, ld, gp->Dinv, reinterpret_cast<unsigned char*>(gp->Lt));
        return BOPY_OK;
    });
    CUDA_TRY(cudaGetLastError());
    if (gp->probe_capable && gp->n_blocks > 1) {
        pack_m_kernel<<<gp->n_blocks - 1, 256, 0, st>>>(L_dev, n, ld, gp->Dinv, reinterpret_cast<unsigned char*>(gp->Mt));
        CUDA_TRY(cudaGetLastError());
    }
    return BOPY_OK;
}

int pack_x(bopy_gp* gp, const double* X_dev, const double* alpha_dev, const LsParam& ls, cudaStream_t st) {
    pack_x_kernel<<<gp->n_blocks, BM, 0, st>>>(X_dev, alpha_dev, (int)gp->n, gp->d, ls, gp->Xt);
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

template <int KIND> void launch_gram_t(const bopy_gp* gp, const double* X, const LsParam& ls, double amp, double diag,
                                       double* A, int ld, int n_fill, cudaStream_t st) {
    dim3 block(32, 8), grid((unsigned)((n_fill + 31) / 32), (unsigned)((n_fill + 7) / 8));
    gram_kernel<KIND><<<grid, block, 0, st>>>(X, (int)gp->n, gp->d, ls, amp, diag, A, ld, n_fill);
}

void launch_gram(const bopy_gp* gp, const double* X, const LsParam& ls, double amp, double diag, double* A, int ld,
                 int n_fill, cudaStream_t st) {
    switch (gp->kernel) {
        case BOPY_KERNEL_RBF: launch_gram_t<K_RBF>(gp, X, ls, amp, diag, A, ld, n_fill, st); break;
        case BOPY_KERNEL_MATERN12: launch_gram_t<K_M12>(gp, X, ls, amp, diag, A, ld, n_fill, st); break;
        case BOPY_KERNEL_MATERN32: launch_gram_t<K_M32>(gp, X, ls, amp, diag, A, ld, n_fill, st); break;
        default: launch_gram_t<K_M52>(gp, X, ls, amp, diag, A, ld, n_fill, st); break;
    }
}

// in-place blocked Cholesky of the lower triangle of A (n x n, leading dimension ld); Dinv receives inv(L_JJ)
// Look-ahead of depth one: the serial chain chol(J) -> panel(J) -> trailing update of block column J+1 runs on a
// HIGH-priority side stream, the bulk of the trailing update of step J (block columns >= J+2) stays on the caller's
// stream, so that the chain of step J+1 runs beside it (thread blocks of the higher-priority stream are dispatched
// first as multiprocessors free up; with equal priorities the bulk's queue starves the chain and nothing overlaps).
struct CholStreams {
    cudaStream_t chain = nullptr;
    cudaEvent_t start = nullptr, end = nullptr, panel_done[2] = {nullptr, nullptr}, rest_done[2] = {nullptr, nullptr};
    bool ok = false;
};
CholStreams& chol_streams() {
    static thread_local CholStreams per_device[64];   // a stream belongs to the device that was current at its creation
    int dev = 0;
    cudaGetDevice(&dev);
    CholStreams& cs = per_device[dev & 63];
    if (!cs.ok) {
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        bool good = cudaStreamCreateWithPriority(&cs.chain, cudaStreamNonBlocking, greatest) == cudaSuccess;
        good = good && cudaEventCreateWithFlags(&cs.start, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&cs.end, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < 2 && good; ++i)
            good = cudaEventCreateWithFlags(&cs.panel_done[i], cudaEventDisableTiming) == cudaSuccess &&
                   cudaEventCreateWithFlags(&cs.rest_done[i], cudaEventDisableTiming) == cudaSuccess;
        cs.ok = good;
    }
    return cs;
}

// panel / trailing-update tiles of the blocked Cholesky: gemm_nt_async_kernel (cp.async ring); BOPY_B200_CHOL_GEMM=registers
// selects the register-staged gemm_nt_kernel of the first version for A/B runs
void launch_gemm_nt(bool staged, int grid, cudaStream_t s, double* A, int n, int ld, int J, const double* Dinv, int mode) {
    if (staged) gemm_nt_kernel<<<grid, NT, 0, s>>>(A, n, ld, J, Dinv, mode);
    else gemm_nt_async_kernel<<<grid, NT, TG_SMEM_BYTES, s>>>(A, n, ld, J, Dinv, mode);
}

void launch_cholesky(double* A, int n, int ld, int nb, double* Dinv, int* status, cudaStream_t st) {
    const size_t chol_smem = chol_smem_bytes();
    const char* gemm_env = std::getenv("BOPY_B200_CHOL_GEMM");
    const bool staged = gemm_env != nullptr && std::strcmp(gemm_env, "registers") == 0;
    if (!staged) cudaFuncSetAttribute(gemm_nt_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM_BYTES);
    cudaFuncSetAttribute(chol_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem);
    CholStreams& cs = chol_streams();
    if (!(cs.ok && nb >= 4)) {   // plain right-looking order on one stream
        for (int J = 0; J < nb; ++J) {
            chol_block_kernel<<<1, CHOL_NT, chol_smem, st>>>(A, n, ld, J, Dinv, status);
            const int below = nb - J - 1;
            if (below > 0) {
                launch_gemm_nt(staged, below, st, A, n, ld, J, Dinv, 0);
                launch_gemm_nt(staged, below * (below + 1) / 2, st, A, n, ld, J, Dinv, 1);
            }
        }
        return;
    }
    cudaStream_t hi = cs.chain;
    cudaEventRecord(cs.start, st);            // everything queued on st so far (the Gram matrix) comes first
    cudaStreamWaitEvent(hi, cs.start, 0);
    bool rest_pending = false;
    long long* prof = nullptr;                // development aid: phase time stamps of the first diagonal block to stderr
    if (std::getenv("BOPY_B200_CHOL_PROF") != nullptr && cudaMalloc(&prof, 40 * sizeof(long long)) == cudaSuccess)
        cudaMemset(prof, 0, 40 * sizeof(long long));
    for (int J = 0; J < nb; ++J) {
        chol_block_kernel<<<1, CHOL_NT, chol_smem, hi>>>(A, n, ld, J, Dinv, status, J == 0 ? prof : nullptr);
        if (J == 0 && prof != nullptr) {
            long long h[40];
            cudaStreamSynchronize(hi);
            cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
            cudaFree(prof);
            std::fprintf(stderr, "[chol_prof] cycles per panel (a | b | c):");
            for (int pnl = 0; pnl < 8; ++pnl)
                std::fprintf(stderr, " %lld|%lld|%lld", h[1 + 3 * pnl] - h[3 * pnl], h[2 + 3 * pnl] - h[1 + 3 * pnl], h[3 + 3 * pnl] - h[2 + 3 * pnl]);
            std::fprintf(stderr, "  write-back %lld  inverse %lld  store %lld  total %lld\n", h[25] - h[24], h[26] - h[25], h[27] - h[26], h[27] - h[0]);
        }
        const int below = nb - J - 1;
        if (below <= 0) break;
        launch_gemm_nt(staged, below, hi, A, n, ld, J, Dinv, 0);
        cudaEventRecord(cs.panel_done[J & 1], hi);
        if (rest_pending) cudaStreamWaitEvent(hi, cs.rest_done[(J - 1) & 1], 0);   // step J-1 also updated block column J+1
        launch_gemm_nt(staged, below, hi, A, n, ld, J, Dinv, 2);
        rest_pending = false;
        if (below > 1) {
            cudaStreamWaitEvent(st, cs.panel_done[J & 1], 0);
            launch_gemm_nt(staged, (below - 1) * below / 2, st, A, n, ld, J, Dinv, 3);
            cudaEventRecord(cs.rest_done[J & 1], st);
            rest_pending = true;
        }
    }
    cudaEventRecord(cs.end, hi);
    cudaStreamWaitEvent(st, cs.end, 0);       // the caller's stream continues after the whole factorisation
}

}  // namespace

extern "C" {

int bopy_gp_set_state(bopy_gp* gp, const double* X_dev, const double* L_dev, const double* alpha_dev,
                      const double* length_scale_host, int n_ls, double amplitude, double noise_level,
                      double y_mean, double y_std, void* stream) {
    if (gp == nullptr) return fail(BOPY_ERR_BAD_ARG, "gp handle is NULL");
    if (X_dev == nullptr || L_dev == nullptr || alpha_dev == nullptr)
        return fail(BOPY_ERR_BAD_ARG, "X_dev, L_dev and alpha_dev must be non-NULL");
    int rc = check_hyper(gp, length_scale_host, n_ls, amplitude, noise_level);
    if (rc != BOPY_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(gp->device));
    gp->ready = false;
    state_changed(gp);
    gp->Lfull_valid = false;   // bopy_gp_append needs a factor of the library's own fit (its jitter is known)
    gp->Lfull_factor = false;
    const LsParam ls = store_hyper(gp, length_scale_host, n_ls, amplitude, noise_level, y_mean, y_std);
    dinv_kernel<<<gp->n_blocks, BM, 0, st>>>(L_dev, (int)gp->n, (int)gp->n, gp->Dinv);
    CUDA_TRY(cudaGetLastError());
    rc = pack_state(gp, X_dev, L_dev, (int)gp->n, alpha_dev, ls, st);
    if (rc != BOPY_OK) return rc;
    if (gp->inv_max_m > 0 && gp->inv_mode != 0) {
        // the inverse path builds W = L^-1 from the row-major factor: keep a copy (rows beyond n are never read)
        const size_t ld = (size_t)gp->n_pad, n = (size_t)gp->n;
        if (gp->Lfull == nullptr) {
            CUDA_TRY(cudaMalloc(&gp->Lfull, ld * ld * sizeof(double)));
            CUDA_TRY(cudaMemsetAsync(gp->Lfull, 0, ld * ld * sizeof(double), st));
        }
        CUDA_TRY(cudaMemcpy2DAsync(gp->Lfull, ld * sizeof(double), L_dev, n * sizeof(double), n * sizeof(double), n,
                                   cudaMemcpyDeviceToDevice, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    gp->Lfull_factor = gp->inv_max_m > 0 && gp->inv_mode != 0;
    gp->ready = true;
    return BOPY_OK;
}

int bopy_gp_fit(bopy_gp* gp, const double* X_dev, const double* yn_dev, const double* length_scale_host, int n_ls,
                double amplitude, double noise_level, double alpha_reg, double y_mean, double y_std,
                double* L_out_dev, double* alpha_out_dev, void* stream) {
    if (gp == nullptr) return fail(BOPY_ERR_BAD_ARG, "gp handle is NULL");
    if (X_dev == nullptr || yn_dev == nullptr) return fail(BOPY_ERR_BAD_ARG, "X_dev and yn_dev must be non-NULL");
    if (!(alpha_reg >= 0.0)) return fail(BOPY_ERR_BAD_ARG, "alpha_reg must be non-negative");
    int rc = check_hyper(gp, length_scale_host, n_ls, amplitude, noise_level);
    if (rc != BOPY_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(gp->device));
    gp->ready = false;
    state_changed(gp);
    gp->Lfull_factor = false;
    const LsParam ls = store_hyper(gp, length_scale_host, n_ls, amplitude, noise_level, y_mean, y_std);
    const int n = (int)gp->n, nb = gp->n_blocks;
    // the working matrix is the handle's own n_pad x n_pad factor (kept for bopy_gp_append); scratch: z (n_pad),
    // alpha (n), status
    const int ld = gp->n_pad;
    gp->Lfull_valid = false;
    if (gp->Lfull == nullptr) {
        CUDA_TRY(cudaMalloc(&gp->Lfull, (size_t)ld * ld * sizeof(double)));
        CUDA_TRY(cudaMemsetAsync(gp->Lfull, 0, (size_t)ld * ld * sizeof(double), st));   // the upper triangle stays 0
    }
    double* const A = gp->Lfull;
    double* scratch = nullptr;
    int* status = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&scratch), ((size_t)gp->n_pad + n) * sizeof(double) + 16, st);
    if (e != cudaSuccess) return fail(BOPY_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    double* z = scratch;
    double* alpha = alpha_out_dev != nullptr ? alpha_out_dev : scratch + gp->n_pad;
    status = reinterpret_cast<int*>(scratch + gp->n_pad + n);
    auto cleanup = [&]() { cudaFreeAsync(scratch, st); };
    cudaMemsetAsync(status, 0, sizeof(int), st);
    const double diag = (amplitude + noise_level) + alpha_reg;   // kernel_(X) diagonal, then += alpha
    launch_gram(gp, X_dev, ls, gp->amp, diag, A, ld, n, st);
    launch_cholesky(A, n, ld, nb, gp->Dinv, status, st);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        rc = fail(BOPY_ERR_CUDA, "fit kernels failed to launch: %s", cudaGetErrorString(e));
    } else if (gp->probe_capable) {
        // the packed factor first: alpha_ is then solved on it by the chained kernels
        rc = pack_factor(gp, A, ld, st);
        if (rc == BOPY_OK) rc = solve_alpha_chain(gp, yn_dev, alpha, st);
        if (rc == BOPY_OK) rc = pack_x(gp, X_dev, alpha, ls, st);
    } else {
        solve_alpha_kernel<<<1, 1024, 0, st>>>(A, n, ld, nb, gp->Dinv, yn_dev, z, alpha);
        rc = pack_state(gp, X_dev, A, ld, alpha, ls, st);
    }
    if (rc == BOPY_OK && L_out_dev != nullptr)   // export: (n, n) row-major, zeros above the diagonal
        e = cudaMemcpy2DAsync(L_out_dev, (size_t)n * sizeof(double), A, (size_t)ld * sizeof(double), (size_t)n * sizeof(double),
                              (size_t)n, cudaMemcpyDeviceToDevice, st);
    int host_status = 0;
    if (rc == BOPY_OK) e = cudaMemcpyAsync(&host_status, status, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (rc == BOPY_OK && e == cudaSuccess) e = cudaStreamSynchronize(st);
    cleanup();
    if (rc != BOPY_OK) return rc;
    if (e != cudaSuccess) return fail(BOPY_ERR_CUDA, "bopy_gp_fit failed: %s", cudaGetErrorString(e));
    if (host_status != 0)
        return fail(BOPY_ERR_NOT_POSITIVE_DEFINITE,
                    "K + alpha I is not positive definite (non-positive pivot in block column %d); increase alpha",
                    host_status - 1);
    gp->Lfull_valid = true;
    gp->Lfull_factor = true;
    gp->alpha_reg = alpha_reg;
    gp->ready = true;
    return BOPY_OK;
}

int bopy_gp_append(bopy_gp* gp, const double* X_dev, const double* yn_dev, double y_mean, double y_std,
                   double* alpha_out_dev, void* stream) {
    int rc = check_ready(gp);
    if (rc != BOPY_OK) return rc;
    if (X_dev == nullptr || yn_dev == nullptr) return fail(BOPY_ERR_BAD_ARG, "X_dev and yn_dev must be non-NULL");
    if (!gp->Lfull_valid || !gp->probe_capable)
        return fail(BOPY_ERR_NOT_READY, "bopy_gp_append needs a state built by bopy_gp_fit on an fp64 handle");
    const int n = (int)gp->n, nb = gp->n_blocks, ld = gp->n_pad;
    if ((n + 1 + BM - 1) / BM != nb)
        return fail(BOPY_ERR_BAD_ARG, "the grown data set (%d points) needs another 128-row block: refit", n + 1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(gp->device));
    LsParam ls;
    for (int q = 0; q < MAX_D; ++q) ls.v[q] = q < gp->d ? gp->ls[q] : 1.0;
    double* scratch = nullptr;
    CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&scratch), ((size_t)gp->n_pad + n + 1) * sizeof(double) + 16, st));
    double* const z = scratch;
    double* const alpha = alpha_out_dev != nullptr ? alpha_out_dev : scratch + gp->n_pad;
    int* const status = reinterpret_cast<int*>(scratch + gp->n_pad + n + 1);
    cudaMemsetAsync(status, 0, sizeof(int), st);
    // 1. v = L^-1 k(X, x_new): the latency path's forward solve of the new point against the CURRENT state
    gp->ready = false;
    state_changed(gp);
    rc = launch_probe(gp, probe_plan(gp, 1), X_dev + (size_t)n * gp->d, 1, BOPY_ACQ_NONE, 0.0, 0.0, nullptr, nullptr, nullptr, 0,
                      nullptr, 1, st);
    if (rc != BOPY_OK) {
        cudaFreeAsync(scratch, st);
        return rc;
    }
    // 2. the new row of the factor, the inverse of the one diagonal block it touches
    append_row_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const double*>(gp->Vws), n, ld, (gp->amp + gp->noise) + gp->alpha_reg,
                                         gp->Lfull, status);
    gp->n = n + 1;
    gp->y_mean = y_mean;
    gp->y_std = y_std;
    const size_t chol_smem = chol_smem_bytes();
    cudaFuncSetAttribute(dinv_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem);
    dinv_block_kernel<<<1, CHOL_NT, chol_smem, st>>>(gp->Lfull, n + 1, ld, n / BM, gp->Dinv);
    // 3. the packed factor, alpha for the (re-normalised) targets solved on it, then X / alpha
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = fail(BOPY_ERR_CUDA, "append kernels failed to launch: %s", cudaGetErrorString(e));
    if (rc == BOPY_OK) rc = pack_factor(gp, gp->Lfull, ld, st);
    if (rc == BOPY_OK) rc = solve_alpha_chain(gp, yn_dev, alpha, st);
    if (rc == BOPY_OK) rc = pack_x(gp, X_dev, alpha, ls, st);
    (void)z;
    int host_status = 0;
    if (rc == BOPY_OK) e = cudaMemcpyAsync(&host_status, status, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (rc == BOPY_OK && e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(scratch, st);
    if (rc != BOPY_OK || e != cudaSuccess || host_status != 0) {
        gp->Lfull_valid = false;   // row n was written: only a refit restores a consistent state
        gp->Lfull_factor = false;
        if (rc != BOPY_OK) return rc;
        if (e != cudaSuccess) return fail(BOPY_ERR_CUDA, "bopy_gp_append failed: %s", cudaGetErrorString(e));
        return fail(BOPY_ERR_NOT_POSITIVE_DEFINITE, "the grown K + alpha I is not positive definite (new pivot <= 0)");
    }
    gp->ready = true;
    return BOPY_OK;
}

int bopy_gp_truncate(bopy_gp* gp, int64_t n_new, const double* X_dev, const double* yn_dev, double y_mean, double y_std,
                     double* alpha_out_dev, void* stream) {
    int rc = check_ready(gp);
    if (rc != BOPY_OK) return rc;
    if (X_dev == nullptr || yn_dev == nullptr) return fail(BOPY_ERR_BAD_ARG, "X_dev and yn_dev must be non-NULL");
    if (!gp->Lfull_valid || !gp->probe_capable)
        return fail(BOPY_ERR_NOT_READY, "bopy_gp_truncate needs a state built by bopy_gp_fit on an fp64 handle");
    if (n_new < 1 || n_new >= gp->n || (n_new + BM - 1) / BM != gp->n_blocks)
        return fail(BOPY_ERR_BAD_ARG, "n_new = %lld must be in [1, n) and keep the handle's %d block rows: refit",
                    (long long)n_new, gp->n_blocks);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(gp->device));
    LsParam ls;
    for (int q = 0; q < MAX_D; ++q) ls.v[q] = q < gp->d ? gp->ls[q] : 1.0;
    double* scratch = nullptr;
    if (alpha_out_dev == nullptr) CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&scratch), (size_t)n_new * sizeof(double), st));
    double* const alpha = alpha_out_dev != nullptr ? alpha_out_dev : scratch;
    // the factor of the first n_new points is the leading block of the kept factor: only the padding of the last
    // diagonal block, the targets and their normalisation change
    gp->ready = false;
    state_changed(gp);
    gp->n = n_new;
    gp->y_mean = y_mean;
    gp->y_std = y_std;
    const int ld = gp->n_pad;
    const size_t chol_smem = chol_smem_bytes();
    cudaFuncSetAttribute(dinv_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem);
    dinv_block_kernel<<<1, CHOL_NT, chol_smem, st>>>(gp->Lfull, (int)n_new, ld, gp->n_blocks - 1, gp->Dinv);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = fail(BOPY_ERR_CUDA, "truncate kernels failed to launch: %s", cudaGetErrorString(e));
    if (rc == BOPY_OK) rc = pack_factor(gp, gp->Lfull, ld, st);
    if (rc == BOPY_OK) rc = solve_alpha_chain(gp, yn_dev, alpha, st);
    if (rc == BOPY_OK) rc = pack_x(gp, X_dev, alpha, ls, st);
    if (rc == BOPY_OK) e = cudaStreamSynchronize(st);
    if (scratch) cudaFreeAsync(scratch, st);
    if (rc != BOPY_OK || e != cudaSuccess) {
        gp->Lfull_valid = false;
        gp->Lfull_factor = false;
        return rc != BOPY_OK ? rc : fail(BOPY_ERR_CUDA, "bopy_gp_truncate failed: %s", cudaGetErrorString(e));
    }
    gp->ready = true;
    return BOPY_OK;
}

int bopy_gp_lml(bopy_gp* gp, const double* X_dev, const double* yn_dev, const double* length_scale_host, int n_ls,
                double amplitude, double noise_level, double alpha_reg, double* lml_out_host, double* grad_out_host,
                void* stream) {
    if (gp == nullptr) return fail(BOPY_ERR_BAD_ARG, "gp handle is NULL");
    if (X_dev == nullptr || yn_dev == nullptr || lml_out_host == nullptr)
        return fail(BOPY_ERR_BAD_ARG, "X_dev, yn_dev and lml_out_host must be non-NULL");
    if (!(alpha_reg >= 0.0)) return fail(BOPY_ERR_BAD_ARG, "alpha_reg must be non-negative");
    int rc = check_hyper(gp, length_scale_host, n_ls, amplitude, noise_level);
    if (rc != BOPY_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(gp->device));
    LsParam ls;
    for (int q = 0; q < MAX_D; ++q) ls.v[q] = q < gp->d ? length_scale_host[n_ls == 1 ? 0 : q] : 1.0;
    const int n = (int)gp->n, nb = gp->n_blocks, np = gp->n_pad;
    const bool want_grad = grad_out_host != nullptr;
    const int gx = (n + 15) / 16;
    // one pooled allocation: A (np^2), W (np^2, gradient only), Dinv, z, alpha, scalars, partials, status
    const size_t mat = (size_t)np * np;
    const size_t n_part = want_grad ? (size_t)gx * gx * LML_NG : 0;
    const size_t doubles = mat * (want_grad ? 2 : 1) + (size_t)nb * BM * BM + 2 * (size_t)np + 64 + n_part + 2;
    double* buf = nullptr;
    CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&buf), doubles * sizeof(double), st));
    double* A = buf;
    double* W = want_grad ? A + mat : nullptr;
    double* Dinv = A + mat * (want_grad ? 2 : 1);
    double* z = Dinv + (size_t)nb * BM * BM;
    double* alpha = z + np;
    double* scal = alpha + np;            // [0] lml, [1..] gradient
    double* partial = scal + 64;
    int* status = reinterpret_cast<int*>(partial + n_part);
    cudaMemsetAsync(A, 0, mat * (want_grad ? 2 : 1) * sizeof(double), st);
    cudaMemsetAsync(alpha, 0, (size_t)np * sizeof(double), st);
    cudaMemsetAsync(status, 0, sizeof(int), st);
    launch_gram(gp, X_dev, ls, amplitude, (amplitude + noise_level) + alpha_reg, A, np, np, st);
    launch_cholesky(A, n, np, nb, Dinv, status, st);
    solve_alpha_kernel<<<1, 1024, 0, st>>>(A, n, np, nb, Dinv, yn_dev, z, alpha);
    lml_value_kernel<<<1, 1024, 0, st>>>(A, n, np, yn_dev, alpha, scal);
    if (want_grad) {
        if (launch_trtri(A, W, Dinv, nb, np, st) != BOPY_OK) {   // W = L^-1
            cudaFreeAsync(buf, st);
            return BOPY_ERR_CUDA;
        }
        if (trtri_by_diagonals())
            tile_gemm_kernel<<<nb * (nb + 1) / 2, NT, 0, st>>>(2, nb, 0, A, W, Dinv, A, np);   // K^-1 (lower) over L
        else
            tile_gemm_async_kernel<<<nb * (nb + 1) / 2, NT, TG_SMEM_BYTES, st>>>(3, nb, 0, A, W, A, np);
        dim3 block(16, 16), grid(gx, gx);
        switch (gp->kernel) {
            case BOPY_KERNEL_RBF:
                lml_grad_kernel<K_RBF><<<grid, block, 0, st>>>(X_dev, n, gp->d, n_ls, ls, amplitude, noise_level, alpha, A, np, partial);
                break;
            case BOPY_KERNEL_MATERN12:
                lml_grad_kernel<K_M12><<<grid, block, 0, st>>>(X_dev, n, gp->d, n_ls, ls, amplitude, noise_level, alpha, A, np, partial);
                break;
            case BOPY_KERNEL_MATERN32:
                lml_grad_kernel<K_M32><<<grid, block, 0, st>>>(X_dev, n, gp->d, n_ls, ls, amplitude, noise_level, alpha, A, np, partial);
                break;
            default:
                lml_grad_kernel<K_M52><<<grid, block, 0, st>>>(X_dev, n, gp->d, n_ls, ls, amplitude, noise_level, alpha, A, np, partial);
                break;
        }
        lml_grad_finalize_kernel<<<1, 256, 0, st>>>(partial, gx, gx, n_ls + 2, scal + 1);
    }
    cudaError_t e = cudaGetLastError();
    double host[1 + LML_NG];
    int host_status = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(host, scal, sizeof(double) * (1 + (want_grad ? n_ls + 2 : 0)), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&host_status, status, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(buf, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(BOPY_ERR_CUDA, "bopy_gp_lml failed: %s", cudaGetErrorString(e));
    if (host_status != 0)
        return fail(BOPY_ERR_NOT_POSITIVE_DEFINITE, "K + alpha I is not positive definite (block column %d)", host_status - 1);
    *lml_out_host = host[0];
    if (want_grad)
        for (int q = 0; q < n_ls + 2; ++q) grad_out_host[q] = host[1 + q];
    return BOPY_OK;
}

int bopy_gp_posterior_acq(bopy_gp* gp, const double* Xs_dev, int64_t m, int acq, double eta, double kappa,
                          double* mean_out, double* var_out, double* acq_out, int64_t index_base,
                          double* min_val_out, int64_t* min_idx_out, void* stream) {
    int rc = check_ready(gp);
    if (rc != BOPY_OK) return rc;
    if (Xs_dev == nullptr) return fail(BOPY_ERR_BAD_ARG, "Xs_dev is NULL");
    if (m < 1) return fail(BOPY_ERR_BAD_ARG, "m must be >= 1 (got %lld)", (long long)m);
    if (acq < BOPY_ACQ_NONE || acq > BOPY_ACQ_POI) return fail(BOPY_ERR_BAD_ARG, "unknown acquisition id %d", acq);
    if (acq == BOPY_ACQ_NONE && (acq_out != nullptr || min_val_out != nullptr || min_idx_out != nullptr))
        return fail(BOPY_ERR_BAD_ARG, "acquisition outputs requested with BOPY_ACQ_NONE");
    CUDA_TRY(cudaSetDevice(gp->device));
    return run_sweep(gp, Xs_dev, m, acq, eta, kappa, mean_out, var_out, acq_out, index_base, min_val_out,
                     reinterpret_cast<long long*>(min_idx_out), gp->Vws, 0, static_cast<cudaStream_t>(stream));
}

int bopy_gp_resize(bopy_gp* gp, int64_t n) {
    if (gp == nullptr) return fail(BOPY_ERR_BAD_ARG, "gp handle is NULL");
    gp->Lfull_valid = false;
    gp->Lfull_factor = false;
    state_changed(gp);
    if (n < 1 || (n + BM - 1) / BM != gp->n_blocks)
        return fail(BOPY_ERR_BAD_ARG, "n = %lld does not fit this handle's %d block rows of %d (create a new handle)",
                    (long long)n, gp->n_blocks, BM);
    gp->n = n;
    gp->ready = false;
    return BOPY_OK;
}

int bopy_gp_set_group_mode(bopy_gp* gp, int group_size, int slots, int lead, int* effective_out) {
    if (gp == nullptr) return fail(BOPY_ERR_BAD_ARG, "gp handle is NULL");
    if (group_size != -2) {                // -2: report only
        CUDA_TRY(cudaSetDevice(gp->device));
        CUDA_TRY(cudaDeviceSynchronize()); // no sweep of this handle may be in flight while its control buffers change
        configure_group_mode(gp, group_size, slots, lead);
        cudaFree(gp->gctl);
        cudaFree(gp->gpart);
        gp->gctl = nullptr;
        gp->gpart = nullptr;
    }
    if (effective_out) {
        effective_out[0] = gp->group_size;
        effective_out[1] = gp->group_slots;
        effective_out[2] = gp->group_lead;
    }
    return BOPY_OK;
}

int bopy_group_schedule(int64_t job, int n_blocks, int lead, int* tile_seq_out, int* row_out) {
    if (job < 0 || job > 0x7fffffffLL || n_blocks < 1 || lead < 0 || 2 * lead > n_blocks || tile_seq_out == nullptr || row_out == nullptr)
        return fail(BOPY_ERR_BAD_ARG, "bopy_group_schedule: job >= 0, n_blocks >= 1, 0 <= 2 lead <= n_blocks, outputs non-NULL");
    group_job((unsigned)job, n_blocks, lead, *tile_seq_out, *row_out);
    return BOPY_OK;
}

int bopy_gp_probe_trace(bopy_gp* gp, int64_t* stamps_out_host) {
    if (gp == nullptr) return fail(BOPY_ERR_BAD_ARG, "gp handle is NULL");
    if (!gp->probe_capable) return fail(BOPY_ERR_UNSUPPORTED, "this handle has no latency path");
    const size_t bytes = (size_t)gp->n_blocks * 8 * sizeof(long long);
    if (stamps_out_host == nullptr) {   // arm: the next latency-path launches record their batch-0 time stamps
        if (gp->probe_trace == nullptr) CUDA_TRY(cudaMalloc(&gp->probe_trace, bytes));
        CUDA_TRY(cudaMemset(gp->probe_trace, 0, bytes));
        return BOPY_OK;
    }
    if (gp->probe_trace == nullptr) return fail(BOPY_ERR_NOT_READY, "tracing was not armed");
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(stamps_out_host, gp->probe_trace, bytes, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaFree(gp->probe_trace));
    gp->probe_trace = nullptr;
    return BOPY_OK;
}

int bopy_gp_set_latency_path(bopy_gp* gp, int64_t max_m, int64_t* effective_out) {
    if (gp == nullptr) return fail(BOPY_ERR_BAD_ARG, "gp handle is NULL");
    if (max_m < 0) return fail(BOPY_ERR_BAD_ARG, "max_m must be >= 0 (got %lld)", (long long)max_m);
    gp->probe_max_m = gp->probe_capable ? std::min<long long>(max_m, probe_capacity(gp)) : 0;
    if (effective_out) *effective_out = gp->probe_max_m;
    return BOPY_OK;
}

int bopy_gp_set_inverse_path(bopy_gp* gp, int mode, int64_t* max_m_out) {
    if (gp == nullptr) return fail(BOPY_ERR_BAD_ARG, "gp handle is NULL");
    if (mode < -1 || mode > 1) return fail(BOPY_ERR_BAD_ARG, "mode must be -1 (auto), 0 (off) or 1 (on) (got %d)", mode);
    gp->inv_mode = mode;
    gp->inv_small_calls = 0;
    if (max_m_out)
        *max_m_out = (mode != 0 && gp->Lfull_factor)
                         ? (gp->dtype == BOPY_F32 ? gp->inv_max_m : std::min(gp->inv_max_m, gp->probe_max_m))
                         : 0;
    return BOPY_OK;
}

int bopy_acq_value_and_grad(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_dev, int64_t m,
                            double* acq_out, double* grad_out, double* mean_out, double* var_out, void* stream) {
    int rc = check_ready(gp);
    if (rc != BOPY_OK) return rc;
    if (Xs_dev == nullptr || grad_out == nullptr) return fail(BOPY_ERR_BAD_ARG, "Xs_dev / grad_out is NULL");
    if (m < 1) return fail(BOPY_ERR_BAD_ARG, "m must be >= 1 (got %lld)", (long long)m);
    if (acq < BOPY_ACQ_LCB || acq > BOPY_ACQ_POI) return fail(BOPY_ERR_BAD_ARG, "unknown acquisition id %d", acq);
    if (!gp->probe_capable)
        return fail(BOPY_ERR_UNSUPPORTED, "the acquisition gradient needs an fp64 handle with n <= %d", gp->sm_count * BM);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(gp->device));
    const int gb = gp->sm_count, nb = gp->n_blocks;            // batches per chunk
    const long long chunk = (long long)gb * PROBE_MAX_NC;       // candidates per chunk: V and W share the sweep workspace
    rc = ensure_grad_buffers(gp);
    if (rc != BOPY_OK) return rc;
    for (long long off = 0; off < m; off += chunk) {
        const long long mc = std::min<long long>(chunk, m - off);
        const ProbePlan pl = probe_plan(gp, mc);
        double* const mean = mean_out ? mean_out + off : gp->grad_mv;
        double* const var = var_out ? var_out + off : gp->grad_mv + chunk;
        rc = launch_probe(gp, pl, Xs_dev + off * gp->d, mc, acq, eta, kappa, mean, var, acq_out ? acq_out + off : nullptr, 0,
                          nullptr, 1, st);
        if (rc != BOPY_OK) return rc;
        GradParams q;
        std::memset(&q, 0, sizeof(q));
        q.Lt = reinterpret_cast<const unsigned char*>(gp->Lt);
        q.Xt = gp->Xt;
        q.V = reinterpret_cast<const double*>(gp->Vws);
        q.W = reinterpret_cast<double*>(gp->Vws) + chunk * gp->n_pad;
        q.Xs = Xs_dev + off * gp->d;
        q.m = mc;
        q.nbatch = pl.nbatch;
        q.groups = pl.groups;
        q.n = (int)gp->n;
        q.n_blocks = nb;
        q.d = gp->d;
        for (int k = 0; k < gp->d; ++k) q.ls[k] = gp->ls[k];
        q.amp = gp->amp;
        q.y_std = gp->y_std;
        q.y_var = gp->y_std * gp->y_std;
        q.acq = acq;
        q.eta = eta;
        q.kappa = kappa;
        q.mean = mean;
        q.var = var;
        q.grad_out = grad_out + off * gp->d;
        q.flags = gp->grad_flags;
        q.flags2 = gp->grad_flags + (size_t)gb * nb;
        q.gpart = gp->grad_part;
        q.ticket = gp->probe_flags;
        q.ticket_base = gp->probe_ticket_base;
        if (++gp->grad_epoch == 0) {
            CUDA_TRY(cudaMemsetAsync(gp->grad_flags, 0, (size_t)2 * gb * nb * sizeof(unsigned), st));
            gp->grad_epoch = 1;
        }
        q.epoch = gp->grad_epoch;
        rc = pl.na == 1 ? launch_grad_k<1>(gp->kernel, q, pl.grid, st)
                        : (pl.na == 2 ? launch_grad_k<2>(gp->kernel, q, pl.grid, st) : launch_grad_k<4>(gp->kernel, q, pl.grid, st));
        if (rc != BOPY_OK) return rc;
        gp->probe_ticket_base += (unsigned)pl.grid;
    }
    return BOPY_OK;
}

int bopy_acq_eval_host(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_host, int64_t m,
                       double* acq_out_host, double* mean_out_host, double* var_out_host, void* stream) {
    int rc = check_ready(gp);
    if (rc != BOPY_OK) return rc;
    if (Xs_host == nullptr) return fail(BOPY_ERR_BAD_ARG, "Xs_host is NULL");
    if (m < 1 || m > HOST_CALL_MAX_M)
        return fail(BOPY_ERR_BAD_ARG, "m must be in [1, %lld] for the host-buffer entry (got %lld)", HOST_CALL_MAX_M, (long long)m);
    if (acq < BOPY_ACQ_NONE || acq > BOPY_ACQ_POI) return fail(BOPY_ERR_BAD_ARG, "unknown acquisition id %d", acq);
    if (acq == BOPY_ACQ_NONE && acq_out_host != nullptr)
        return fail(BOPY_ERR_BAD_ARG, "acquisition output requested with BOPY_ACQ_NONE");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(gp->device));
    auto pinned_out = [&](double** out_dev) -> int {   // [3][INV_MAX_NC] mapped pinned: acq / mean / var
        if (gp->inv_host_out == nullptr)
            CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&gp->inv_host_out), (size_t)3 * INV_MAX_NC * sizeof(double),
                                   cudaHostAllocMapped));
        CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(out_dev), gp->inv_host_out, 0));
        return BOPY_OK;
    };
    auto fetch_pinned = [&]() {
        const size_t nbytes = (size_t)m * sizeof(double);
        if (acq_out_host) std::memcpy(acq_out_host, gp->inv_host_out, nbytes);
        if (mean_out_host) std::memcpy(mean_out_host, gp->inv_host_out + INV_MAX_NC, nbytes);
        if (var_out_host) std::memcpy(var_out_host, gp->inv_host_out + 2 * INV_MAX_NC, nbytes);
    };
    static_assert(SMALL_INLINE_M == INV_MAX_NC, "one pinned result buffer serves both kernels");
    if (small_applies(gp, 0) && m <= SMALL_INLINE_M) {
        // n <= 32 (the reference's own examples): the thread-per-candidate kernel, candidates as kernel parameters
        double* out_dev = nullptr;
        rc = pinned_out(&out_dev);
        if (rc != BOPY_OK) return rc;
        rc = run_sweep(gp, nullptr, m, acq, eta, kappa, mean_out_host ? out_dev + INV_MAX_NC : nullptr,
                       var_out_host ? out_dev + 2 * INV_MAX_NC : nullptr, acq_out_host ? out_dev : nullptr, 0, nullptr, nullptr,
                       gp->Vws, 0, st, nullptr, true, false, Xs_host);
        if (rc != BOPY_OK) return rc;
        CUDA_TRY(cudaStreamSynchronize(st));
        fetch_pinned();
        return BOPY_OK;
    }
    bool tried_inv = false;
    if (inv_applies(gp, m, 0, nullptr)) {
        // the DIRECT probe on the inverse path: candidates as kernel parameters, results through mapped pinned memory --
        // one launch and one stream synchronisation, no memcpy call on either side
        bool use = false;
        rc = inv_prepare(gp, st, &use);
        if (rc != BOPY_OK) return rc;
        tried_inv = true;
        if (use) {
            double* out_dev = nullptr;
            rc = pinned_out(&out_dev);
            if (rc != BOPY_OK) return rc;
            rc = launch_inv(gp, nullptr, m, acq, eta, kappa, mean_out_host ? out_dev + INV_MAX_NC : nullptr,
                            var_out_host ? out_dev + 2 * INV_MAX_NC : nullptr, acq_out_host ? out_dev : nullptr, 0, nullptr,
                            nullptr, st, Xs_host);
            if (rc != BOPY_OK) return rc;
            CUDA_TRY(cudaStreamSynchronize(st));
            fetch_pinned();
            return BOPY_OK;
        }
    }
    if (gp->host_x == nullptr) {
        CUDA_TRY(cudaMalloc(&gp->host_x, (size_t)HOST_CALL_MAX_M * gp->d * sizeof(double)));
        CUDA_TRY(cudaMalloc(&gp->host_out, (size_t)3 * HOST_CALL_MAX_M * sizeof(double)));
    }
    double* const a_dev = acq_out_host ? gp->host_out : nullptr;
    double* const m_dev = mean_out_host ? gp->host_out + HOST_CALL_MAX_M : nullptr;
    double* const v_dev = var_out_host ? gp->host_out + 2 * HOST_CALL_MAX_M : nullptr;
    CUDA_TRY(cudaMemcpyAsync(gp->host_x, Xs_host, (size_t)m * gp->d * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = run_sweep(gp, gp->host_x, m, acq, eta, kappa, m_dev, v_dev, a_dev, 0, nullptr, nullptr, gp->Vws, 0, st, nullptr, true,
                   !tried_inv);   // (the call has been counted above)
    if (rc != BOPY_OK) return rc;
    const size_t bytes = (size_t)m * sizeof(double);
    if (a_dev) CUDA_TRY(cudaMemcpyAsync(acq_out_host, a_dev, bytes, cudaMemcpyDeviceToHost, st));
    if (m_dev) CUDA_TRY(cudaMemcpyAsync(mean_out_host, m_dev, bytes, cudaMemcpyDeviceToHost, st));
    if (v_dev) CUDA_TRY(cudaMemcpyAsync(var_out_host, v_dev, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return BOPY_OK;
}

int bopy_gp_predict_diag(bopy_gp* gp, const double* Xs_dev, int64_t m, double* mean_out, double* var_out,
                         void* stream) {
    return bopy_gp_posterior_acq(gp, Xs_dev, m, BOPY_ACQ_NONE, 0.0, 0.0, mean_out, var_out, nullptr, 0, nullptr,
                                 nullptr, stream);
}

int bopy_acq_eval(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_dev, int64_t m, double* acq_out,
                  void* stream) {
    if (acq_out == nullptr) return fail(BOPY_ERR_BAD_ARG, "acq_out is NULL");
    return bopy_gp_posterior_acq(gp, Xs_dev, m, acq, eta, kappa, nullptr, nullptr, acq_out, 0, nullptr, nullptr, stream);
}

int bopy_acq_argmin(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_dev, int64_t m,
                    int64_t index_base, double* min_val_out, int64_t* min_idx_out, void* stream) {
    if (min_val_out == nullptr || min_idx_out == nullptr) return fail(BOPY_ERR_BAD_ARG, "min outputs are NULL");
    return bopy_gp_posterior_acq(gp, Xs_dev, m, acq, eta, kappa, nullptr, nullptr, nullptr, index_base, min_val_out,
                                 min_idx_out, stream);
}

int bopy_acq_argmin_pruned(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_dev, int64_t m,
                           int64_t index_base, double* min_val_out, int64_t* min_idx_out, int64_t* stats_out_host,
                           void* stream) {
    int rc = check_ready(gp);
    if (rc != BOPY_OK) return rc;
    if (Xs_dev == nullptr || min_val_out == nullptr || min_idx_out == nullptr)
        return fail(BOPY_ERR_BAD_ARG, "Xs_dev / min outputs are NULL");
    if (m < 1) return fail(BOPY_ERR_BAD_ARG, "m must be >= 1 (got %lld)", (long long)m);
    if (acq < BOPY_ACQ_LCB || acq > BOPY_ACQ_POI) return fail(BOPY_ERR_BAD_ARG, "unknown acquisition id %d", acq);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(gp->device));
    long long* const min_idx = reinterpret_cast<long long*>(min_idx_out);
    // every sweep in here runs the THROUGHPUT kernel, whatever its size: the values must be the ones the plain sweep of
    // all m candidates produces, bit for bit (the latency path orders a few partial sums differently)
    const long long sample = std::min<long long>(m, 8192);
    if (m <= 4 * sample) {   // nothing to gain: plain sweep
        if (stats_out_host) stats_out_host[0] = m, stats_out_host[1] = 0, stats_out_host[2] = m;
        return run_sweep(gp, Xs_dev, m, acq, eta, kappa, nullptr, nullptr, nullptr, index_base, min_val_out, min_idx,
                         gp->Vws, 0, st);   // exactly what bopy_acq_argmin does for this m
    }
    const int d = gp->d;
    const long long per_block = (long long)PRUNE_NT * COMPACT_ITEMS;
    const int nblocks = (int)((m + per_block - 1) / per_block);
    // one pooled scratch: bounds (m), incumbent value (1), sample rows (sample x d), survivor rows (m x d worst case is
    // never needed: sized after the count), block counts, total
    double* bound = nullptr;
    CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&bound),
                             ((size_t)m + 2 + (size_t)sample * d) * sizeof(double) + ((size_t)nblocks + 4) * sizeof(unsigned) + 16, st));
    double* const thr = bound + m;                    // incumbent value, then reused
    double* const srows = thr + 2;
    unsigned* const counts = reinterpret_cast<unsigned*>(srows + (size_t)sample * d);
    long long* const total = reinterpret_cast<long long*>(thr + 1);
    auto cleanup = [&]() { cudaFreeAsync(bound, st); };
    // 1. lower bounds from the posterior mean
    const double sd_max = std::sqrt((gp->amp + gp->noise)) * std::fabs(gp->y_std);
    launch_mean_bound(gp, Xs_dev, m, sd_max, acq, eta, kappa, nullptr, bound, st);
    // 2. incumbent: the true minimum over a strided sample
    const long long stride = m / sample;
    strided_rows_kernel<<<(unsigned)((sample * d + 255) / 256), 256, 0, st>>>(Xs_dev, d, stride, sample, srows);
    rc = run_sweep(gp, srows, sample, acq, eta, kappa, nullptr, nullptr, nullptr, 0, thr, min_idx, gp->Vws, 0, st, nullptr, false);
    if (rc != BOPY_OK) {
        cleanup();
        return rc;
    }
    // 3. survivors = {bound <= incumbent}, ascending order kept
    compact_count_kernel<<<nblocks, PRUNE_NT, 0, st>>>(bound, m, thr, m, counts);
    compact_scan_kernel<<<1, 1024, 0, st>>>(counts, nblocks, total);
    long long count = 0;
    cudaError_t e = cudaMemcpyAsync(&count, total, sizeof(long long), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cleanup();
        return fail(BOPY_ERR_CUDA, "pruned argmin failed: %s", cudaGetErrorString(e));
    }
    if (stats_out_host) stats_out_host[0] = m, stats_out_host[1] = sample, stats_out_host[2] = count;
    if (count <= 0 || count > m / 2) {   // bound not selective (or an all-NaN incumbent): the plain sweep is as good
        cleanup();
        if (stats_out_host) stats_out_host[2] = m;
        return run_sweep(gp, Xs_dev, m, acq, eta, kappa, nullptr, nullptr, nullptr, index_base, min_val_out, min_idx,
                         gp->Vws, 0, st, nullptr, false);
    }
    long long* idx_list = nullptr;
    e = cudaMallocAsync(reinterpret_cast<void**>(&idx_list), (size_t)count * (sizeof(long long) + (size_t)d * sizeof(double)), st);
    if (e != cudaSuccess) {
        cleanup();
        return fail(BOPY_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    }
    double* const rows = reinterpret_cast<double*>(idx_list + count);
    compact_scatter_kernel<<<nblocks, PRUNE_NT, 0, st>>>(bound, m, thr, m, counts, Xs_dev, d, idx_list, rows);
    // 4. the full fused sweep over the survivors only; local index -> original index
    rc = run_sweep(gp, rows, count, acq, eta, kappa, nullptr, nullptr, nullptr, 0, min_val_out, min_idx, gp->Vws, 0, st, nullptr, false);
    if (rc == BOPY_OK) {
        remap_index_kernel<<<1, 1, 0, st>>>(idx_list, count, index_base, min_idx);
        if (cudaGetLastError() != cudaSuccess) rc = fail(BOPY_ERR_CUDA, "remap_index_kernel failed to launch");
    }
    cudaFreeAsync(idx_list, st);
    cleanup();
    return rc;
}

int bopy_acq_segment_argmin(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_dev, int64_t m,
                            int64_t seg_len, int64_t index_base, double* seg_val_out, int64_t* seg_idx_out,
                            void* stream) {
    int rc = check_ready(gp);
    if (rc != BOPY_OK) return rc;
    if (Xs_dev == nullptr || seg_val_out == nullptr || seg_idx_out == nullptr)
        return fail(BOPY_ERR_BAD_ARG, "Xs_dev / seg_val_out / seg_idx_out is NULL");
    if (m < 1) return fail(BOPY_ERR_BAD_ARG, "m must be >= 1 (got %lld)", (long long)m);
    if (acq < BOPY_ACQ_LCB || acq > BOPY_ACQ_POI) return fail(BOPY_ERR_BAD_ARG, "unknown acquisition id %d", acq);
    if (seg_len < BN || seg_len % BN != 0)
        return fail(BOPY_ERR_BAD_ARG, "seg_len must be a positive multiple of %d (got %lld)", BN, (long long)seg_len);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(gp->device));
    const long long ntiles = (m + BN - 1) / BN, nseg = (m + seg_len - 1) / seg_len;
    MinLoc* records = nullptr;
    CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&records), (size_t)ntiles * sizeof(MinLoc), st));
    rc = run_sweep(gp, Xs_dev, m, acq, eta, kappa, nullptr, nullptr, nullptr, index_base, nullptr, nullptr, gp->Vws, 0,
                   st, records);
    if (rc == BOPY_OK) {
        segment_minloc_kernel<<<(unsigned)((nseg + 127) / 128), 128, 0, st>>>(
            records, ntiles, (int)(seg_len / BN), nseg, seg_val_out, reinterpret_cast<long long*>(seg_idx_out));
        if (cudaGetLastError() != cudaSuccess) rc = fail(BOPY_ERR_CUDA, "segment_minloc_kernel failed to launch");
    }
    cudaFreeAsync(records, st);
    return rc;
}

int bopy_acq_segment_argmin_pruned(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_dev, int64_t m,
                                   int64_t seg_len, int64_t index_base, double* seg_val_out, int64_t* seg_idx_out,
                                   int64_t* stats_out_host, void* stream) {
    int rc = check_ready(gp);
    if (rc != BOPY_OK) return rc;
    if (Xs_dev == nullptr || seg_val_out == nullptr || seg_idx_out == nullptr)
        return fail(BOPY_ERR_BAD_ARG, "Xs_dev / seg_val_out / seg_idx_out is NULL");
    if (acq < BOPY_ACQ_LCB || acq > BOPY_ACQ_POI) return fail(BOPY_ERR_BAD_ARG, "unknown acquisition id %d", acq);
    const long long stride = 16;   // one candidate in 16 of every segment is evaluated up front: the segment's incumbent
    if (seg_len < BN || seg_len % BN != 0)
        return fail(BOPY_ERR_BAD_ARG, "seg_len must be a positive multiple of %d (got %lld)", BN, (long long)seg_len);
    if (m < seg_len || m % seg_len != 0)
        return fail(BOPY_ERR_BAD_ARG, "m must be a positive multiple of seg_len for the pruned segmented arg-min");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(gp->device));
    const int d = gp->d;
    const long long nseg = m / seg_len, nsample = m / stride, per_seg = seg_len / stride;
    const long long per_block = (long long)PRUNE_NT * COMPACT_ITEMS;
    const int nblocks = (int)((m + per_block - 1) / per_block);
    double* bound = nullptr;
    CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&bound),
                             ((size_t)m + (size_t)nseg + 2 + (size_t)nsample * (d + 1)) * sizeof(double) +
                                 ((size_t)nblocks + 4) * sizeof(unsigned) + 16, st));
    double* const inc = bound + m;                         // [nseg] incumbents
    long long* const total = reinterpret_cast<long long*>(inc + nseg);
    double* const svals = inc + nseg + 2;                  // [nsample] sample values
    double* const srows = svals + nsample;                 // [nsample][d] sample rows
    unsigned* const counts = reinterpret_cast<unsigned*>(srows + (size_t)nsample * d);
    auto cleanup = [&]() { cudaFreeAsync(bound, st); };
    const double sd_max = std::sqrt((gp->amp + gp->noise)) * std::fabs(gp->y_std);
    launch_mean_bound(gp, Xs_dev, m, sd_max, acq, eta, kappa, nullptr, bound, st);
    strided_rows_kernel<<<(unsigned)((nsample * d + 255) / 256), 256, 0, st>>>(Xs_dev, d, stride, nsample, srows);
    rc = run_sweep(gp, srows, nsample, acq, eta, kappa, nullptr, nullptr, svals, 0, nullptr, nullptr, gp->Vws, 0, st, nullptr, false);
    if (rc != BOPY_OK) {
        cleanup();
        return rc;
    }
    segment_incumbent_kernel<<<(unsigned)((nseg + 127) / 128), 128, 0, st>>>(svals, nsample, per_seg, nseg, inc);
    compact_count_kernel<<<nblocks, PRUNE_NT, 0, st>>>(bound, m, inc, seg_len, counts);
    compact_scan_kernel<<<1, 1024, 0, st>>>(counts, nblocks, total);
    long long count = 0;
    cudaError_t e = cudaMemcpyAsync(&count, total, sizeof(long long), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cleanup();
        return fail(BOPY_ERR_CUDA, "pruned segmented argmin failed: %s", cudaGetErrorString(e));
    }
    if (stats_out_host) stats_out_host[0] = m, stats_out_host[1] = nsample, stats_out_host[2] = count;
    if (count <= 0 || count > m / 2) {   // not selective: the plain segmented sweep
        cleanup();
        if (stats_out_host) stats_out_host[2] = m;
        return bopy_acq_segment_argmin(gp, acq, eta, kappa, Xs_dev, m, seg_len, index_base, seg_val_out, seg_idx_out, stream);
    }
    long long* idx_list = nullptr;
    e = cudaMallocAsync(reinterpret_cast<void**>(&idx_list), (size_t)count * (sizeof(long long) + (size_t)(d + 1) * sizeof(double)), st);
    if (e != cudaSuccess) {
        cleanup();
        return fail(BOPY_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    }
    double* const rows = reinterpret_cast<double*>(idx_list + count);
    double* const vals = rows + (size_t)count * d;
    compact_scatter_kernel<<<nblocks, PRUNE_NT, 0, st>>>(bound, m, inc, seg_len, counts, Xs_dev, d, idx_list, rows);
    rc = run_sweep(gp, rows, count, acq, eta, kappa, nullptr, nullptr, vals, 0, nullptr, nullptr, gp->Vws, 0, st, nullptr, false);
    if (rc == BOPY_OK) {
        segment_argmin_survivors_kernel<<<(unsigned)((nseg + 127) / 128), 128, 0, st>>>(
            idx_list, vals, count, seg_len, nseg, index_base, seg_val_out, reinterpret_cast<long long*>(seg_idx_out));
        if (cudaGetLastError() != cudaSuccess) rc = fail(BOPY_ERR_CUDA, "segment_argmin_survivors_kernel failed to launch");
    }
    cudaFreeAsync(idx_list, st);
    cleanup();
    return rc;
}

int bopy_candidates_around(uint64_t seed, const double* starts_dev, int64_t S, int P, int d,
                           const double* halfwidth_host, const double* lowers_host, const double* uppers_host,
                           double* out_dev, void* stream) {
    if (starts_dev == nullptr || halfwidth_host == nullptr || lowers_host == nullptr || uppers_host == nullptr ||
        out_dev == nullptr)
        return fail(BOPY_ERR_BAD_ARG, "starts/halfwidth/lowers/uppers/out must be non-NULL");
    if (S < 1 || P < 1 || d < 1 || d > MAX_D) return fail(BOPY_ERR_BAD_ARG, "need S >= 1, P >= 1, 1 <= d <= %d", MAX_D);
    BoxParam half, box;
    for (int q = 0; q < MAX_D; ++q) {
        half.lo[q] = half.hi[q] = q < d ? halfwidth_host[q] : 0.0;
        box.lo[q] = q < d ? lowers_host[q] : 0.0;
        box.hi[q] = q < d ? uppers_host[q] : 1.0;
    }
    const long long total = (long long)S * P * d;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    candidates_around_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, starts_dev, S, P, d, half, box,
                                                                               out_dev);
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

int bopy_multistart_step(int64_t S, int d, const double* lowers_host, const double* uppers_host, double* xc_dev,
                         double* fc_dev, double* gc_dev, double* xt_dev, const double* ft_dev, const double* gt_dev,
                         double* alpha_dev, int first, void* stream) {
    if (lowers_host == nullptr || uppers_host == nullptr || xc_dev == nullptr || fc_dev == nullptr || gc_dev == nullptr ||
        xt_dev == nullptr || ft_dev == nullptr || gt_dev == nullptr || alpha_dev == nullptr)
        return fail(BOPY_ERR_BAD_ARG, "bopy_multistart_step: NULL argument");
    if (S < 1 || d < 1 || d > MAX_D) return fail(BOPY_ERR_BAD_ARG, "need S >= 1 and 1 <= d <= %d", MAX_D);
    BoxParam box;
    for (int q = 0; q < MAX_D; ++q) {
        box.lo[q] = q < d ? lowers_host[q] : 0.0;
        box.hi[q] = q < d ? uppers_host[q] : 1.0;
    }
    multistart_step_kernel<<<(unsigned)((S + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        S, d, box, xc_dev, fc_dev, gc_dev, xt_dev, ft_dev, gt_dev, alpha_dev, first);
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

int bopy_gather_rows(const double* Xs_dev, int64_t m, int d, const int64_t* idx_dev, int64_t S, int64_t index_base,
                     double* out_dev, void* stream) {
    if (Xs_dev == nullptr || idx_dev == nullptr || out_dev == nullptr)
        return fail(BOPY_ERR_BAD_ARG, "Xs_dev / idx_dev / out_dev is NULL");
    if (m < 1 || S < 1 || d < 1) return fail(BOPY_ERR_BAD_ARG, "need m, S, d >= 1");
    const long long total = (long long)S * d;
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        Xs_dev, reinterpret_cast<const long long*>(idx_dev), S, d, index_base, m, out_dev);
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

int bopy_acq_from_moments(int acq, double eta, double kappa, const double* mean_dev, const double* var_dev,
                          int64_t m, double* acq_out, int64_t index_base, double* min_val_out,
                          int64_t* min_idx_out, void* stream) {
    if (mean_dev == nullptr || var_dev == nullptr) return fail(BOPY_ERR_BAD_ARG, "mean_dev / var_dev is NULL");
    if (m < 1) return fail(BOPY_ERR_BAD_ARG, "m must be >= 1 (got %lld)", (long long)m);
    if (acq < BOPY_ACQ_LCB || acq > BOPY_ACQ_POI) return fail(BOPY_ERR_BAD_ARG, "unknown acquisition id %d", acq);
    moments_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
        acq, eta, kappa, mean_dev, var_dev, m, acq_out, index_base, min_val_out,
        reinterpret_cast<long long*>(min_idx_out));
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

int bopy_gp_predict_cov(bopy_gp* gp, const double* Xs_dev, int64_t m, double* mean_out, double* cov_out,
                        void* stream) {
    int rc = check_ready(gp);
    if (rc != BOPY_OK) return rc;
    if (Xs_dev == nullptr || cov_out == nullptr) return fail(BOPY_ERR_BAD_ARG, "Xs_dev / cov_out is NULL");
    if (m < 1) return fail(BOPY_ERR_BAD_ARG, "m must be >= 1 (got %lld)", (long long)m);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(gp->device));
    if (m == 1) {
        // ONE point (Surrogate.predict inside a DIRECT objective or the Kriging believer, bopy/acquisition.py:189): the 1 x 1
        // covariance is the posterior variance -- the latency / inverse path instead of one thread block walking all of L
        rc = run_sweep(gp, Xs_dev, 1, BOPY_ACQ_NONE, 0.0, 0.0, mean_out, cov_out, nullptr, 0, nullptr, nullptr, gp->Vws, 0, st);
        if (rc != BOPY_OK) return rc;
        CUDA_TRY(cudaStreamSynchronize(st));
        return BOPY_OK;
    }
    if (!small_applies(gp, 0) && probe_applies(gp, m, 0, nullptr) && m <= probe_capacity(gp)) {
        // a small candidate set (the reference's plotting grids: a few hundred points): V from the latency path -- the solve
        // spread over the block rows of L -- instead of one thread block per 128 candidates walking all of L
        const ProbePlan pl = probe_plan(gp, m);
        rc = launch_probe(gp, pl, Xs_dev, m, BOPY_ACQ_NONE, 0.0, 0.0, mean_out, nullptr, nullptr, 0, nullptr, 1, st);
        if (rc != BOPY_OK) return rc;
        LsParam ls;
        for (int q = 0; q < MAX_D; ++q) ls.v[q] = q < gp->d ? gp->ls[q] : 1.0;
        dim3 block(16, 16), grid((unsigned)((m + 15) / 16), (unsigned)((m + 15) / 16));
        const double kss = gp->amp + gp->noise, yv = gp->y_std * gp->y_std;
        const double* V = reinterpret_cast<const double*>(gp->Vws);
        switch (gp->kernel) {
            case BOPY_KERNEL_RBF:
                cov_probe_kernel<K_RBF><<<grid, block, 0, st>>>(V, pl.na, gp->n_pad, (int)gp->n, Xs_dev, m, gp->d, ls, gp->amp, kss, yv, cov_out);
                break;
            case BOPY_KERNEL_MATERN12:
                cov_probe_kernel<K_M12><<<grid, block, 0, st>>>(V, pl.na, gp->n_pad, (int)gp->n, Xs_dev, m, gp->d, ls, gp->amp, kss, yv, cov_out);
                break;
            case BOPY_KERNEL_MATERN32:
                cov_probe_kernel<K_M32><<<grid, block, 0, st>>>(V, pl.na, gp->n_pad, (int)gp->n, Xs_dev, m, gp->d, ls, gp->amp, kss, yv, cov_out);
                break;
            default:
                cov_probe_kernel<K_M52><<<grid, block, 0, st>>>(V, pl.na, gp->n_pad, (int)gp->n, Xs_dev, m, gp->d, ls, gp->amp, kss, yv, cov_out);
                break;
        }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(st));
        return BOPY_OK;
    }
    const long long ntiles = (m + BN - 1) / BN;
    void* Vall = nullptr;   // stream-ordered: the pool keeps the block between calls (plotting grids call this in a loop)
    CUDA_TRY(cudaMallocAsync(&Vall, (size_t)ntiles * gp->n_pad * BN * v_entry_bytes(gp), st));
    rc = run_sweep(gp, Xs_dev, m, BOPY_ACQ_NONE, 0.0, 0.0, mean_out, nullptr, nullptr, 0, nullptr, nullptr, Vall, 1, st);
    if (rc == BOPY_OK) {
        LsParam ls;
        for (int q = 0; q < MAX_D; ++q) ls.v[q] = q < gp->d ? gp->ls[q] : 1.0;
        rc = dispatch_engine(gp, [&](auto pol) { return launch_cov_k<decltype(pol)>(gp, Vall, Xs_dev, m, ls, cov_out, st); });
    }
    cudaFreeAsync(Vall, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc != BOPY_OK) return rc;
    if (e != cudaSuccess) return fail(BOPY_ERR_CUDA, "predict_cov failed: %s", cudaGetErrorString(e));
    return BOPY_OK;
}

int bopy_candidates_uniform(uint64_t seed, int64_t index_base, int64_t m, int d, const double* lowers_host,
                            const double* uppers_host, double* out_dev, void* stream) {
    if (lowers_host == nullptr || uppers_host == nullptr || out_dev == nullptr)
        return fail(BOPY_ERR_BAD_ARG, "lowers/uppers/out must be non-NULL");
    if (m < 1 || d < 1 || d > MAX_D) return fail(BOPY_ERR_BAD_ARG, "need m >= 1 and 1 <= d <= %d", MAX_D);
    if (index_base < 0) return fail(BOPY_ERR_BAD_ARG, "index_base must be >= 0");
    BoxParam box;
    for (int q = 0; q < MAX_D; ++q) {
        box.lo[q] = q < d ? lowers_host[q] : 0.0;
        box.hi[q] = q < d ? uppers_host[q] : 1.0;
    }
    const long long total = (long long)m * d;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    candidates_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, index_base, m, d, box, out_dev);
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

int bopy_measure_peak(int what, double* tflops_out) {
    if (tflops_out == nullptr) return fail(BOPY_ERR_BAD_ARG, "tflops_out is NULL");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    const int grid = prop.multiProcessorCount * 8, block = 256;
    void* sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, 64));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best_ms = 1e30, flops = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CUDA_TRY(cudaEventRecord(e0));
        if (what == BOPY_PEAK_FP64_FMA) {
            const int iters = 512;
            peak_fma_kernel<double><<<grid, block>>>(reinterpret_cast<double*>(sink), iters, 1.0000001, 1e-9);
            flops = 2.0 * grid * block * (double)iters * 64;
        } else if (what == BOPY_PEAK_FP32_FMA) {
            const int iters = 1024;
            peak_fma_kernel<float><<<grid, block>>>(reinterpret_cast<float*>(sink), iters, 1.0000001f, 1e-9f);
            flops = 2.0 * grid * block * (double)iters * 64;
        } else if (what == BOPY_PEAK_FP64_MMA) {
            const int iters = 2048;
            const int g2 = prop.multiProcessorCount * 2;   // 16 warps per SM, 32 independent DMMAs each
            peak_dmma_kernel<<<g2, block>>>(reinterpret_cast<double*>(sink), iters, 1.0000001, 1e-9);
            flops = 2.0 * g2 * (block / 32) * (double)iters * 32 * 256;
        } else if (what == BOPY_PEAK_TF32_MMA_SYNC) {
            const int iters = 4096;
            const int g2 = prop.multiProcessorCount * 4;
            peak_tf32_mma_kernel<<<g2, block>>>(reinterpret_cast<float*>(sink), iters, 1.0000001f, 1e-9f);
            flops = 2.0 * g2 * (block / 32) * (double)iters * 16 * 1024;
        } else if (what == BOPY_PEAK_TF32_TCGEN05) {
            const int iters = 4096;
            const size_t smem = 4 * tc::TF32_TILE_BYTES + 64;
            peak_tcgen05_tf32_kernel<<<prop.multiProcessorCount, 128, smem>>>(iters);
            flops = 2.0 * prop.multiProcessorCount * (double)iters * 3 * BM * BN * 8;
        } else {
            cudaFree(sink);
            return fail(BOPY_ERR_BAD_ARG, "unknown peak id %d", what);
        }
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best_ms = std::min(best_ms, (double)ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    CUDA_TRY(cudaGetLastError());
    *tflops_out = flops / (best_ms * 1e-3) / 1e12;
    return BOPY_OK;
}

int bopy_gp_set_nan_policy(bopy_gp* gp, int policy) {
    if (gp == nullptr) return fail(BOPY_ERR_BAD_ARG, "gp handle is NULL");
    if (policy != BOPY_NAN_FIRST && policy != BOPY_NAN_SKIP) return fail(BOPY_ERR_BAD_ARG, "unknown NaN policy %d", policy);
    gp->nan_skip = policy == BOPY_NAN_SKIP ? 1 : 0;
    return BOPY_OK;
}

int bopy_multistart_refine(bopy_gp* gp, int acq, double eta, double kappa, int64_t S, const double* lowers_host,
                           const double* uppers_host, int iterations, double* xt_dev, double* xc_dev, double* fc_dev,
                           double* work_dev, void* stream) {
    int rc = check_ready(gp);
    if (rc != BOPY_OK) return rc;
    if (lowers_host == nullptr || uppers_host == nullptr || xt_dev == nullptr || xc_dev == nullptr || fc_dev == nullptr ||
        work_dev == nullptr)
        return fail(BOPY_ERR_BAD_ARG, "bopy_multistart_refine: NULL argument");
    if (S < 1 || iterations < 0) return fail(BOPY_ERR_BAD_ARG, "need S >= 1 and iterations >= 0");
    const int d = gp->d;
    // work_dev: gc (S,d) | gt (S,d) | ft (S) | alpha (S)
    double* const gc = work_dev;
    double* const gt = gc + (size_t)S * d;
    double* const ft = gt + (size_t)S * d;
    double* const alpha = ft + S;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(gp->device));
    fill_kernel<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(alpha, S, 1.0);
    CUDA_TRY(cudaGetLastError());
    for (int k = 0; k <= iterations; ++k) {
        rc = bopy_acq_value_and_grad(gp, acq, eta, kappa, xt_dev, S, ft, gt, nullptr, nullptr, stream);
        if (rc != BOPY_OK) return rc;
        rc = bopy_multistart_step(S, d, lowers_host, uppers_host, xc_dev, fc_dev, gc, xt_dev, ft, gt, alpha, k == 0 ? 1 : 0, stream);
        if (rc != BOPY_OK) return rc;
    }
    return BOPY_OK;
}

int bopy_topk_min_distance(const double* x_dev, const double* a_dev, int64_t N, int d, int k, double min_distance,
                           const double* scale_host, int64_t* idx_out_dev, double* val_out_dev, void* stream) {
    if (x_dev == nullptr || a_dev == nullptr || idx_out_dev == nullptr || val_out_dev == nullptr)
        return fail(BOPY_ERR_BAD_ARG, "bopy_topk_min_distance: NULL argument");
    if (N < 1 || d < 1 || d > MAX_D || k < 1) return fail(BOPY_ERR_BAD_ARG, "need N, k >= 1 and 1 <= d <= %d", MAX_D);
    if (!(min_distance >= 0.0)) return fail(BOPY_ERR_BAD_ARG, "min_distance must be >= 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LsParam scale;
    for (int q = 0; q < MAX_D; ++q) scale.v[q] = (scale_host != nullptr && q < d) ? scale_host[q] : 1.0;
    for (int q = 0; q < d; ++q)
        if (!(scale.v[q] > 0.0)) return fail(BOPY_ERR_BAD_ARG, "scale[%d] must be positive", q);
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = (int)std::min<long long>((N + 255) / 256, (long long)sms * 8);
    unsigned char* scratch = nullptr;
    const size_t alive_bytes = ((size_t)N + 255) / 256 * 256;
    CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&scratch), alive_bytes + (size_t)grid * sizeof(MinLoc), st));
    MinLoc* const partials = reinterpret_cast<MinLoc*>(scratch + alive_bytes);
    for (int c = 0; c < k; ++c) {
        topk_pass_kernel<<<grid, 256, 0, st>>>(x_dev, a_dev, N, d, scale, min_distance * min_distance,
                                               reinterpret_cast<const long long*>(idx_out_dev), c, scratch, partials);
        minloc_finalize_kernel<<<1, 256, 0, st>>>(partials, grid, val_out_dev + c, reinterpret_cast<long long*>(idx_out_dev) + c);
    }
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) return fail(BOPY_ERR_CUDA, "bopy_topk_min_distance launch failed: %s", cudaGetErrorString(e));
    return BOPY_OK;
}

// ---- the sharded sweep's single exchange step -----------------------------------------------------------------------

struct bopy_comm {
    NcclApi::comm_t nccl = nullptr;
    int world = 1, rank = 0, device = 0;
    MinLoc* send = nullptr;   // this rank's record
    MinLoc* recv = nullptr;   // [world]
};

int bopy_comm_unique_id(void* id_out, int64_t id_bytes) {
    if (id_out == nullptr || id_bytes < (int64_t)sizeof(NcclApi::unique_id))
        return fail(BOPY_ERR_BAD_ARG, "id_out must hold %d bytes", (int)sizeof(NcclApi::unique_id));
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(BOPY_ERR_UNSUPPORTED, "NCCL is not available: %s", api.error.c_str());
    NcclApi::unique_id id;
    const int r = api.GetUniqueId(&id);
    if (r != 0) return fail(BOPY_ERR_NCCL, "ncclGetUniqueId failed: %s", api.GetErrorString(r));
    std::memcpy(id_out, &id, sizeof(id));
    return BOPY_OK;
}

int bopy_comm_create(bopy_comm** out, const void* unique_id, int world_size, int rank, int device) {
    if (out == nullptr || unique_id == nullptr) return fail(BOPY_ERR_BAD_ARG, "out / unique_id is NULL");
    *out = nullptr;
    if (world_size < 1 || rank < 0 || rank >= world_size) return fail(BOPY_ERR_BAD_ARG, "bad rank %d of %d", rank, world_size);
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(BOPY_ERR_UNSUPPORTED, "NCCL is not available: %s", api.error.c_str());
    CUDA_TRY(cudaSetDevice(device));
    bopy_comm* c = new (std::nothrow) bopy_comm();
    if (c == nullptr) return fail(BOPY_ERR_BAD_ARG, "out of host memory");
    c->world = world_size;
    c->rank = rank;
    c->device = device;
    NcclApi::unique_id id;
    std::memcpy(&id, unique_id, sizeof(id));
    const int r = api.CommInitRank(&c->nccl, world_size, id, rank);
    if (r != 0) {
        delete c;
        return fail(BOPY_ERR_NCCL, "ncclCommInitRank failed: %s", api.GetErrorString(r));
    }
    cudaError_t e = cudaMalloc(&c->send, sizeof(MinLoc));
    if (e == cudaSuccess) e = cudaMalloc(&c->recv, (size_t)world_size * sizeof(MinLoc));
    if (e != cudaSuccess) {
        bopy_comm_destroy(c);
        return fail(BOPY_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    }
    *out = c;
    return BOPY_OK;
}

void bopy_comm_destroy(bopy_comm* comm) {
    if (comm == nullptr) return;
    cudaSetDevice(comm->device);
    if (comm->nccl != nullptr) nccl_api().CommDestroy(comm->nccl);
    cudaFree(comm->send);
    cudaFree(comm->recv);
    delete comm;
}

int bopy_minloc_allreduce(bopy_comm* comm, double* val_dev, int64_t* idx_dev, int nan_policy, void* stream) {
    if (comm == nullptr || val_dev == nullptr || idx_dev == nullptr) return fail(BOPY_ERR_BAD_ARG, "comm / val_dev / idx_dev is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(comm->device));
    minloc_pack_kernel<<<1, 1, 0, st>>>(val_dev, reinterpret_cast<const long long*>(idx_dev), comm->send);
    CUDA_TRY(cudaGetLastError());
    const int r = nccl_api().AllGather(comm->send, comm->recv, sizeof(MinLoc), NCCL_INT8, comm->nccl, st);
    if (r != 0) return fail(BOPY_ERR_NCCL, "ncclAllGather failed: %s", nccl_api().GetErrorString(r));
    minloc_gathered_kernel<<<1, 32, 0, st>>>(comm->recv, comm->world, nan_policy == BOPY_NAN_SKIP ? 1 : 0, val_dev,
                                              reinterpret_cast<long long*>(idx_dev));
    CUDA_TRY(cudaGetLastError());
    return BOPY_OK;
}

int bopy_gp_launch_info(const bopy_gp* gp, int64_t m, int* grid_out, int* launches_out, int64_t* workspace_bytes_out) {
    if (gp == nullptr) return fail(BOPY_ERR_BAD_ARG, "gp handle is NULL");
    const long long ntiles = (m + BN - 1) / BN;
    if (small_applies(gp, 0)) {
        if (grid_out) *grid_out = small_grid(gp, m);
    } else if (inv_applies(gp, m, 0, nullptr) && (gp->Winv_valid || gp->inv_mode == 1)) {
        if (grid_out) *grid_out = inv_grid(gp);
        if (launches_out) *launches_out = 1;  // probe_inv_kernel finishes the call itself
        if (workspace_bytes_out) *workspace_bytes_out = 0;
        return BOPY_OK;
    } else if (probe_applies(gp, m, 0, nullptr)) {
        if (grid_out) *grid_out = probe_plan(gp, m).grid;
    } else if (warp_applies(gp, 0)) {
        if (grid_out) *grid_out = (int)std::min<long long>(ntiles, 2LL * gp->sm_count);
        if (launches_out) *launches_out = 2;  // sweep_warp_kernel + minloc_finalize_kernel
        if (workspace_bytes_out) *workspace_bytes_out = 0;
        return BOPY_OK;
    } else if (group_applies(gp, ntiles, 0, gp->Vws)) {
        const GroupPlan pl = group_plan(gp, ntiles);
        if (grid_out) *grid_out = pl.grid;
        if (launches_out) *launches_out = 2;  // sweep_group_kernel + minloc_finalize_kernel (the control block is reset by a memset node)
        if (workspace_bytes_out) *workspace_bytes_out = (int64_t)pl.ngroups * gp->group_slots * gp->n_pad * BN * (int64_t)sizeof(double);
        return BOPY_OK;
    } else if (grid_out) {
        *grid_out = (int)std::min<long long>(ntiles, gp->sm_count);
    }
    if (launches_out) *launches_out = 2;  // sweep_kernel + minloc_finalize_kernel (argmin); 1 without argmin
    if (workspace_bytes_out) *workspace_bytes_out = (int64_t)std::min<long long>(ntiles, gp->sm_count) * gp->n_pad * BN * (int64_t)v_entry_bytes(gp);
    return BOPY_OK;
}

}  // extern "C"
