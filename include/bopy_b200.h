/*
 * bopy_b200.h -- C ABI of the B200-native GP posterior -> acquisition -> argmin path.
 *
 * The reference (tompretty/bopy) is pure Python and has no FFI / plugin registry; its seams for
 * this path are three template methods.  Each entry point below names the reference interface it
 * stands behind (paths relative to /root/reference; $SK = sklearn/gaussian_process of
 * scikit-learn 1.9.0, the library the reference delegates the arithmetic to):
 *
 *   bopy_gp_set_state      <- state left by ScipyGPSurrogate._fit     bopy/surrogate.py:87-88
 *                             ($SK/_gpr.py:349-367: X_train_, L_, alpha_, kernel_, y mean/std)
 *   bopy_gp_fit            <- ScipyGPSurrogate._fit itself for fixed hyper-parameters (gp.fit with optimizer=None)
 *   bopy_gp_predict_cov    <- ScipyGPSurrogate._predict                bopy/surrogate.py:90-91
 *                             ($SK/_gpr.py:446-473, return_cov branch)
 *   bopy_gp_predict_diag   <- np.diag(sigma) as every acquisition uses it   bopy/acquisition.py:84-85,100-101,124-125
 *   bopy_acq_eval          <- LCB._f / EI._f / POI._f                  bopy/acquisition.py:83-85, 99-106, 123-128
 *   bopy_acq_argmin        <- Optimizer._optimize() -> (x_min, f_min)  bopy/optimizer.py:65-67, 99-107
 *                             (np.argmin rules: first minimum, first NaN wins)
 *   bopy_candidates_uniform<- the candidate set an optimiser sweeps (bounds: bopy/bounds.py:36-54)
 *
 * Conventions
 *   - plain C, no CUDA / torch types: a stream is passed as `void*` (a cudaStream_t), device
 *     buffers as raw pointers.  The caller owns every buffer it passes and keeps it alive until
 *     the stream has been synchronised; the handle owns only its private packed copies/workspace.
 *   - all state and all inputs/outputs are IEEE fp64 (what the reference computes in); `dtype`
 *     selects the arithmetic of the triangular solve (fp64, or fp32 with fp64 kernel tile / mean /
 *     variance accumulation).
 *   - every call returns BOPY_OK (0) or a negative error code; the message is available from
 *     bopy_last_error() (thread-local).  Nothing throws across the boundary.  There is no CPU
 *     fallback: without a CUDA device every compute entry returns BOPY_ERR_CUDA.
 *   - calls are asynchronous on the given stream unless stated otherwise.  A handle is not
 *     thread-safe; use one handle per device (one process per GPU), and issue its calls on ONE stream
 *     (or synchronise when changing streams): launches of a handle share its solve workspace and the
 *     latency-path kernels hand thread-block roles out from a per-handle counter, so two launches of the
 *     same handle must not overlap in time.  For the same reason these calls cannot be captured into a CUDA graph.
 */
#ifndef BOPY_B200_H
#define BOPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BOPY_B200_ABI_VERSION 2

typedef struct bopy_gp bopy_gp;

enum bopy_status {
    BOPY_OK = 0,
    BOPY_ERR_BAD_ARG = -1,
    BOPY_ERR_CUDA = -2,
    BOPY_ERR_UNSUPPORTED = -3,
    BOPY_ERR_NOT_READY = -4,
    BOPY_ERR_NOT_POSITIVE_DEFINITE = -5,
    BOPY_ERR_NCCL = -6
};

enum bopy_dtype { BOPY_F64 = 0, BOPY_F32 = 1 };

/* base kernel k(r), r = |x/l - x'/l|  ($SK/kernels.py:1558-1570 RBF, :1713-1729 Matern) */
enum bopy_kernel { BOPY_KERNEL_RBF = 0, BOPY_KERNEL_MATERN12 = 1, BOPY_KERNEL_MATERN32 = 2, BOPY_KERNEL_MATERN52 = 3 };

/* acquisition functions, MINIMISATION convention (bopy/acquisition.py:14-17) */
enum bopy_acq { BOPY_ACQ_NONE = -1, BOPY_ACQ_LCB = 0, BOPY_ACQ_EI = 1, BOPY_ACQ_POI = 2 };

/* what bopy_measure_peak times */
enum bopy_peak {
    BOPY_PEAK_FP64_FMA = 0,
    BOPY_PEAK_FP32_FMA = 1,
    BOPY_PEAK_FP64_MMA = 2,
    BOPY_PEAK_TF32_MMA_SYNC = 3,
    BOPY_PEAK_TF32_TCGEN05 = 4 /* tcgen05.mma kind::tf32 M=128 N=128 K=8, operands in shared memory, accumulators in TMEM */
};

/* what the arg-min does with NaN acquisition values (posterior variance rounded to <= 0: sqrt / scipy's scale > 0 rule) */
enum bopy_nan_policy {
    BOPY_NAN_FIRST = 0, /* np.argmin: the first NaN wins -- what `np.argmin(acq(X*))` over the reference's values returns */
    BOPY_NAN_SKIP = 1   /* np.nanargmin: a NaN never wins; index -1 if every value is NaN (what an optimiser wants) */
};

typedef struct bopy_comm bopy_comm;

int bopy_abi_version(void);
const char* bopy_last_error(void);

/* Create a handle for a GP with n training points in d dimensions (1 <= d <= 32) on CUDA `device`. */
int bopy_gp_create(bopy_gp** out, int device, int dtype, int kernel, int64_t n, int d);
void bopy_gp_destroy(bopy_gp* gp);

/* Change the number of training points of a handle without reallocating, as long as ceil(n / 128) stays the same
 * (a BayesOpt loop grows the data set by one point per trial, bopy/bayes_opt.py:255-262; the Kriging believer by one
 * per batch member, bopy/acquisition.py:188-192).  The state must be installed again (bopy_gp_set_state / bopy_gp_fit).
 * BOPY_ERR_BAD_ARG if n needs another number of 128-row blocks: create a new handle then. */
int bopy_gp_resize(bopy_gp* gp, int64_t n);

/*
 * Install the fitted state.  X_dev (n,d) row-major, L_dev (n,n) row-major lower Cholesky factor of
 * K + alpha*I (entries above the diagonal are ignored), alpha_dev (n,): device pointers, fp64.
 * length_scale_host: n_ls = 1 (isotropic) or d (ARD) host doubles.  amplitude = ConstantKernel
 * value, noise_level = WhiteKernel level (enters k(x,x) only), y_mean / y_std = target
 * normalisation ($SK/_gpr.py:275-285).  Packs L into the solve layout (negated off-diagonal tiles,
 * inverted 128x128 diagonal blocks) on the device; synchronises the stream before returning.
 */
int bopy_gp_set_state(bopy_gp* gp, const double* X_dev, const double* L_dev, const double* alpha_dev,
                      const double* length_scale_host, int n_ls, double amplitude, double noise_level,
                      double y_mean, double y_std, void* stream);

/*
 * Fit with FIXED hyper-parameters entirely on the device and install the state:
 *   K = k(X, X) + (noise_level + alpha_reg) I;  L = cholesky(K);  alpha = K^-1 yn      ($SK/_gpr.py:349-367)
 * X_dev (n,d) and yn_dev (n,) are device fp64; yn is the target vector AFTER the caller's normalisation
 * ((y - y_mean) / y_std, $SK/_gpr.py:275-285: an O(n) host step kept in Python so it is bit-identical).
 * L_out_dev (n,n row-major; only the lower triangle is meaningful) and alpha_out_dev (n,) receive the factor
 * and the weights when non-NULL.  Returns BOPY_ERR_NOT_POSITIVE_DEFINITE if a pivot is not positive.
 * Synchronises the stream before returning.  Replaces bopy_gp_set_state for this handle.
 */
int bopy_gp_fit(bopy_gp* gp, const double* X_dev, const double* yn_dev, const double* length_scale_host, int n_ls,
                double amplitude, double noise_level, double alpha_reg, double y_mean, double y_std,
                double* L_out_dev, double* alpha_out_dev, void* stream);

/*
 * Grow the fitted data set by ONE point without refactorising (same hyper-parameters): a BayesOpt trial adds one
 * observation (bopy/bayes_opt.py:255-262, 239-242), the Kriging believer one fantasy per batch member
 * (bopy/acquisition.py:188-192).  X_dev (n+1,d): the old points followed by the new one; yn_dev (n+1,): ALL targets after
 * the caller's (new) normalisation.  The new row of the Cholesky factor is the latency path's forward solve of the new
 * point, L[n][:n] = L^-1 k(X, x_new), L[n][n] = sqrt(k(x,x) + noise + alpha - |L[n][:n]|^2); only the diagonal block it
 * touches is re-inverted; alpha_ is re-solved for the new targets.  Needs a state installed by bopy_gp_fit on an fp64
 * handle and n+1 within the handle's 128-row blocks (BOPY_ERR_BAD_ARG otherwise: refit).  Synchronises the stream.
 */
int bopy_gp_append(bopy_gp* gp, const double* X_dev, const double* yn_dev, double y_mean, double y_std,
                   double* alpha_out_dev, void* stream);

/* The reverse: keep only the first n_new points (the Kriging believer restores the real data when a batch is finished,
 * bopy/acquisition.py:194-197).  The factor of a leading subset is the leading block of the kept factor, so nothing is
 * refactorised: the last diagonal block is re-inverted for its new padding, alpha_ re-solved for yn_dev (n_new,), the
 * state repacked.  Same preconditions as bopy_gp_append; n_new must keep the handle's number of 128-row blocks. */
int bopy_gp_truncate(bopy_gp* gp, int64_t n_new, const double* X_dev, const double* yn_dev, double y_mean, double y_std,
                     double* alpha_out_dev, void* stream);

/*
 * Log marginal likelihood of the (normalised) targets under the given hyper-parameters, and optionally its gradient
 * with respect to the LOG hyper-parameters ($SK/_gpr.py:541-656).  Does not touch the
 * installed state.  lml_out_host: one double.  grad_out_host (nullable): n_ls + 2 doubles laid out as
 * [d/dlog amplitude, d/dlog length_scale[0..n_ls-1], d/dlog noise_level].  Synchronous.
 * BOPY_ERR_NOT_POSITIVE_DEFINITE if K + alpha I has a non-positive pivot (scikit-learn returns -inf there).
 */
int bopy_gp_lml(bopy_gp* gp, const double* X_dev, const double* yn_dev, const double* length_scale_host, int n_ls,
                double amplitude, double noise_level, double alpha_reg, double* lml_out_host, double* grad_out_host,
                void* stream);

/*
 * The fused sweep.  For candidates Xs_dev (m,d) row-major fp64 computes posterior mean and variance
 * (diagonal only, never the m x m matrix), the acquisition `acq` (eta = min(y) for EI/POI, kappa for
 * LCB) and the argmin over the m candidates.  Every output pointer may be NULL: mean_out/var_out/
 * acq_out (m,) fp64; min_val_out/min_idx_out one fp64 / int64 (index = index_base + local index;
 * first-minimum and first-NaN rules of np.argmin).
 */
int bopy_gp_posterior_acq(bopy_gp* gp, const double* Xs_dev, int64_t m, int acq, double eta, double kappa,
                          double* mean_out, double* var_out, double* acq_out,
                          int64_t index_base, double* min_val_out, int64_t* min_idx_out, void* stream);

/*
 * Small candidate sets (the reference's own calling pattern: one point per DIRECT probe, bopy/optimizer.py:95-107;
 * a few hundred for plots and fantasies) take a latency path: the forward substitution of one batch of 8-32
 * candidates is spread over the block rows of L, one CTA per block row, instead of one CTA walking all of L.
 * It serves fp64 handles for m <= max_m (default 4096, or the smaller m at which the group-mode sweep overtakes it:
 * ~2000 at n = 2048, ~450 at n = 8192; at most 148*128; 0 switches it off).  Same arithmetic as the
 * throughput path; the order of a few partial sums differs, so the two agree to rounding (~1e-15), not bit for bit.
 * Writes the limit now in force to effective_out (nullable); fp32 handles always report 0.
 */
int bopy_gp_set_latency_path(bopy_gp* gp, int64_t max_m, int64_t* effective_out);

/*
 * Inverse path: the same DIRECT probe (bopy/optimizer.py:95-107: ONE point per call, up to 20 000 calls per trial on
 * one fitted state; each is predict(return_cov=True) on a (1, d) array, bopy/surrogate.py:83-92) in one hop.  The
 * latency path above leaves a chain of n/128 dependent hops per call; a state that is probed thousands of times can pay
 * for W = L^-1 once (recursive blocked inversion on the fp64 tile kernel of bopy_gp_lml: 0.64 ms at n = 2048, 7.9 ms at 8192) and then serve every call of up to 8 candidates
 * (fewer when n_pad x 8 candidates of K* do not fit in shared memory: 2 at n = 8192) as one matrix-vector product,
 * v = W k*, spread over all SMs.  fp64 handles with n_pad <= 8192 whose latency path is on, and fp32 handles (which keep the
 * fp64 factor, inv(L_II), X / l and alpha_: their small calls are then answered in fp64); fp64 arithmetic, K* /
 * de-normalisation / acquisition shared with the other paths, results agree with them to rounding (parity bound 1e-9).
 * mode -1 (default): W is built at the 16th call of at most that many candidates on one state (the build costs what 6 chained
 * calls cost at n = 2048 and 24 at n = 8192); 1: at the first; 0: never.  Every change of the state (set_state, fit, append, truncate, resize) drops W.  The row-major factor W is
 * built from is the one bopy_gp_fit keeps, or a copy bopy_gp_set_state takes while the mode is not 0 -- switching the
 * mode on after bopy_gp_set_state takes effect at the next state.  Writes the number of candidates per call served by
 * this path for the current state to max_m_out (nullable; 0 = none).  BOPY_B200_INVERSE_PATH presets the mode.
 */
int bopy_gp_set_inverse_path(bopy_gp* gp, int mode, int64_t* max_m_out);

/*
 * Group mode of the fp64 throughput sweep.  By default every thread block owns one tile of 128 candidates and keeps that
 * tile's V = L^-1 K*^T (n_pad x 128 fp64) in a workspace slot: 148 slots, 310 MB at n = 2048 against a 126 MB L2, so half
 * of the re-reads of V go to DRAM.  In group mode `group_size` thread blocks share a tile -- the unit of work is one
 * (tile, block row of L) -- and a group keeps `slots` tiles in flight, with the first `lead` block rows of its next tile
 * interleaved with the last `lead` of the current one; tiles in flight: ceil(148 / group_size) x slots.  The arithmetic
 * and its order are those of the one-tile-per-block kernel, the results bit-identical (group_size 0).
 * group_size -2: change nothing, only report; -1: choose from n (the handle's default), 0: one tile per block, 1: one tile per block with
 * the round-1 zig-zag order of the V reads (differs from the others by rounding), >= 2: that many blocks per tile.
 * slots (2..4) / lead (0 .. n_blocks/2) < 0: choose.  effective_out (nullable) receives {group_size, slots, lead} in force.
 * fp64 DMMA handles with at least 4 block rows only; others stay at 0.  Synchronises the device.
 */
int bopy_gp_set_group_mode(bopy_gp* gp, int group_size, int slots, int lead, int* effective_out);

/*
 * The job order of group mode, for tests (host only, no device needed): the group's job-th unit of work is block row
 * *row_out of the group's *tile_seq_out-th tile.  Every (tile, row) appears exactly once, after (tile, row - 1) and
 * after (tile - 2, n_blocks - 1).
 */
int bopy_group_schedule(int64_t job, int n_blocks, int lead, int* tile_seq_out, int* row_out);

/*
 * Acquisition value AND its gradient with respect to the candidate, for the multi-start refinement behind
 * Optimizer._optimize (bopy/optimizer.py:65-67; SURVEY.md section 8f rank 3 -- the reference has no gradient code):
 *   d mean/dx = y_std sum_i alpha_i dk_i/dx,  d var/dx = -2 y_std^2 sum_i w_i dk_i/dx,  w = L^-T (L^-1 k*),
 * chained with the partials of LCB / EI / POI.  grad_out (m,d) row-major is required; acq_out, mean_out, var_out (m,)
 * may be NULL.  Rows whose posterior standard deviation is not > 0 get NaN gradients (the scale > 0 rule).
 * Runs the latency path's forward solve plus a mirrored backward solve; fp64 handles only (BOPY_ERR_UNSUPPORTED else).
 */
int bopy_acq_value_and_grad(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_dev, int64_t m,
                            double* acq_out, double* grad_out, double* mean_out, double* var_out, void* stream);

/*
 * The same for a SMALL candidate set in HOST memory -- the reference's own calling pattern, one point per DIRECT
 * probe: AcquisitionFunction.__call__(x.reshape(1, -1)), bopy/optimizer.py:96-97 -> bopy/acquisition.py:26-42.
 * Xs_host (m,d) and the requested outputs (m,) are plain host pointers (pageable or pinned), 1 <= m <= 4096; the call
 * copies in, runs the fused sweep (latency path), copies out and synchronises the stream before returning.
 * acq = BOPY_ACQ_NONE with mean/var outputs = np.diag of Surrogate.predict.
 */
int bopy_acq_eval_host(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_host, int64_t m,
                       double* acq_out_host, double* mean_out_host, double* var_out_host, void* stream);

/* Measurement aid for the latency path: called with NULL it arms tracing (the following latency-path launches record
 * %globaltimer stamps of their first batch: per block row [start, K* done, last V_J flag seen, GEMM done, diagonal solve
 * done, V_I published, -, -], nanoseconds); called with a host buffer of n_blocks * 8 int64 it synchronises, copies the
 * stamps out and disarms. */
int bopy_gp_probe_trace(bopy_gp* gp, int64_t* stamps_out_host);

/* Named views of the fused sweep. */
int bopy_gp_predict_diag(bopy_gp* gp, const double* Xs_dev, int64_t m, double* mean_out, double* var_out,
                         void* stream);
int bopy_acq_eval(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_dev, int64_t m,
                  double* acq_out, void* stream);
int bopy_acq_argmin(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_dev, int64_t m,
                    int64_t index_base, double* min_val_out, int64_t* min_idx_out, void* stream);

/*
 * The same arg-min by branch and bound (opt-in).  A rigorous lower bound of the acquisition is formed per candidate from
 * the posterior mean alone (var <= prior variance; LCB / EI / POI are monotone in the standard deviation on the side
 * that matters) at 1/40 of the cost of the full posterior; the true minimum over a strided sample of 8192 candidates is
 * the incumbent; only candidates whose bound does not exceed it -- kept in their original order -- go through the fused
 * sweep.  Index and value are those of bopy_acq_argmin bit for bit, EXCEPT that a candidate whose variance rounds to
 * <= 0 (NaN acquisition, which np.argmin would return first) may be pruned.  stats_out_host (nullable, 3 values):
 * candidates, sample size, candidates that went through the full sweep.  Synchronises the stream once internally.
 */
int bopy_acq_argmin_pruned(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_dev, int64_t m,
                           int64_t index_base, double* min_val_out, int64_t* min_idx_out, int64_t* stats_out_host,
                           void* stream);

/* Segmented arg-min: the m candidates are cut into consecutive segments of seg_len (a multiple of 128) and the
 * fused sweep returns one (value, index) per segment: seg_val_out / seg_idx_out (ceil(m / seg_len),).  This is the
 * batched multi-start primitive: one segment per start, all starts in one launch. */
int bopy_acq_segment_argmin(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_dev, int64_t m,
                            int64_t seg_len, int64_t index_base, double* seg_val_out, int64_t* seg_idx_out,
                            void* stream);

/* The segmented arg-min by branch and bound (opt-in; same caveat as bopy_acq_argmin_pruned): every 16th candidate of a
 * segment is evaluated up front, their minimum is the segment's incumbent, and only candidates whose mean-only lower bound
 * does not exceed the incumbent of THEIR segment go through the fused sweep.  Per-segment index and value equal
 * bopy_acq_segment_argmin's bit for bit.  m must be a multiple of seg_len.  stats_out_host as above. */
int bopy_acq_segment_argmin_pruned(bopy_gp* gp, int acq, double eta, double kappa, const double* Xs_dev, int64_t m,
                                   int64_t seg_len, int64_t index_base, double* seg_val_out, int64_t* seg_idx_out,
                                   int64_t* stats_out_host, void* stream);

/* Local candidate clouds around S starts (S,d): out_dev (S*P, d); row s*P is the start itself, the other P-1 rows
 * are start + U(-halfwidth, halfwidth) clipped to [lowers, uppers] (counter-based, seed). */
int bopy_candidates_around(uint64_t seed, const double* starts_dev, int64_t S, int P, int d,
                           const double* halfwidth_host, const double* lowers_host, const double* uppers_host,
                           double* out_dev, void* stream);

/* One step of the batched multi-start refinement, all S starts in lock step: monotone projected gradient with
 * Barzilai-Borwein step lengths inside [lowers, uppers].  (xt, ft, gt) = trial points with their just-evaluated
 * acquisition values / gradients (bopy_acq_value_and_grad); (xc, fc, gc) = accepted points; alpha_dev (S,) step
 * lengths.  Accepts a trial if it does not increase the value (NaN never beats a number), updates alpha, and
 * overwrites xt with the next trial points.  first != 0: the trial points are the starts themselves. */
int bopy_multistart_step(int64_t S, int d, const double* lowers_host, const double* uppers_host, double* xc_dev,
                         double* fc_dev, double* gc_dev, double* xt_dev, const double* ft_dev, const double* gt_dev,
                         double* alpha_dev, int first, void* stream);

/* out_dev[s] = Xs_dev[idx_dev[s] - index_base] for s < S (rows of d doubles). */
int bopy_gather_rows(const double* Xs_dev, int64_t m, int d, const int64_t* idx_dev, int64_t S, int64_t index_base,
                     double* out_dev, void* stream);

/* The acquisition epilogue alone, on posterior moments already on the device (mean_dev, var_dev (m,) fp64):
 * the same device code as the fused sweep's epilogue.  Serves surrogates that are not B200-native (a
 * user-defined Surrogate subclass whose predict() ran elsewhere); intended for small m. */
int bopy_acq_from_moments(int acq, double eta, double kappa, const double* mean_dev, const double* var_dev,
                          int64_t m, double* acq_out, int64_t index_base, double* min_val_out,
                          int64_t* min_idx_out, void* stream);

/* The Surrogate.predict contract: mean (m,) and the full covariance (m,m) row-major, fp64.
 * Needs an (n_pad x m) workspace inside the handle; intended for small m (plots, fantasies). */
int bopy_gp_predict_cov(bopy_gp* gp, const double* Xs_dev, int64_t m, double* mean_out, double* cov_out,
                        void* stream);

/* Counter-based uniform candidates in a box: out_dev (m,d) row-major fp64,
 * x[i][j] = lo[j] + u * (hi[j] - lo[j]), u = (splitmix64(seed + G*((index_base+i)*d + j + 1)) >> 11) * 2^-53.
 * Identical for every sharding of the index range (restated in oracle/gp_oracle.py). */
int bopy_candidates_uniform(uint64_t seed, int64_t index_base, int64_t m, int d, const double* lowers_host,
                            const double* uppers_host, double* out_dev, void* stream);

/* Register-resident FMA / MMA microbenchmark on the current device: the roofline denominator of the
 * solve.  Synchronous.  Writes TFLOP/s (2 flops per FMA). */
int bopy_measure_peak(int what, double* tflops_out);

/* Arg-min policy of this handle for NaN acquisition values (default BOPY_NAN_FIRST = np.argmin, the parity rule).
 * Applies to bopy_gp_posterior_acq / bopy_acq_argmin / bopy_acq_segment_argmin and the sweeps inside the pruned variants. */
int bopy_gp_set_nan_policy(bopy_gp* gp, int policy);

/* The whole gradient refinement of the batched multi-start behind Optimizer._optimize (bopy/optimizer.py:65-67) in ONE
 * call: iterations + 1 rounds of (bopy_acq_value_and_grad at the trial points, bopy_multistart_step), queued back to back
 * on the stream with no host round trip.  xt_dev (S,d): the starts on entry, scratch afterwards; xc_dev (S,d) / fc_dev (S,):
 * the refined points and their acquisition values; work_dev: 2*S*d + 2*S doubles of scratch. */
int bopy_multistart_refine(bopy_gp* gp, int acq, double eta, double kappa, int64_t S, const double* lowers_host,
                           const double* uppers_host, int iterations, double* xt_dev, double* xc_dev, double* fc_dev,
                           double* work_dev, void* stream);

/* One-shot batch selection on the device-resident evaluation log of a sweep (bopy/optimizer.py:186-232, 271-276; the
 * log itself: bopy/acquisition.py:200-242): the k best evaluations a_dev (N,) at points x_dev (N,d) such that any two
 * picks are at least min_distance apart after dividing coordinate j by scale_host[j] (NULL: 1).  Greedy, best first,
 * np.argmin's tie rule, NaN values never picked.  idx_out_dev / val_out_dev (k,): picks in order; -1 / 0.0 where fewer
 * than k evaluations qualify. */
int bopy_topk_min_distance(const double* x_dev, const double* a_dev, int64_t N, int d, int k, double min_distance,
                           const double* scale_host, int64_t* idx_out_dev, double* val_out_dev, void* stream);

/* The sharded sweep's ONE exchange step: candidates are cut into contiguous index ranges, one
 * process per GPU runs bopy_acq_argmin with index_base = range start, and the per-rank 16-byte (value, global index)
 * records are reduced to the global winner.  A bopy_comm wraps an ncclComm_t (NCCL is bound at run time: the libnccl.so.2
 * already loaded into the process, else the system one; BOPY_ERR_UNSUPPORTED if there is none):
 *   rank 0:      bopy_comm_unique_id(id, 128)           -> ship the 128 bytes to every rank (any out-of-band channel)
 *   every rank:  bopy_comm_create(&comm, id, world, rank, device)          (collective: ncclCommInitRank)
 *   per sweep:   bopy_minloc_allreduce(comm, val_dev, idx_dev, policy, stream)
 * bopy_minloc_allreduce replaces *val_dev / *idx_dev on every rank by the winner under np.argmin's ordering (NaN first
 * -- or never, with BOPY_NAN_SKIP --, then the smaller value, then the lower index): ncclAllGather of the records on
 * `stream` + a one-warp kernel; no host synchronisation. */
int bopy_comm_unique_id(void* id_out, int64_t id_bytes);
int bopy_comm_create(bopy_comm** out, const void* unique_id, int world_size, int rank, int device);
void bopy_comm_destroy(bopy_comm* comm);
int bopy_minloc_allreduce(bopy_comm* comm, double* val_dev, int64_t* idx_dev, int nan_policy, void* stream);

/* Introspection used by bench.py / tests: thread blocks one sweep launches and kernels per sweep. */
int bopy_gp_launch_info(const bopy_gp* gp, int64_t m, int* grid_out, int* launches_out, int64_t* workspace_bytes_out);

#ifdef __cplusplus
}
#endif
#endif /* BOPY_B200_H */
