"""Tolerance conventions of the parity tests (also stated in DESIGN.md, section "Parity").

fp64 path (north star: posterior mean/var within 1e-9 relative):
    |mean - mean_ref| <= 1e-9 * |mean_ref| + 1e-9 * y_std
    |var  - var_ref | <= 1e-9 * |var_ref|  + 1e-11 * prior_var         prior_var = (amplitude + noise) * y_std^2
  The absolute terms are the floor below which "relative" is undefined for this path: the mean is a
  cancelling sum of n terms alpha_i k_i of size up to |alpha|_max, and var = prior - sum v^2 cancels to
  ~alpha_reg * prior at training points, where the reference's own two formulations (full-cov gemm vs
  diag einsum) already differ by ~1e-12 * prior.
fp32 mode (north star: 1e-4).  Only L_IJ and V are stored/multiplied in fp32; K*, the mean, the residual
accumulation across 128-column blocks, the diagonal solve and sum v^2 are fp64.  What remains is the first-order
sensitivity of sum v^2 to rounding L and V to fp32, S = 2 eps32 |w|^T |L| |v| with w = K^-1 k* (1e-6..4e-6 of the
prior on the BASELINE configs):
    |mean - mean_ref| : as fp64
    |var  - var_ref | <= 1e-4 * max(|var_ref|, 1e-1 * prior_var)
  i.e. 1e-4 relative wherever the posterior variance is at least a tenth of the prior, 1e-5 * prior below that.
Arg-min: index identical, or a *stated tie*: the reference's own acquisition values at the two indices
differ by no more than the acquisition's error bound implied by the var/mean tolerances above
(tie_tol, relative to the spread of the reference acquisition values).
"""
import numpy as np

TOL = {
    "f64": dict(mean_rtol=1e-9, mean_atol=1e-9, var_rtol=1e-9, var_atol=1e-11, var_floor=0.0, tie=1e-9),
    "f32": dict(mean_rtol=1e-9, mean_atol=1e-9, var_rtol=1e-4, var_atol=0.0, var_floor=1e-1, tie=1e-4),
}


def prior_var(state):
    return (state.kernel.amplitude + state.kernel.noise_level) * state.y_std ** 2


def check_mean(mean, ref, state, dtype):
    t = TOL[dtype]
    err = np.abs(mean - ref)
    bound = t["mean_rtol"] * np.abs(ref) + t["mean_atol"] * state.y_std
    return err, bound


def check_var(var, ref, state, dtype):
    t = TOL[dtype]
    pv = prior_var(state)
    err = np.abs(var - ref)
    bound = t["var_rtol"] * np.maximum(np.abs(ref), t["var_floor"] * pv) + t["var_atol"] * pv
    return err, bound


def is_stated_tie(acq_ref, idx_ref, idx_got, dtype):
    """idx_got is acceptable if the REFERENCE values at idx_ref and idx_got are within tie tolerance."""
    if idx_ref == idx_got:
        return True
    a, b = acq_ref[idx_ref], acq_ref[idx_got]
    if np.isnan(a) or np.isnan(b):
        return bool(np.isnan(a) and np.isnan(b))
    finite = acq_ref[np.isfinite(acq_ref)]
    spread = float(np.max(finite) - np.min(finite)) if finite.size else 1.0
    return abs(a - b) <= TOL[dtype]["tie"] * max(spread, abs(a))
