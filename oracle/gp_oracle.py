"""CPU oracle for the bopy hot path: GP posterior -> acquisition -> argmin.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``bopy_b200/`` imports ``oracle``.

What it restates (all fp64, numpy/scipy):

* the reference's thin layer
    - ``bopy/surrogate.py:87-91``    ScipyGPSurrogate._fit/_predict  (gp.fit / gp.predict(return_cov=True))
    - ``bopy/acquisition.py:83-85``  LCB._f    mean - kappa*sqrt(diag(cov))
    - ``bopy/acquisition.py:99-109`` EI._f/_fit  -var*pdf(eta;mu,sd) + (mu-eta)*cdf(eta;mu,sd), eta=min(y)
    - ``bopy/acquisition.py:123-131`` POI._f/_fit 1 - cdf(eta;mu,sd)
    - ``bopy/optimizer.py:99-107``   the argmin a global optimiser returns (np.argmin rules)
* the third-party arithmetic the reference delegates to.  It is NOT in /root/reference:
  scikit-learn (``setup.py:13``, unpinned; this image has 1.9.0) and scipy (1.18.1).
    - sklearn ``gaussian_process/_gpr.py:446-473``  predict, ``return_cov`` branch (no clamp of
      negative variances on this branch)
    - sklearn ``gaussian_process/_gpr.py:275-285, 349-367`` y normalisation, K + alpha*I, Cholesky, alpha_
    - sklearn ``gaussian_process/kernels.py:1558-1570`` RBF (cdist(X/l, Y/l, 'sqeuclidean'), exp(-d/2))
    - sklearn ``gaussian_process/kernels.py:1713-1729`` Matern nu in {0.5, 1.5, 2.5}
    - sklearn ``gaussian_process/kernels.py:971, 1278-1282, 1418-1423`` Product / Constant / White
    - scipy ``stats/_continuous_distns.py:358-371`` norm pdf/cdf and
      ``stats/_distn_infrastructure.py:2075-2091, 2157-2175`` the ``scale > 0`` -> NaN rule

Pinning: the reference's own tests hold NO numeric vectors for this path, so
the oracle is pinned against outputs of the unmodified reference run in the authoring container
(``tools/make_golden.py`` -> ``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular
from scipy.spatial.distance import cdist
from scipy.special import ndtr

KIND_RBF = "rbf"
KIND_MATERN = "matern"

ACQ_LCB = "lcb"
ACQ_EI = "ei"
ACQ_POI = "poi"


# --------------------------------------------------------------------------------------
# kernel + fitted state
# --------------------------------------------------------------------------------------
@dataclass
class KernelSpec:
    """amplitude * base(x, x'; length_scale) [+ noise_level on k(x, x)]."""

    kind: str = KIND_RBF           # 'rbf' | 'matern'
    nu: float = 1.5                # matern only: 0.5, 1.5, 2.5
    length_scale: np.ndarray = field(default_factory=lambda: np.ones(1))  # (1,) or (d,)
    amplitude: float = 1.0         # ConstantKernel value (kernels.py:1278-1282)
    noise_level: float = 0.0       # WhiteKernel: only on k(x,x) (kernels.py:1418-1423)


@dataclass
class GPState:
    """What sklearn's fit leaves behind for predict (_gpr.py:349-367)."""

    kernel: KernelSpec
    X_train: np.ndarray            # (n, d)
    L: np.ndarray                  # (n, n) lower Cholesky factor of K + alpha I
    alpha: np.ndarray              # (n,)   (K + alpha I)^-1 y_normalised
    y_mean: float
    y_std: float


def kernel_from_sklearn(k) -> KernelSpec:
    """Flatten a fitted sklearn kernel into a KernelSpec (supports C*RBF, C*Matern, + White)."""
    from sklearn.gaussian_process import kernels as sk

    spec = KernelSpec()

    def base(kk):
        if isinstance(kk, sk.Matern):  # Matern subclasses RBF: test first
            spec.kind, spec.nu = KIND_MATERN, float(kk.nu)
            spec.length_scale = np.atleast_1d(np.asarray(kk.length_scale, dtype=np.float64))
        elif isinstance(kk, sk.RBF):
            spec.kind = KIND_RBF
            spec.length_scale = np.atleast_1d(np.asarray(kk.length_scale, dtype=np.float64))
        else:
            raise ValueError(f"unsupported base kernel {kk!r}")

    def product(kk):
        if isinstance(kk, sk.Product):
            for part in (kk.k1, kk.k2):
                if isinstance(part, sk.ConstantKernel):
                    spec.amplitude *= float(part.constant_value)
                else:
                    product(part)
        else:
            base(kk)

    if isinstance(k, sk.Sum):
        parts = [k.k1, k.k2]
        white = [p for p in parts if isinstance(p, sk.WhiteKernel)]
        rest = [p for p in parts if not isinstance(p, sk.WhiteKernel)]
        if len(white) != 1 or len(rest) != 1:
            raise ValueError(f"unsupported kernel sum {k!r}")
        spec.noise_level = float(white[0].noise_level)
        product(rest[0])
    else:
        product(k)
    return spec


def state_from_sklearn(gp) -> GPState:
    """Lift (X_train_, L_, alpha_, kernel_, y stats) out of a fitted GaussianProcessRegressor."""
    return GPState(
        kernel=kernel_from_sklearn(gp.kernel_),
        X_train=np.ascontiguousarray(gp.X_train_, dtype=np.float64),
        L=np.ascontiguousarray(gp.L_, dtype=np.float64),
        alpha=np.ascontiguousarray(gp.alpha_, dtype=np.float64).reshape(-1),
        y_mean=float(np.asarray(gp._y_train_mean).reshape(-1)[0]),
        y_std=float(np.asarray(gp._y_train_std).reshape(-1)[0]),
    )


def kernel_cross(spec: KernelSpec, Xa: np.ndarray, Xb: np.ndarray) -> np.ndarray:
    """k(Xa, Xb), (ma, mb).  Direct differences like cdist, never |x|^2+|y|^2-2xy."""
    ls = spec.length_scale
    A, B = Xa / ls, Xb / ls
    if spec.kind == KIND_RBF:                       # kernels.py:1569-1570
        K = np.exp(-0.5 * cdist(A, B, metric="sqeuclidean"))
    elif spec.kind == KIND_MATERN:                  # kernels.py:1720-1729
        r = cdist(A, B, metric="euclidean")
        if spec.nu == 0.5:
            K = np.exp(-r)
        elif spec.nu == 1.5:
            K = r * math.sqrt(3)
            K = (1.0 + K) * np.exp(-K)
        elif spec.nu == 2.5:
            K = r * math.sqrt(5)
            K = (1.0 + K + K ** 2 / 3.0) * np.exp(-K)
        else:
            raise ValueError("matern nu must be 0.5, 1.5 or 2.5")
    else:
        raise ValueError(spec.kind)
    return spec.amplitude * K                       # Product with Constant, kernels.py:971


def kernel_self_diag(spec: KernelSpec, m: int) -> np.ndarray:
    """diag k(X*, X*): normalised base kernel -> 1 (kernels.py:473-490), times amplitude, plus white."""
    return np.full(m, spec.amplitude + spec.noise_level, dtype=np.float64)


def fit_state(X: np.ndarray, y: np.ndarray, spec: KernelSpec, alpha: float, normalize_y: bool) -> GPState:
    """Fixed-hyper-parameter fit, restating _gpr.py:275-285 and :349-367."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if normalize_y:
        y_mean = float(np.mean(y))
        y_std = float(np.std(y))
        if y_std < 10 * np.finfo(np.float64).eps:   # sklearn _handle_zeros_in_scale
            y_std = 1.0
        yn = (y - y_mean) / y_std
    else:
        y_mean, y_std, yn = 0.0, 1.0, y
    K = kernel_cross(spec, X, X)
    # sklearn evaluates k(X) with pdist/squareform and fills the diagonal with exactly 1
    K = 0.5 * (K + K.T)
    K[np.diag_indices_from(K)] = spec.amplitude + spec.noise_level
    K[np.diag_indices_from(K)] += alpha
    L = cholesky(K, lower=True, check_finite=False)
    a = cho_solve((L, True), yn, check_finite=False)
    return GPState(kernel=spec, X_train=X, L=L, alpha=a, y_mean=y_mean, y_std=y_std)


# --------------------------------------------------------------------------------------
# posterior
# --------------------------------------------------------------------------------------
def posterior_full(state: GPState, Xs: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(mean (m,), cov (m,m)) exactly as _gpr.py:446-473 orders the arithmetic."""
    Kt = kernel_cross(state.kernel, Xs, state.X_train)                     # :446
    mean = Kt @ state.alpha                                                # :447
    mean = state.y_std * mean + state.y_mean                               # :450
    V = solve_triangular(state.L, Kt.T, lower=True, check_finite=False)    # :460-462
    Kss = kernel_cross(state.kernel, Xs, Xs)
    Kss = Kss.copy()
    Kss[np.diag_indices_from(Kss)] = state.kernel.amplitude + state.kernel.noise_level
    cov = Kss - V.T @ V                                                    # :466
    cov = cov * state.y_std ** 2                                           # :469
    return mean, cov


def posterior_diag(state: GPState, Xs: np.ndarray, chunk: int = 8192) -> Tuple[np.ndarray, np.ndarray]:
    """(mean (m,), var (m,)) = diagonal of posterior_full without the m x m matrix.

    var = k(x,x) - sum_i V[i,c]^2, NOT clamped (the return_cov branch does not clamp).
    """
    m = Xs.shape[0]
    mean = np.empty(m)
    var = np.empty(m)
    for s in range(0, m, chunk):
        e = min(m, s + chunk)
        Kt = kernel_cross(state.kernel, Xs[s:e], state.X_train)
        mean[s:e] = state.y_std * (Kt @ state.alpha) + state.y_mean
        V = solve_triangular(state.L, Kt.T, lower=True, check_finite=False)
        v = kernel_self_diag(state.kernel, e - s) - np.einsum("ij,ij->j", V, V)
        var[s:e] = v * state.y_std ** 2
    return mean, var


# --------------------------------------------------------------------------------------
# acquisition epilogues (minimisation convention, bopy/acquisition.py:14-17)
# --------------------------------------------------------------------------------------
_SQRT_2PI = math.sqrt(2.0 * math.pi)


def _norm_pdf(x, loc, scale):
    """scipy.stats.norm.pdf(x, loc, scale): exp(-z^2/2)/sqrt(2pi)/scale; NaN unless scale > 0."""
    with np.errstate(all="ignore"):
        z = (x - loc) / scale
        out = np.exp(-z * z / 2.0) / _SQRT_2PI / scale
    bad = ~(scale > 0) | np.isnan(loc)
    return np.where(bad, np.nan, out)


def _norm_cdf(x, loc, scale):
    """scipy.stats.norm.cdf(x, loc, scale) = ndtr((x-loc)/scale); NaN unless scale > 0."""
    with np.errstate(all="ignore"):
        z = (x - loc) / scale
        out = ndtr(z)
    bad = ~(scale > 0) | np.isnan(loc)
    return np.where(bad, np.nan, out)


def acquisition(kind: str, mean: np.ndarray, var: np.ndarray, eta: float = 0.0, kappa: float = 2.0) -> np.ndarray:
    """LCB / EI / POI from (mean, var) following bopy/acquisition.py:83-85, 99-106, 123-128."""
    with np.errstate(invalid="ignore"):
        std = np.sqrt(var)                           # negative var -> NaN, as in the reference
    if kind == ACQ_LCB:
        return mean - kappa * std
    if kind == ACQ_EI:
        return -var * _norm_pdf(eta, mean, std) + (mean - eta) * _norm_cdf(eta, mean, std)
    if kind == ACQ_POI:
        return 1 - _norm_cdf(eta, mean, std)
    raise ValueError(kind)


def argmin_first(values: np.ndarray) -> Tuple[int, float]:
    """np.argmin rules: first occurrence of the minimum; the first NaN wins if any."""
    i = int(np.argmin(values))
    return i, float(values[i])


def acquisition_sweep(state: GPState, kind: str, Xs: np.ndarray, eta: float = 0.0, kappa: float = 2.0,
                      chunk: int = 8192):
    """mean, var, acq, (argmin index, value) over a candidate set (diag-only path)."""
    mean, var = posterior_diag(state, Xs, chunk)
    a = acquisition(kind, mean, var, eta, kappa)
    return mean, var, a, argmin_first(a)


# --------------------------------------------------------------------------------------
# gradient of the acquisition with respect to the candidate (for the multi-start refinement)
# --------------------------------------------------------------------------------------
def kernel_cross_grad_factor(spec: KernelSpec, Xa: np.ndarray, Xb: np.ndarray) -> np.ndarray:
    """kd (ma, mb) with  d k(xa, xb) / d xa_q = kd * (xa_q - xb_q) / l_q^2.

    From the closed forms kernel_cross() evaluates (kernels.py:1569-1570, 1720-1729), r = |xa/l - xb/l|:
    RBF -k; Matern 1/2 -amp e^-r / r (0 at r = 0, where the kernel has a kink); 3/2 -3 amp e^(-sqrt3 r);
    5/2 -(5/3) (1 + sqrt5 r) amp e^(-sqrt5 r).  The reference has no gradient code: this extension is pinned by
    finite differences of the (pinned) acquisition above, tests/test_oracle_gradient.py."""
    ls = spec.length_scale
    A, B = Xa / ls, Xb / ls
    if spec.kind == KIND_RBF:
        return -spec.amplitude * np.exp(-0.5 * cdist(A, B, metric="sqeuclidean"))
    r = cdist(A, B, metric="euclidean")
    if spec.nu == 0.5:
        with np.errstate(divide="ignore", invalid="ignore"):
            out = -spec.amplitude * np.exp(-r) / r
        return np.where(r > 0, out, 0.0)
    if spec.nu == 1.5:
        return -3.0 * spec.amplitude * np.exp(-math.sqrt(3) * r)
    if spec.nu == 2.5:
        k = math.sqrt(5) * r
        return -(5.0 / 3.0) * (1.0 + k) * spec.amplitude * np.exp(-k)
    raise ValueError("matern nu must be 0.5, 1.5 or 2.5")


def acquisition_partials(kind: str, mean, var, eta: float = 0.0, kappa: float = 2.0):
    """(d acq / d mean, d acq / d var) of LCB / EI / POI as acquisition() defines them; NaN unless sqrt(var) > 0."""
    with np.errstate(all="ignore"):
        sd = np.sqrt(var)
        z = (eta - mean) / sd
        pdf = np.exp(-z * z / 2.0) / _SQRT_2PI
        if kind == ACQ_LCB:
            dm, dv = np.ones_like(mean), -kappa / (2.0 * sd)
        elif kind == ACQ_EI:           # a = -sd pdf(z) - (eta - mean) cdf(z)
            dm, dv = ndtr(z), -pdf / (2.0 * sd)
        elif kind == ACQ_POI:          # a = 1 - cdf(z)
            dm, dv = pdf / sd, pdf * z / (2.0 * var)
        else:
            raise ValueError(kind)
    bad = ~(sd > 0)
    return np.where(bad, np.nan, dm), np.where(bad, np.nan, dv)


def acquisition_value_and_grad(state: GPState, kind: str, Xs: np.ndarray, eta: float = 0.0, kappa: float = 2.0):
    """acq (m,), d acq / d x (m, d), plus (mean, var):
        d mean / dx = y_std   * sum_i alpha_i dk_i/dx
        d var  / dx = -2 y_var * sum_i w_i dk_i/dx,   w = K^-1 k* = L^-T (L^-1 k*)."""
    spec = state.kernel
    ls = np.broadcast_to(spec.length_scale, (state.X_train.shape[1],))
    Kt = kernel_cross(spec, Xs, state.X_train)
    mean = state.y_std * (Kt @ state.alpha) + state.y_mean
    V = solve_triangular(state.L, Kt.T, lower=True, check_finite=False)
    var = (kernel_self_diag(spec, Xs.shape[0]) - np.einsum("ij,ij->j", V, V)) * state.y_std ** 2
    W = solve_triangular(state.L.T, V, lower=False, check_finite=False)           # (n, m)
    kd = kernel_cross_grad_factor(spec, Xs, state.X_train)                         # (m, n)
    diff = (Xs[:, None, :] - state.X_train[None, :, :]) / (ls * ls)                # (m, n, d)
    gm = np.einsum("mn,mnd->md", kd * state.alpha[None, :], diff)
    gv = np.einsum("mn,mnd->md", kd * W.T, diff)
    dm, dv = acquisition_partials(kind, mean, var, eta, kappa)
    grad = dm[:, None] * (state.y_std * gm) + dv[:, None] * (-2.0 * state.y_std ** 2 * gv)
    return acquisition(kind, mean, var, eta, kappa), grad, mean, var


def multistart_step(lowers, uppers, xc, fc, gc, xt, ft, gt, alpha, first):
    """One lock-step step of the batched multi-start refinement, restating multistart_step_kernel
    (bopy_b200/csrc/aux_kernels.cuh) operation by operation.  Arrays are updated in place."""
    lowers, uppers = np.asarray(lowers, dtype=np.float64), np.asarray(uppers, dtype=np.float64)
    S, d = xt.shape
    for s in range(S):
        a = alpha[s]
        f_new, f_cur = ft[s], fc[s]
        if first:
            accept = True
            gmax, span = 0.0, uppers[0] - lowers[0]
            for q in range(d):
                if not np.isnan(gt[s, q]):
                    gmax = max(gmax, abs(gt[s, q]))
                span = min(span, uppers[q] - lowers[q])
            a = 0.1 * span / gmax if gmax > 0.0 else 1.0
        else:
            new_nan, cur_nan = np.isnan(f_new), np.isnan(f_cur)
            accept = (not new_nan) and (cur_nan or f_new <= f_cur)
            if accept:
                ss = sy = 0.0
                for q in range(d):
                    sq, yq = xt[s, q] - xc[s, q], gt[s, q] - gc[s, q]
                    ss = ss + sq * sq
                    sy = sy + sq * yq
                a = ss / sy if sy > 0.0 else 4.0 * a
                a = min(max(a, 1e-12), 1e12)
            else:
                a = 0.25 * a
        if accept:
            fc[s] = f_new
            xc[s] = xt[s]
            gc[s] = gt[s]
        alpha[s] = a
        for q in range(d):
            gq = gc[s, q]
            step = a * gq if not np.isnan(gq) else 0.0
            xt[s, q] = min(max(xc[s, q] - step, lowers[q]), uppers[q])


# --------------------------------------------------------------------------------------
# the reference's own call sequence, with its m x m covariance (used as the CPU baseline)
# --------------------------------------------------------------------------------------
def reference_style_acquisition(state: GPState, kind: str, Xs: np.ndarray, eta: float, kappa: float = 2.0,
                                chunk: int = 64) -> np.ndarray:
    """What bopy does per call: predict -> full cov -> np.diag -> epilogue, in chunks of `chunk`
    candidates (bopy/optimizer.py:96-97 calls it with chunk = 1)."""
    out = np.empty(Xs.shape[0])
    for s in range(0, Xs.shape[0], chunk):
        e = min(Xs.shape[0], s + chunk)
        mean, cov = posterior_full(state, Xs[s:e])
        out[s:e] = acquisition(kind, mean, np.diag(cov), eta, kappa)
    return out


# --------------------------------------------------------------------------------------
# counter-based candidate generator (restates bopy_b200/csrc: bopy_candidates_uniform)
# --------------------------------------------------------------------------------------
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def candidates_uniform(seed: int, index_base: int, m: int, lowers, uppers) -> np.ndarray:
    """x[i, j] = lo_j + u * (hi_j - lo_j), u = (splitmix64(seed + G*((base+i)*d + j + 1)) >> 11) * 2^-53."""
    lowers = np.asarray(lowers, dtype=np.float64)
    uppers = np.asarray(uppers, dtype=np.float64)
    d = lowers.shape[0]
    with np.errstate(over="ignore"):
        ctr = (np.arange(m, dtype=np.uint64)[:, None] + np.uint64(index_base)) * np.uint64(d) \
            + np.arange(d, dtype=np.uint64)[None, :] + np.uint64(1)
        z = np.uint64(seed) + _GOLDEN * ctr
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lowers[None, :] + u * (uppers - lowers)[None, :]
