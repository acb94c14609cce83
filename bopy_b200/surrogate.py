"""Surrogate models: the `fit` / `predict` interface of bopy/surrogate.py with the posterior on a B200.

`Surrogate` keeps the reference's template (validate -> store references -> `_fit` -> mark fitted;
validate -> `_predict`, bopy/surrogate.py:26-69).  `B200GPSurrogate` takes the same constructor
argument as the reference's `ScipyGPSurrogate` (a scikit-learn `GaussianProcessRegressor`), lets
scikit-learn do the fit exactly as the reference does (bopy/surrogate.py:87-88; the fit is not on
the hot path), then hands the fitted state to the C ABI.  Every prediction afterwards --
`predict` (mean + full covariance), `predict_diag`, the fused acquisition values and the fused
acquisition argmin -- runs in the sm_100a kernels.  numpy in, numpy out.
"""
from abc import ABC, abstractmethod
from typing import Tuple

import numpy as np

from . import _native
from .kernel_spec import flatten_sklearn_kernel
from .mixin import FittableMixin


class Surrogate(FittableMixin, ABC):
    """A probabilistic stand-in for the objective: `fit(x, y)` then `predict(x) -> (mean, cov)`."""

    def __init__(self):
        super().__init__()
        self.has_been_fitted = False
        self.n_dimensions = -1
        self.x = np.array([])
        self.y = np.array([])

    def fit(self, x: np.ndarray, y: np.ndarray) -> None:
        """x: (n_samples, n_dimensions), y: (n_samples,).  Keeps references to both."""
        self._validate_ok_for_fitting(x, y)
        self.x, self.y = x, y
        self._fit(x, y)
        self._confirm_fit()

    def predict(self, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """x: (n_samples, n_dimensions) -> mean (n_samples,), covariance (n_samples, n_samples)."""
        self._validate_ok_for_predicting(x)
        return self._predict(x)

    @abstractmethod
    def _fit(self, x: np.ndarray, y: np.ndarray) -> None:
        ...

    @abstractmethod
    def _predict(self, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        ...


class B200GPSurrogate(Surrogate):
    """scikit-learn GP regressor whose posterior is evaluated by the B200 kernels.

    Parameters
    ----------
    gp : sklearn.gaussian_process.GaussianProcessRegressor
        Same object the reference's ScipyGPSurrogate wraps (bopy/surrogate.py:82-85).
    dtype : 'f64' | 'f32'
        Arithmetic of the triangular solve.  'f64' matches the reference to ~1e-12; 'f32' keeps the
        kernel tile, the mean and the variance reduction in fp64 and solves in fp32.
    device : int | str | torch.device | None
        CUDA device (default: current).
    device_fit : 'auto' | bool
        Where the fit runs.
        'auto'  fixed hyper-parameters (`gp.optimizer is None`): Gram matrix, Cholesky factorisation and alpha_ on the
                device (`bopy_gp_fit`), nothing n x n crosses PCIe; otherwise scikit-learn optimises the
                hyper-parameters on the host exactly as the reference does and the fitted state is uploaded.
        True    always on the device: with an optimizer, scikit-learn's own optimiser (L-BFGS-B, restarts) drives the
                log marginal likelihood and its gradient evaluated by `bopy_gp_lml` -- every O(n^3) step on the GPU.
                The optimiser path is sensitive to rounding, so the hyper-parameters can differ from the host
                route's in the last digits (or pick another local optimum where scikit-learn's would, too).
        False   always the host route.
    latency_max_m : int | None
        Calls with at most this many candidates take the latency path of the library (`bopy_gp_set_latency_path`:
        the solve of a small batch is spread over the block rows of L, ~n/128 hops of a few microseconds, instead
        of one thread block walking all of L).  None keeps the library default (4096, less where the group-mode sweep overtakes it earlier); 0 switches it off.
    inverse_path : 'auto' | bool
        Calls of a handful of candidates (DIRECT probes the acquisition ONE point per call, thousands of times per
        trial, `bopy/optimizer.py:95-107`) as one matrix-vector product with W = L^-1 (`bopy_gp_set_inverse_path`).
        'auto' (the library default) builds W at the 16th such call on one fitted state (the build costs what 6 chained
        calls cost at n = 2048, 24 at n = 8192), True at the first, False never.
        The paths agree to rounding, not bit for bit: with False a candidate's value never depends on how many
        calls came before it.  dtype='f32' surrogates are served too, in fp64 (the handle keeps the fp64 factor).
    """

    def __init__(self, gp, dtype: str = "f64", device=None, device_fit="auto", latency_max_m=None, inverse_path="auto"):
        super().__init__()
        if dtype not in ("f64", "f32"):
            raise ValueError("dtype must be 'f64' or 'f32'")
        self.gp = gp
        self.dtype = dtype
        self.device = device
        self.device_fit = device_fit
        self.latency_max_m = latency_max_m
        if inverse_path not in ("auto", True, False):
            raise ValueError("inverse_path must be 'auto', True or False")
        self.inverse_path = inverse_path
        self.native = None          # _native.NativeGP once fitted
        self.kernel_spec = None
        self.fitted_on_device = False

    # -- fit ---------------------------------------------------------------------------------------
    def _fit(self, x: np.ndarray, y: np.ndarray) -> None:
        if self._device_fit_applies():
            self._fit_on_device(x, y)
        else:
            self.gp.fit(x, y)          # host fit incl. hyper-parameter optimisation, bopy/surrogate.py:87-88
            self.load_fitted_state()

    def _device_fit_applies(self) -> bool:
        if self.device_fit is False:
            return False
        fixed = getattr(self.gp, "optimizer", None) is None
        scalar_alpha = np.ndim(self.gp.alpha) == 0
        if self.device_fit is True:
            if not scalar_alpha:
                raise ValueError("device_fit=True needs a scalar alpha")
            return True
        return fixed and scalar_alpha

    def _optimise_on_device(self, kernel_, native, Xd, yd):
        """$SK/_gpr.py:299-344 with the objective on the device: maximise the log marginal likelihood over theta."""
        from operator import itemgetter

        from sklearn.utils import check_random_state

        from .kernel_spec import theta_gradient
        gp = self.gp
        alpha_reg = float(gp.alpha)

        def obj_func(theta, eval_gradient=True):
            k = kernel_.clone_with_theta(theta)
            flat = flatten_sklearn_kernel(k)
            try:
                lml, grad = native.lml(Xd, yd, flat.length_scale, amplitude=flat.amplitude,
                                       noise_level=flat.noise_level, alpha_reg=alpha_reg, want_grad=eval_gradient)
            except np.linalg.LinAlgError:       # scikit-learn: -inf likelihood, zero gradient
                return (np.inf, np.zeros_like(theta)) if eval_gradient else np.inf
            if eval_gradient:
                return -lml, -theta_gradient(k, grad)
            return -lml

        gp._rng = check_random_state(gp.random_state)
        optima = [gp._constrained_optimization(obj_func, kernel_.theta, kernel_.bounds)]
        if gp.n_restarts_optimizer > 0:
            if not np.isfinite(kernel_.bounds).all():
                raise ValueError("Multiple optimizer restarts (n_restarts_optimizer>0) requires that all bounds "
                                 "are finite.")
            bounds = kernel_.bounds
            for _ in range(gp.n_restarts_optimizer):
                theta0 = gp._rng.uniform(bounds[:, 0], bounds[:, 1])
                optima.append(gp._constrained_optimization(obj_func, theta0, bounds))
        values = list(map(itemgetter(1), optima))
        kernel_.theta = optima[int(np.argmin(values))][0]
        kernel_._check_bounds_params()
        gp.log_marginal_likelihood_value_ = -np.min(values)

    def _native_for(self, n: int, d: int, kernel: str):
        if self.native is not None and self.native.d == d and self.native.kernel == kernel and self.native.resize(n):
            return self.native        # same number of 128-row blocks: workspaces are reused (one more point per trial)
        if (self.native is None or self.native.n != n or self.native.d != d or self.native.kernel != kernel):
            if self.native is not None:
                self.native.close()
            self.native = _native.NativeGP(n, d, kernel=kernel, dtype=self.dtype, device=self.device)
            if getattr(self, "latency_max_m", None) is not None:
                self.native.set_latency_path(self.latency_max_m)
            if getattr(self, "inverse_path", "auto") != "auto":
                self.native.set_inverse_path(1 if self.inverse_path else 0)
        return self.native

    def _fit_on_device(self, x: np.ndarray, y: np.ndarray) -> None:
        """gp.fit(x, y) with optimizer=None, restated: $SK/_gpr.py:275-285 on the host (O(n)), :349-367 on the GPU."""
        from sklearn.base import clone
        from sklearn.gaussian_process.kernels import RBF, ConstantKernel
        gp = self.gp
        kernel = gp.kernel if gp.kernel is not None else \
            ConstantKernel(1.0, constant_value_bounds="fixed") * RBF(1.0, length_scale_bounds="fixed")
        kernel_ = clone(kernel)
        spec = flatten_sklearn_kernel(kernel_)      # raises for kernels the device does not implement
        X = np.ascontiguousarray(x, dtype=np.float64)
        yv = np.asarray(y, dtype=np.float64)
        if gp.normalize_y:
            y_mean = np.mean(yv, axis=0)
            y_std = np.std(yv, axis=0)
            if y_std < 10 * np.finfo(np.float64).eps:   # sklearn's _handle_zeros_in_scale
                y_std = 1.0
            yn = (yv - y_mean) / y_std
        else:
            y_mean, y_std, yn = 0.0, 1.0, yv
        n, d = X.shape
        if self._try_append(X, yn, y_mean, y_std, spec, kernel_):
            return
        native = self._native_for(n, d, spec.kernel)
        if gp.optimizer is not None and kernel_.n_dims > 0:
            Xd, yd = native._dev64(X, (n, d)), native._dev64(yn, (n,))
            self._optimise_on_device(kernel_, native, Xd, yd)
            spec = flatten_sklearn_kernel(kernel_)
        alpha, _ = native.fit(X, yn, spec.length_scale, amplitude=spec.amplitude, noise_level=spec.noise_level,
                              alpha_reg=float(gp.alpha), y_mean=float(y_mean), y_std=float(y_std))
        # leave the scikit-learn object in the state its own fit would leave (minus the n x n factor, which stays
        # on the device; `export_factor()` fetches it)
        gp.kernel_ = kernel_
        gp.X_train_ = np.copy(X) if gp.copy_X_train else X
        gp.y_train_ = np.copy(yn) if gp.copy_X_train else yn
        gp._y_train_mean, gp._y_train_std = y_mean, y_std
        gp.alpha_ = alpha.cpu().numpy()
        self.kernel_spec = spec
        self.fitted_on_device = True
        self._append_key = self._spec_key(spec) if gp.optimizer is None else None

    def _spec_key(self, spec):
        return (spec.kernel, tuple(np.ravel(spec.length_scale).tolist()), float(spec.amplitude), float(spec.noise_level),
                float(self.gp.alpha), bool(self.gp.normalize_y))

    def _try_append(self, X, yn, y_mean, y_std, spec, kernel_) -> bool:
        """A refit whose data are the previous data plus ONE new last point, with the same fixed hyper-parameters, is a
        one-row extension of the factor (`bopy_gp_append`): the BayesOpt loop's per-trial refit
        (bopy/bayes_opt.py:239-242) and the Kriging believer's per-member refit (bopy/acquisition.py:188-192)."""
        gp, native = self.gp, self.native
        prev = getattr(gp, "X_train_", None)
        if (native is None or not self.fitted_on_device or gp.optimizer is not None or prev is None
                or getattr(self, "_append_key", None) != self._spec_key(spec) or native.n != prev.shape[0]
                or X.shape[1] != prev.shape[1] or not native.gradient_capable()
                or (X.shape[0] + 127) // 128 != (native.n + 127) // 128):
            return False
        grow = X.shape[0] == prev.shape[0] + 1 and np.array_equal(X[:-1], prev)
        # a leading subset of the previous data (the Kriging believer's restore): the factor is the leading block
        shrink = X.shape[0] < prev.shape[0] and np.array_equal(X, prev[:X.shape[0]])
        if not (grow or shrink):
            return False
        try:
            alpha = (native.append if grow else native.truncate)(X, yn, y_mean=float(y_mean), y_std=float(y_std))
        except np.linalg.LinAlgError:
            raise
        except _native.NativeLibraryError:
            return False          # the generic path refits from scratch
        gp.kernel_ = kernel_
        gp.X_train_ = np.copy(X) if gp.copy_X_train else X
        gp.y_train_ = np.copy(yn) if gp.copy_X_train else yn
        gp._y_train_mean, gp._y_train_std = y_mean, y_std
        gp.alpha_ = alpha.cpu().numpy()
        self.kernel_spec = spec
        if grow:
            self.appended_rows = getattr(self, "appended_rows", 0) + 1
        else:
            self.truncations = getattr(self, "truncations", 0) + 1
        return True

    def export_factor(self) -> np.ndarray:
        """The lower Cholesky factor L_ (n, n) as numpy: refits on the device with the factor exported."""
        spec, gp = self.kernel_spec, self.gp
        _, L = self.native.fit(gp.X_train_, gp.y_train_, spec.length_scale, amplitude=spec.amplitude,
                               noise_level=spec.noise_level, alpha_reg=float(gp.alpha),
                               y_mean=float(np.ravel(gp._y_train_mean)[0]), y_std=float(np.ravel(gp._y_train_std)[0]),
                               want_factor=True)
        return np.tril(L.cpu().numpy())

    def load_fitted_state(self) -> None:
        """(Re)install the state of `self.gp` -- X_train_, L_, alpha_, kernel_, y mean/std -- on the device."""
        gp = self.gp
        spec = flatten_sklearn_kernel(gp.kernel_)
        X = np.ascontiguousarray(gp.X_train_, dtype=np.float64)
        n, d = X.shape
        if np.ndim(gp.alpha_) != 1:
            raise ValueError("multi-target GPs are not supported")
        self._native_for(n, d, spec.kernel).set_state(
            X, gp.L_, gp.alpha_, spec.length_scale, amplitude=spec.amplitude, noise_level=spec.noise_level,
            y_mean=float(np.ravel(gp._y_train_mean)[0]), y_std=float(np.ravel(gp._y_train_std)[0]))
        self.kernel_spec = spec
        self.fitted_on_device = False

    # -- the reference contract --------------------------------------------------------------------
    def _predict(self, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        if isinstance(x, np.ndarray) and x.shape[0] == 1:
            # ONE point (a DIRECT objective built on predict, the Kriging believer's bopy/acquisition.py:189): the 1 x 1
            # covariance is the posterior variance -- one host-buffer call on the latency / inverse path
            _, mean, var = self.native.eval_host(np.ascontiguousarray(x, dtype=np.float64), want_acq=False, want_mean=True,
                                                 want_var=True)
            return mean, var.reshape(1, 1)
        xs = self.native.candidates(x)
        mean, cov = self.native.predict_cov(xs)
        return mean.cpu().numpy(), cov.cpu().numpy()

    # -- additive, diagonal-only and fused entry points ---------------------------------------------
    def _small_host_call(self, x):
        """x as a C-contiguous float64 array if it qualifies for the one-call host-buffer entry, else None."""
        if isinstance(x, np.ndarray) and x.ndim == 2 and 0 < x.shape[0] <= self.native.HOST_CALL_MAX_M \
                and x.shape[1] == self.native.d:
            return np.ascontiguousarray(x, dtype=np.float64)
        return None

    def predict_diag(self, x) -> Tuple[np.ndarray, np.ndarray]:
        """mean (m,), var (m,) = diagonal of `predict` without the m x m matrix."""
        self._validate_ok_for_predicting(x)
        xh = self._small_host_call(x)
        if xh is not None:
            _, mean, var = self.native.eval_host(xh, want_acq=False, want_mean=True, want_var=True)
            return mean, var
        out = self.native.sweep(self.native.candidates(x), want_mean=True, want_var=True)
        return out["mean"].cpu().numpy(), out["var"].cpu().numpy()

    def acquisition_values(self, kind: str, x, eta: float = 0.0, kappa: float = 2.0) -> np.ndarray:
        """Fused posterior -> acquisition; (m,) numpy."""
        xh = self._small_host_call(x)
        if xh is not None:       # one point per DIRECT probe (bopy/optimizer.py:96-97): one native call, no torch ops
            return self.native.eval_host(xh, kind, eta=eta, kappa=kappa)[0]
        out = self.native.sweep(self.native.candidates(x), acq=kind, eta=eta, kappa=kappa, want_acq=True)
        return out["acq"].cpu().numpy()

    def acquisition_argmin(self, kind: str, x, eta: float = 0.0, kappa: float = 2.0, index_base: int = 0,
                           prune: bool = False, nan_policy: str = "first", on_device: bool = False):
        """Fused posterior -> acquisition -> argmin over the rows of x (numpy or device tensor).

        Returns (index, value).  nan_policy='first': np.argmin's rules (first minimum; a NaN value -- posterior variance
        rounded to <= 0 -- wins, the first one): what `np.argmin(acq(X*))` over the reference's values returns, and the
        parity rule.  nan_policy='skip': np.nanargmin's (a NaN never wins; index -1 if everything is NaN): what an
        optimiser wants, so that it never proposes a point on top of a training point because its variance cancelled.
        prune=True: branch and bound (`bopy_acq_argmin_pruned`): a lower bound from the posterior mean discards most
        candidates before the full sweep; same index and value, except that candidates with a NaN acquisition may be
        skipped (with nan_policy='skip' the two routes agree).  `last_prune_stats` holds how many were fully evaluated.
        on_device=True: (index, value) stay on the device as (1,) tensors (no synchronisation)."""
        self.native.set_nan_policy(nan_policy)
        try:
            if prune:
                minv, mini, self.last_prune_stats = self.native.argmin_pruned(self.native.candidates(x), kind, eta=eta,
                                                                              kappa=kappa, index_base=index_base)
            else:
                out = self.native.sweep(self.native.candidates(x), acq=kind, eta=eta, kappa=kappa, want_min=True,
                                        index_base=index_base)
                minv, mini = out["min_val"], out["min_idx"]
        finally:
            self.native.set_nan_policy("first")
        if on_device:
            return mini, minv
        import torch
        both = torch.stack([mini[0], minv.view(torch.int64)[0]]).cpu().numpy()       # one device-to-host copy
        return int(both[0]), float(both[1:].view(np.float64)[0])


    def supports_gradient(self) -> bool:
        """True if `acquisition_value_and_grad` is available (fp64 solve, n <= 148 * 128)."""
        return self.native is not None and self.native.gradient_capable()

    def acquisition_value_and_grad(self, kind: str, x, eta: float = 0.0, kappa: float = 2.0):
        """Acquisition values (m,) and their gradient with respect to the rows of x (m, d), numpy
        (`bopy_acq_value_and_grad`: the forward solve plus a mirrored backward solve on the device)."""
        val, grad, _, _ = self.native.value_and_grad(self.native.candidates(x), kind, eta=eta, kappa=kappa)
        return val.cpu().numpy(), grad.cpu().numpy()

    def acquisition_segment_argmin(self, kind: str, xs, seg_len: int, eta: float = 0.0, kappa: float = 2.0,
                                   index_base: int = 0, prune: bool = False, nan_policy: str = "first"):
        """Per-segment fused argmin over consecutive segments of `seg_len` rows (a multiple of 128) of the device
        tensor / array `xs`: (values (nseg,), indices (nseg,)) as device tensors.  One launch for all segments.
        prune=True: branch and bound per segment (`bopy_acq_segment_argmin_pruned`), same results.
        nan_policy: as in `acquisition_argmin` (a segment of NaNs only has index -1 under 'skip')."""
        self.native.set_nan_policy(nan_policy)
        try:
            if prune and len(xs) % seg_len == 0:
                vals, idxs, self.last_prune_stats = self.native.segment_argmin_pruned(
                    self.native.candidates(xs), seg_len, kind, eta=eta, kappa=kappa, index_base=index_base)
                return vals, idxs
            return self.native.segment_argmin(self.native.candidates(xs), seg_len, kind, eta=eta, kappa=kappa,
                                              index_base=index_base)
        finally:
            self.native.set_nan_policy("first")


class GPyGPSurrogate(B200GPSurrogate):
    """GPy `GPRegression` surrogate whose posterior is evaluated by the B200 kernels (reference:
    bopy/surrogate.py:94-146).

    Same constructor and fit flow as the reference: `gp_initializer(x, y[:, None])` builds the GPy model on the
    first fit, `set_XY` updates it afterwards, `optimize_restarts(n_restarts)` tunes it on the host.  The fitted
    state GPy keeps for `predict_noiseless` -- `posterior.woodbury_chol` (Cholesky factor of K + sigma^2 I),
    `posterior.woodbury_vector` (K^-1 y), kernel variance / lengthscale, the target normalizer -- is then installed on
    the device, and `predict` / the fused acquisitions run there (noise-free predictive variance, as
    `predict_noiseless(full_cov=True)` returns).

    GPy is not installed in the image this was written in: the attribute names above follow GPy 1.x and the class is
    exercised in the tests with a duck-typed stand-in for the GPy model, not with GPy itself.
    """

    _KERNELS = {"rbf": "rbf", "Mat32": "matern32", "Mat52": "matern52", "Exponential": "matern12"}

    def __init__(self, gp_initializer, n_restarts: int = 1, dtype: str = "f64", device=None):
        Surrogate.__init__(self)
        if dtype not in ("f64", "f32"):
            raise ValueError("dtype must be 'f64' or 'f32'")
        self.gp_initializer = gp_initializer
        self.n_restarts = n_restarts
        self.gp = None
        self.dtype, self.device, self.device_fit, self.latency_max_m = dtype, device, False, None
        self.native, self.kernel_spec, self.fitted_on_device = None, None, False

    def _fit(self, x: np.ndarray, y: np.ndarray) -> None:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            if self.gp is None:
                self.gp = self.gp_initializer(x, y.reshape(-1, 1))
            else:
                self.gp.set_XY(x, y.reshape(-1, 1))
        self.gp.optimize_restarts(self.n_restarts)
        self.load_fitted_state()

    def load_fitted_state(self) -> None:
        from .kernel_spec import FlatKernel, UnsupportedKernelError
        gp = self.gp
        kern = gp.kern
        name = getattr(kern, "name", type(kern).__name__)
        if name not in self._KERNELS:
            raise UnsupportedKernelError(f"GPy kernel {name!r} is not supported (rbf, Mat32, Mat52, Exponential are)")
        spec = FlatKernel(kernel=self._KERNELS[name],
                          length_scale=np.atleast_1d(np.asarray(kern.lengthscale, dtype=np.float64)).ravel().copy(),
                          amplitude=float(np.ravel(kern.variance)[0]), noise_level=0.0)
        X = np.ascontiguousarray(np.asarray(gp.X), dtype=np.float64)
        n, d = X.shape
        L = np.ascontiguousarray(np.asarray(gp.posterior.woodbury_chol), dtype=np.float64)
        alpha = np.asarray(gp.posterior.woodbury_vector, dtype=np.float64).reshape(-1)
        normalizer = getattr(gp, "normalizer", None)
        y_mean = float(np.ravel(normalizer.mean)[0]) if normalizer is not None else 0.0
        y_std = float(np.ravel(normalizer.std)[0]) if normalizer is not None else 1.0
        self._native_for(n, d, spec.kernel).set_state(X, L, alpha, spec.length_scale, amplitude=spec.amplitude,
                                                      noise_level=0.0, y_mean=y_mean, y_std=y_std)
        self.kernel_spec = spec


class ScipyGPSurrogate(B200GPSurrogate):
    """The reference's class name (bopy/surrogate.py:72-91), same constructor argument: a scikit-learn
    `GaussianProcessRegressor`.  What differs from the reference, and is reported instead of silently approximated:

    * kernels: Constant x {RBF, Matern nu in 0.5 / 1.5 / 2.5 / inf} (+ White), isotropic or ARD.  Anything else
      (RationalQuadratic, DotProduct, ExpSineSquared, sums / products of two stationary kernels) raises
      `UnsupportedKernelError` at `fit` -- there is no CPU route behind this class (the reference's own
      `gp.predict(return_cov=True)` is one `pip install bopy` away for such kernels);
    * one target column, a CUDA device;
    * after a fit that ran on the device the scikit-learn object holds `X_train_`, `alpha_`, `kernel_` and the
      normalisation as its own fit would, and `L_` is fetched from the device the first time it is read (so
      `gp.predict(..., return_std / return_cov=True)` on the wrapped object keeps working)."""

    def _fit_on_device(self, x, y):
        super()._fit_on_device(x, y)
        _attach_lazy_factor(self)

    def _try_append(self, X, yn, y_mean, y_std, spec, kernel_):
        done = super()._try_append(X, yn, y_mean, y_std, spec, kernel_)
        if done:
            self.gp.__dict__.pop("L_", None)       # stale: re-fetched on next read
        return done


def _attach_lazy_factor(sur):
    """Give the wrapped scikit-learn object an `L_` that is downloaded from the device on first read."""
    gp = sur.gp
    cls = type(gp)
    if not getattr(cls, "_bopy_b200_lazy_factor", False):
        def L_(self):
            owner = self.__dict__.get("_bopy_b200_owner")
            if owner is None:
                raise AttributeError("L_")
            val = owner.export_factor()
            self.__dict__["L_"] = val
            return val
        # a non-data descriptor: an instance attribute `L_` (set by a host fit, or cached above) takes precedence
        class _Lazy:
            def __get__(self, obj, objtype=None):
                if obj is None:
                    return self
                return L_(obj)
        lazy_cls = type(cls.__name__, (cls,), {"L_": _Lazy(), "_bopy_b200_lazy_factor": True, "__module__": cls.__module__})
        gp.__class__ = lazy_cls
    gp.__dict__.pop("L_", None)
    gp.__dict__["_bopy_b200_owner"] = sur
