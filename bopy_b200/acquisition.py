"""Acquisition functions (to be MINIMISED, as in bopy/acquisition.py:14-17).

`LCB`, `EI`, `POI` keep the reference's constructors and `__call__` / `fit` template.  Their `_f`
never evaluates a formula on the host: with a B200 surrogate the posterior and the epilogue run in
one fused sweep; with any other `Surrogate` subclass the moments `predict` returned are pushed to the
device and the same epilogue code runs there (`bopy_acq_from_moments`).
"""
from abc import ABC, abstractmethod
from typing import Tuple

import numpy as np

from . import _native
from .mixin import FittableMixin
from .surrogate import Surrogate


class AcquisitionFunction(FittableMixin, ABC):
    """Expected-loss style acquisition: smaller is better."""

    def __init__(self, surrogate: Surrogate):
        super().__init__()
        self.surrogate = surrogate
        self.has_been_fitted = False
        self.n_dimensions = -1

    def __call__(self, x: np.ndarray) -> np.ndarray:
        """x: (n_samples, n_dimensions) -> (n_samples,)."""
        self._validate_ok_for_predicting(x)
        return self._f(x)

    def fit(self, x: np.ndarray, y: np.ndarray) -> None:
        """x: (n_samples, n_dimensions), y: (n_samples,)."""
        self._validate_ok_for_fitting(x, y)
        self._fit(x, y)
        self._confirm_fit()

    @abstractmethod
    def _f(self, x: np.ndarray) -> np.ndarray:
        ...

    def _fit(self, x: np.ndarray, y: np.ndarray) -> None:
        """Nothing to learn by default."""


class _MomentAcquisition(AcquisitionFunction):
    """An acquisition that is a pointwise function of the posterior (mean, variance)."""

    kind = ""          # 'lcb' | 'ei' | 'poi'  (bopy_acq enum names)

    def native_args(self) -> dict:
        """eta / kappa handed to the device epilogue."""
        return {}

    def _foreign_moments(self, x):
        """A non-native surrogate's predict() ran wherever it runs; push (mean, diag(cov)) to the device."""
        mean, sigma = self.surrogate.predict(x)
        torch = _native.require_cuda()
        dev = _native.resolve_device(None)
        mean_d = torch.as_tensor(np.ascontiguousarray(mean, dtype=np.float64), device=dev)
        var_d = torch.as_tensor(np.ascontiguousarray(np.diag(sigma), dtype=np.float64), device=dev)
        return mean_d, var_d

    def _f(self, x: np.ndarray) -> np.ndarray:
        sur = self.surrogate
        if hasattr(sur, "acquisition_values"):
            return sur.acquisition_values(self.kind, x, **self.native_args())
        mean_d, var_d = self._foreign_moments(x)
        out, _, _ = _native.acquisition_from_moments(self.kind, mean_d, var_d, **self.native_args())
        return out.cpu().numpy()

    def value_and_grad(self, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """(values (m,), gradient with respect to x (m, d)).  Additive to the reference's interface (which has no
        gradients); needs a B200-native surrogate."""
        self._validate_ok_for_predicting(x)
        sur = self.surrogate
        if not hasattr(sur, "acquisition_value_and_grad"):
            raise TypeError("value_and_grad needs a B200GPSurrogate")
        return sur.acquisition_value_and_grad(self.kind, x, **self.native_args())

    def argmin(self, x, index_base: int = 0, prune: bool = False, nan_policy: str = "first",
               on_device: bool = False):
        """Index and value of the smallest acquisition value over the rows of `x` (fused on the device).
        nan_policy='first': np.argmin's rules (first minimum, first NaN wins) -- what np.argmin over the reference's
        values gives; 'skip': np.nanargmin's (a NaN -- posterior variance rounded to <= 0 -- never wins; index -1 if
        all values are NaN): what the optimisers of this package use.  prune=True: branch and bound on a B200 surrogate
        (see B200GPSurrogate.acquisition_argmin).  on_device=True (B200 surrogates): (index, value) as (1,) device
        tensors, nothing synchronised."""
        self._validate_ok_for_predicting(x)
        sur = self.surrogate
        if hasattr(sur, "acquisition_argmin"):
            return sur.acquisition_argmin(self.kind, x, index_base=index_base, prune=prune, nan_policy=nan_policy,
                                          on_device=on_device, **self.native_args())
        mean_d, var_d = self._foreign_moments(x if isinstance(x, np.ndarray) else x.cpu().numpy())
        out, minv, mini = _native.acquisition_from_moments(self.kind, mean_d, var_d, want_min=True,
                                                           index_base=index_base, **self.native_args())
        if nan_policy == "skip":
            v = out.cpu().numpy()
            if np.isnan(v).all():
                return -1, 0.0
            i = int(np.nanargmin(v))
            return index_base + i, float(v[i])
        return int(mini.item()), float(minv.item())


class LCB(_MomentAcquisition):
    """Lower confidence bound: mean - kappa * std  (bopy/acquisition.py:67-85)."""

    kind = "lcb"

    def __init__(self, surrogate: Surrogate, kappa: float = 2.0):
        super().__init__(surrogate)
        self.kappa = kappa

    def native_args(self):
        return {"kappa": float(self.kappa)}


class _ImprovementAcquisition(_MomentAcquisition):
    """Acquisitions measured against the incumbent eta = min(y)  (bopy/acquisition.py:108-109, 130-131)."""

    def __init__(self, surrogate: Surrogate):
        super().__init__(surrogate)
        self._eta = np.inf

    def _fit(self, x: np.ndarray, y: np.ndarray) -> None:
        self._eta = np.min(y)

    def native_args(self):
        return {"eta": float(self._eta)}


class EI(_ImprovementAcquisition):
    """Negated expected improvement, -E[(eta - f(x))+]  (bopy/acquisition.py:88-106)."""

    kind = "ei"


class POI(_ImprovementAcquisition):
    """1 - P(f(x) <= eta)  (bopy/acquisition.py:112-128)."""

    kind = "poi"


class SequentialBatchAcquisitionFunction(AcquisitionFunction):
    """Wraps a base acquisition that is updated as batch members are picked one by one
    (bopy/acquisition.py:134-169)."""

    def __init__(self, base_acquisition: AcquisitionFunction):
        super().__init__(base_acquisition.surrogate)
        self.base_acquisition = base_acquisition

    def _fit(self, x: np.ndarray, y: np.ndarray) -> None:
        self.base_acquisition.fit(x, y)

    def start_batch(self) -> None:
        """Called before the first batch member is chosen."""

    def add_to_batch(self, optimization_result) -> None:
        """Called with each newly chosen batch member."""

    def finish_batch(self) -> None:
        """Called after the last batch member was chosen."""


class KriggingBeliever(SequentialBatchAcquisitionFunction):
    """Fantasise y = posterior mean at each chosen point and refit, restore the data at the end
    (bopy/acquisition.py:172-197)."""

    def _f(self, x: np.ndarray) -> np.ndarray:
        return self.base_acquisition(x)

    def argmin(self, x, index_base: int = 0, prune: bool = False, **kwargs):
        return self.base_acquisition.argmin(x, index_base=index_base, prune=prune, **kwargs)

    def start_batch(self) -> None:
        self.n_data = len(self.surrogate.x)

    def add_to_batch(self, optimization_result) -> None:
        # only the posterior mean is believed: take the diagonal-only entry where the surrogate has one (one-point
        # latency path) instead of predict()'s full covariance
        predict = getattr(self.surrogate, "predict_diag", self.surrogate.predict)
        believed, _ = predict(optimization_result.x_min)
        grown_x = np.concatenate((self.surrogate.x, optimization_result.x_min))
        grown_y = np.concatenate((self.surrogate.y, believed))
        self.surrogate.fit(grown_x, grown_y)

    def finish_batch(self) -> None:
        self.surrogate.fit(self.surrogate.x[: self.n_data], self.surrogate.y[: self.n_data])


class OneShotBatchAcquisitionFunction(AcquisitionFunction):
    """Records every (x, a(x)) the base acquisition is evaluated at (bopy/acquisition.py:200-242)."""

    def __init__(self, base_acquisition: AcquisitionFunction):
        super().__init__(base_acquisition.surrogate)
        self.base_acquisition = base_acquisition
        self.xs = []
        self.a_xs = []

    def _fit(self, x: np.ndarray, y: np.ndarray) -> None:
        self.base_acquisition.fit(x, y)

    def _f(self, x: np.ndarray) -> np.ndarray:
        values = self.base_acquisition(x)
        self.xs.append(np.array(x))
        self.a_xs.append(np.array(values))
        return values

    def argmin(self, x, index_base: int = 0, prune: bool = False, nan_policy: str = "first", on_device: bool = False):
        """The sweep optimisers' entry: one fused device pass evaluates ALL rows of `x` (numpy or device tensor),
        which are logged like any other evaluation; the arg-min follows `nan_policy` (see _MomentAcquisition.argmin).
        (`prune` is ignored: a one-shot strategy chooses from every evaluation, so none may be skipped.)"""
        self._validate_ok_for_predicting(x)
        base = self.base_acquisition
        sur = base.surrogate
        if hasattr(sur, "native") and hasattr(base, "kind"):
            xs = sur.native.candidates(x)
            sur.native.set_nan_policy(nan_policy)
            try:
                out = sur.native.sweep(xs, acq=base.kind, want_acq=True, want_min=True, index_base=index_base,
                                       **base.native_args())
            finally:
                sur.native.set_nan_policy("first")
            self.xs.append(xs)                 # stay on the device until a strategy asks for them
            self.a_xs.append(out["acq"])
            if on_device:
                return out["min_idx"], out["min_val"]
            return int(out["min_idx"].item()), float(out["min_val"].item())
        values = self._f(x if isinstance(x, np.ndarray) else x.cpu().numpy())
        if nan_policy == "skip" and np.isnan(values).all():
            return -1, 0.0
        i = int(np.nanargmin(values)) if nan_policy == "skip" else int(np.argmin(values))
        return index_base + i, float(values[i])

    def start_optimization(self) -> None:
        self.xs, self.a_xs = [], []

    def get_evaluations(self) -> Tuple[np.ndarray, np.ndarray]:
        to_host = lambda a: a if isinstance(a, np.ndarray) else a.cpu().numpy()   # noqa: E731
        return np.concatenate([to_host(a) for a in self.xs]), np.concatenate([to_host(a) for a in self.a_xs])

    def get_evaluations_on_device(self):
        """(x (N, d), a(x) (N,)) as device tensors: what a device-side strategy (top-k) selects from."""
        torch = _native.require_cuda()
        dev = _native.resolve_device(getattr(self.surrogate, "device", None))
        to_dev = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(a, dtype=torch.float64, device=dev)  # noqa: E731
        return torch.cat([to_dev(a) for a in self.xs]), torch.cat([to_dev(a) for a in self.a_xs])
