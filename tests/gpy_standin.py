"""Minimal stand-in for the parts of GPy the reference touches (GPy is not installable offline):
`GPy.models.GPRegression(x, y, kernel=GPy.kern.RBF(input_dim=d), noise_var=..., normalizer=True)` with `set_XY`,
`optimize_restarts`, `predict_noiseless`, and the fitted attributes `GPyGPSurrogate` reads (bopy/surrogate.py:94-146):
`kern.name / variance / lengthscale`, `X`, `Y`, `normalizer.mean / std`, `posterior.woodbury_chol / woodbury_vector`.
Hyper-parameters stay at their initial values (`optimize_restarts` only counts its calls): the stand-in pins the
interface, not GPy's optimiser.  numpy / scipy only.  `install()` registers it as `sys.modules['GPy']`.
"""
import sys
import types

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular


class RBF:
    name = "rbf"

    def __init__(self, input_dim=1, variance=1.0, lengthscale=1.0, ARD=False):
        self.input_dim = input_dim
        self.variance = np.array([float(variance)])
        self.lengthscale = np.atleast_1d(np.asarray(lengthscale, dtype=np.float64)) * (np.ones(input_dim) if ARD else 1.0)

    def K(self, a, b=None):
        b = a if b is None else b
        d2 = (((a[:, None, :] - b[None, :, :]) / self.lengthscale) ** 2).sum(-1)
        return self.variance[0] * np.exp(-0.5 * d2)


class _Standardize:
    def __init__(self, y):
        self.mean, self.std = y.mean(axis=0), y.std(axis=0)


class _Posterior:
    pass


class GPRegression:
    def __init__(self, X, Y, kernel=None, noise_var=1.0, normalizer=None):
        self.kern = kernel if kernel is not None else RBF(X.shape[1])
        self.noise_var = float(noise_var)
        self._normalize = bool(normalizer)
        self.optimized = 0
        self.set_XY(X, Y)

    def set_XY(self, X, Y):
        self.X, self.Y = np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64)
        self.normalizer = _Standardize(self.Y) if self._normalize else None
        yn = (self.Y - self.normalizer.mean) / self.normalizer.std if self._normalize else self.Y
        K = self.kern.K(self.X) + self.noise_var * np.eye(len(self.X))
        L = cholesky(K, lower=True)
        self.posterior = _Posterior()
        self.posterior.woodbury_chol = L
        self.posterior.woodbury_vector = cho_solve((L, True), yn)

    def optimize_restarts(self, num_restarts=1, **_):
        self.optimized += num_restarts

    def predict_noiseless(self, Xnew, full_cov=False):
        Ks = self.kern.K(np.asarray(Xnew, dtype=np.float64), self.X)
        mu = Ks @ self.posterior.woodbury_vector
        V = solve_triangular(self.posterior.woodbury_chol, Ks.T, lower=True)
        cov = self.kern.K(np.asarray(Xnew, dtype=np.float64)) - V.T @ V
        if self._normalize:
            mu = mu * self.normalizer.std + self.normalizer.mean
            cov = cov * self.normalizer.std ** 2
        return mu, (cov if full_cov else np.diag(cov)[:, None])


def install():
    """Register the stand-in as `GPy` (+ `GPy.models`, `GPy.kern`) unless a real GPy is importable."""
    try:
        import GPy  # noqa: F401
        return sys.modules["GPy"]
    except ImportError:
        pass
    gpy = types.ModuleType("GPy")
    gpy.models = types.ModuleType("GPy.models")
    gpy.kern = types.ModuleType("GPy.kern")
    gpy.models.GPRegression = GPRegression
    gpy.kern.RBF = RBF
    gpy.__standin__ = True
    sys.modules.update({"GPy": gpy, "GPy.models": gpy.models, "GPy.kern": gpy.kern})
    return gpy
