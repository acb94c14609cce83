"""The reference's call sequence on top of the REAL third-party libraries it delegates to.

TEST INFRASTRUCTURE (see oracle/gp_oracle.py header).  bopy's hot path is three thin wrappers
(``bopy/surrogate.py:87-91``, ``bopy/acquisition.py:83-85, 99-109, 123-131``) around
``sklearn.gaussian_process.GaussianProcessRegressor.predict(return_cov=True)`` and
``scipy.stats.norm``.  Both libraries are in this image (and on the GPU box), bopy itself is
not (``/root/reference`` does not travel, and its module-level imports need GPy / scipydirect /
dppy which are absent).  This module performs exactly bopy's calls, in bopy's order, against
those libraries; it is what ``bench.py --impl reference`` times and what the numpy restatement
in ``gp_oracle.py`` is validated against (next to the golden vectors made from the unmodified
reference by ``tools/make_golden.py``).
"""
from __future__ import annotations

import numpy as np
from scipy.stats import norm


def predict(gp, x):
    """bopy/surrogate.py:90-91."""
    return gp.predict(x, return_cov=True)


def lcb(gp, x, kappa=2.0):
    """bopy/acquisition.py:83-85."""
    mean, sigma = predict(gp, x)
    return mean - kappa * np.sqrt(np.diag(sigma))


def ei(gp, x, eta):
    """bopy/acquisition.py:99-106."""
    mean, sigma = predict(gp, x)
    var = np.diag(sigma)
    std = np.sqrt(var)
    return -var * norm.pdf(eta, loc=mean, scale=std) + (mean - eta) * norm.cdf(eta, loc=mean, scale=std)


def poi(gp, x, eta):
    """bopy/acquisition.py:123-128."""
    mean, sigma = predict(gp, x)
    std = np.sqrt(np.diag(sigma))
    return 1 - norm.cdf(eta, mean, std)


def acquisition_chunked(gp, kind, x, eta=0.0, kappa=2.0, chunk=64):
    """The acquisition over `x` the only way the reference can do it at scale: `chunk` candidates per
    call (bopy/optimizer.py:96-97 uses chunk=1; each call builds a chunk x chunk covariance)."""
    out = np.empty(x.shape[0])
    fn = {"lcb": lambda xx: lcb(gp, xx, kappa), "ei": lambda xx: ei(gp, xx, eta), "poi": lambda xx: poi(gp, xx, eta)}[kind]
    with np.errstate(invalid="ignore", divide="ignore"):
        for s in range(0, x.shape[0], chunk):
            out[s:s + chunk] = fn(x[s:s + chunk])
    return out
