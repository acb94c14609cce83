"""Initial designs: `generate(bounds, n_points) -> (n_points, d)` (reference: bopy/initial_design.py:10-63).

Host-side, O(n*d), runs once per optimisation -- not on the hot path.  The reference pulls Sobol
points from `sobol_seq` and Latin hypercubes from `pyDOE` (both absent here); `scipy.stats.qmc`
provides the same constructions.
"""
from abc import ABC, abstractmethod

import numpy as np

from .bounds import Bounds


class InitialDesign(ABC):
    def generate(self, bounds: Bounds, n_points: int) -> np.ndarray:
        if n_points <= 0:
            raise ValueError("`n_points` must be positive.")
        unit = np.asarray(self._generate(bounds.n_dimensions, n_points), dtype=np.float64)
        lowers, uppers = np.asarray(bounds.lowers), np.asarray(bounds.uppers)
        return lowers + unit * (uppers - lowers)

    @abstractmethod
    def _generate(self, n_dimensions: int, n_points: int) -> np.ndarray:
        """n_points points in the unit cube [0, 1]^n_dimensions."""


class UniformRandomInitialDesign(InitialDesign):
    def _generate(self, n_dimensions, n_points):
        return np.random.rand(n_points, n_dimensions)


class SobolSequenceInitialDesign(InitialDesign):
    """Sobol points without the origin (bopy/initial_design.py:41-51 uses sobol_seq.i4_sobol_generate; this uses scipy's
    generator, whose Joe-Kuo direction numbers can differ from sobol_seq's tables in higher dimensions)."""

    def _generate(self, n_dimensions, n_points):
        from scipy.stats import qmc
        sampler = qmc.Sobol(d=n_dimensions, scramble=False)
        sampler.fast_forward(1)  # skip the origin, as sobol_seq.i4_sobol_generate does
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)  # n_points need not be a power of two
            return sampler.random(n_points)


class LatinHypercubeInitialDesign(InitialDesign):
    """Latin hypercube (bopy/initial_design.py:54-63 uses pyDOE.lhs, which draws from numpy's GLOBAL random state): the
    sampler is seeded from that state, so `np.random.seed(...)` makes a BO run reproducible here too."""

    def _generate(self, n_dimensions, n_points):
        import numpy as np
        from scipy.stats import qmc
        return qmc.LatinHypercube(d=n_dimensions, seed=int(np.random.randint(2 ** 31 - 1))).random(n_points)
