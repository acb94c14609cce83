"""Box bounds of the search space.

Same public surface as the reference (bopy/bounds.py:5-54: `Bound(lower, upper)`, `Bounds(bounds)` with
`n_dimensions`, `lowers`, `uppers`, and its two ValueError messages), written as plain value classes that also
hand the box to the device-side candidate generators as contiguous arrays (`as_arrays`, `clip`, `contains`).
"""
from typing import Iterable, Iterator, List, Sequence, Tuple

import numpy as np


class Bound:
    """One closed interval [lower, upper]; `lower` must be strictly below `upper`."""

    __slots__ = ("lower", "upper")

    def __init__(self, lower: float, upper: float):
        if not lower < upper:                      # also rejects NaN ends
            raise ValueError("`lower` must be less than `upper`")
        self.lower = lower
        self.upper = upper

    @property
    def width(self) -> float:
        return self.upper - self.lower

    def __iter__(self) -> Iterator[float]:         # lets `lo, hi = bound` and zip(*bounds) work
        yield self.lower
        yield self.upper

    def __eq__(self, other) -> bool:
        return isinstance(other, Bound) and (self.lower, self.upper) == (other.lower, other.upper)

    def __hash__(self) -> int:
        return hash((self.lower, self.upper))

    def __repr__(self) -> str:
        return f"Bound(lower={self.lower!r}, upper={self.upper!r})"


class Bounds:
    """The search box: a non-empty sequence of `Bound`s, one per input dimension, in order."""

    def __init__(self, bounds: Iterable[Bound]):
        self.bounds: List[Bound] = list(bounds)
        if len(self.bounds) == 0:
            raise ValueError("`bounds` must contain at least one bound.")

    # -- the reference's three accessors ---------------------------------------------------------------------------
    @property
    def n_dimensions(self) -> int:
        return len(self.bounds)

    @property
    def lowers(self) -> List[float]:
        return [lo for lo, _ in self.bounds]

    @property
    def uppers(self) -> List[float]:
        return [hi for _, hi in self.bounds]

    # -- what the sweep optimisers use -----------------------------------------------------------------------------
    def as_arrays(self) -> Tuple[np.ndarray, np.ndarray]:
        """(lowers, uppers) as float64 arrays of length n_dimensions."""
        return np.asarray(self.lowers, dtype=np.float64), np.asarray(self.uppers, dtype=np.float64)

    def clip(self, x: np.ndarray) -> np.ndarray:
        """x (…, n_dimensions) clipped into the box."""
        lo, hi = self.as_arrays()
        return np.minimum(np.maximum(x, lo), hi)

    def contains(self, x: np.ndarray) -> np.ndarray:
        """Row-wise membership of x (m, n_dimensions) in the closed box."""
        lo, hi = self.as_arrays()
        x = np.atleast_2d(x)
        return np.logical_and(x >= lo, x <= hi).all(axis=1)

    def __len__(self) -> int:
        return len(self.bounds)

    def __getitem__(self, i: int) -> Bound:
        return self.bounds[i]

    def __eq__(self, other) -> bool:
        return isinstance(other, Bounds) and self.bounds == other.bounds

    def __repr__(self) -> str:
        return f"Bounds(bounds={self.bounds!r})"


def unit_box(n_dimensions: int) -> Bounds:
    """[0, 1]^n."""
    return Bounds([Bound(0.0, 1.0) for _ in range(n_dimensions)])
