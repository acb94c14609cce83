"""Box bounds of the search space (reference: bopy/bounds.py:5-54)."""
from dataclasses import dataclass
from typing import List


@dataclass
class Bound:
    """One closed interval [lower, upper] with lower < upper."""

    lower: float
    upper: float

    def __post_init__(self):
        if not self.lower < self.upper:
            raise ValueError("`lower` must be less than `upper`")


@dataclass
class Bounds:
    """An ordered, non-empty list of `Bound`s, one per input dimension."""

    bounds: List[Bound]

    def __post_init__(self):
        if not self.bounds:
            raise ValueError("`bounds` must contain at least one bound.")

    @property
    def n_dimensions(self) -> int:
        return len(self.bounds)

    @property
    def lowers(self) -> List[float]:
        return [bound.lower for bound in self.bounds]

    @property
    def uppers(self) -> List[float]:
        return [bound.upper for bound in self.bounds]
