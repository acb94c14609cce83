"""Objective functions used by the examples, tests and benchmark configs.

`forrester` and `bohachevsky` exist in the reference (bopy/benchmark_functions.py:4-42);
`branin` and `hartmann6` are the objectives BASELINE.json's configs name and the reference
lacks.  All take (n_samples, d) and return (n_samples,).  They only
generate training targets; none of them is on the hot path.
"""
import numpy as np


def forrester(x: np.ndarray) -> np.ndarray:
    """Forrester et al. 1-D test function on [0, 1]: (6x-2)^2 sin(12x-4)."""
    t = np.asarray(x, dtype=np.float64).reshape(-1)
    return np.square(6.0 * t - 2.0) * np.sin(12.0 * t - 4.0)


def bohachevsky(x: np.ndarray) -> np.ndarray:
    """Bohachevsky no. 1 in 2-D; global minimum 0 at the origin."""
    x = np.asarray(x, dtype=np.float64)
    a, b = x[:, 0], x[:, 1]
    return a * a + 2.0 * b * b - 0.3 * np.cos(3.0 * np.pi * a) - 0.4 * np.cos(4.0 * np.pi * b) + 0.7


def branin(x: np.ndarray) -> np.ndarray:
    """Branin-Hoo on [-5, 10] x [0, 15]; three global minima of value 0.397887."""
    x = np.asarray(x, dtype=np.float64)
    x1, x2 = x[:, 0], x[:, 1]
    b, c, r, s, t = 5.1 / (4.0 * np.pi ** 2), 5.0 / np.pi, 6.0, 10.0, 1.0 / (8.0 * np.pi)
    return (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1.0 - t) * np.cos(x1) + s


_H6_ALPHA = np.array([1.0, 1.2, 3.0, 3.2])
_H6_A = np.array([[10, 3, 17, 3.5, 1.7, 8],
                  [0.05, 10, 17, 0.1, 8, 14],
                  [3, 3.5, 1.7, 10, 17, 8],
                  [17, 8, 0.05, 10, 0.1, 14]], dtype=np.float64)
_H6_P = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886],
                         [2329, 4135, 8307, 3736, 1004, 9991],
                         [2348, 1451, 3522, 2883, 3047, 6650],
                         [4047, 8828, 8732, 5743, 1091, 381]], dtype=np.float64)


def hartmann6(x: np.ndarray) -> np.ndarray:
    """Hartmann 6-D on [0, 1]^6; global minimum -3.32237."""
    x = np.asarray(x, dtype=np.float64)
    diff = x[:, None, :] - _H6_P[None, :, :]
    inner = np.einsum("kj,nkj->nk", _H6_A, diff * diff)
    return -(np.exp(-inner) @ _H6_ALPHA)
