"""Argument validation shared by surrogates and acquisition functions.

Mirrors the contract of the reference's ``FittableMixin`` (bopy/mixin.py:34-64): the same
checks in the same order with the same exception types and messages, because the reference's
tests assert the messages verbatim (tests/test_surrogate.py:55-101).  Validation stays on the
host, in Python, *before* anything is handed to the C-ABI ("Errors").
"""
import numpy as np

from .exceptions import NotFittedError

_FIT_RULES = (
    (lambda x, y: len(x) == 0, "`x` must contain at least one sample"),
    (lambda x, y: len(y) == 0, "`y` must contain at least one sample"),
    (lambda x, y: len(x) != len(y), "`x` and `y` must have the same number of samples"),
    (lambda x, y: np.ndim(x) != 2, "`x` must be 2D"),
    (lambda x, y: np.ndim(y) != 1, "`y` must be 1D"),
)


def check_fit_arguments(x: np.ndarray, y: np.ndarray) -> int:
    """Raise ValueError on the first violated rule; return the input dimensionality."""
    for violated, message in _FIT_RULES:
        if violated(x, y):
            raise ValueError(message)
    return x.shape[1]


def check_predict_arguments(x: np.ndarray, fitted: bool, n_dimensions: int) -> None:
    if not fitted:
        raise NotFittedError("must be fitted first")
    if len(x) == 0:
        raise ValueError("`x` must contain at least one sample")
    if np.ndim(x) != 2:
        raise ValueError("`x` must be 2D")
    if x.shape[1] != n_dimensions:
        raise ValueError("`x` must have the same number of dimensions as the training data")


class FittableMixin:
    """State (`has_been_fitted`, `n_dimensions`) plus the three validation hooks.

    A class using the mixin calls `_validate_ok_for_fitting(x, y)` at the top of its fit-like
    method, `_confirm_fit()` at the bottom, and `_validate_ok_for_predicting(x)` at the top of
    anything predict-like.
    """

    has_been_fitted = False
    n_dimensions = -1

    def _validate_ok_for_fitting(self, x: np.ndarray, y: np.ndarray) -> None:
        self.n_dimensions = check_fit_arguments(x, y)

    def _confirm_fit(self) -> None:
        self.has_been_fitted = True

    def _validate_ok_for_predicting(self, x: np.ndarray) -> None:
        check_predict_arguments(x, self.has_been_fitted, self.n_dimensions)
