"""The oracle's acquisition gradient (oracle/gp_oracle.py: acquisition_value_and_grad) against central finite
differences of the oracle's own -- golden-pinned -- acquisition values.  The reference has no gradient code."""
import numpy as np
import pytest

from conftest import golden_state
from oracle import gp_oracle as O

CASES = ["iso_rbf_bare", "ard_amp_white", "matern15_d2", "matern25_d2", "c1_forrester_rbf_n6", "ragged_n333_d4_opt"]


@pytest.mark.parametrize("acq", ["lcb", "ei", "poi"])
@pytest.mark.parametrize("name", CASES)
def test_gradient_matches_finite_differences(name, acq):
    g, st = golden_state(name)
    rng = np.random.default_rng(3)
    lo, hi = st.X_train.min(0), st.X_train.max(0)
    xs = lo + rng.random((24, st.X_train.shape[1])) * (hi - lo)
    eta = float(g["eta"])
    a, grad, mean, var = O.acquisition_value_and_grad(st, acq, xs, eta=eta, kappa=2.0)
    _, _, a_ref, _ = O.acquisition_sweep(st, acq, xs, eta=eta, kappa=2.0)
    np.testing.assert_allclose(a, a_ref, rtol=1e-12, atol=1e-14)
    span = hi - lo
    checked = 0
    for q in range(xs.shape[1]):
        h = 1e-5 * span[q]
        e = np.zeros(xs.shape[1])
        e[q] = h
        ap = O.acquisition_sweep(st, acq, xs + e, eta=eta, kappa=2.0)[2]
        am = O.acquisition_sweep(st, acq, xs - e, eta=eta, kappa=2.0)[2]
        fd = (ap - am) / (2 * h)
        # where the posterior variance is resolved the difference quotient is trustworthy
        ok = (var > 1e-6 * O.kernel_self_diag(st.kernel, 1)[0] * st.y_std ** 2) & np.isfinite(fd) & np.isfinite(grad[:, q])
        scale = np.abs(grad[ok]).max() if ok.any() else 1.0
        np.testing.assert_allclose(grad[ok, q], fd[ok], rtol=2e-4, atol=2e-5 * scale + 1e-12)
        checked += int(ok.sum())
    assert checked > 0


def test_matern12_gradient_away_from_the_kink():
    g, st = golden_state("matern05_d2")
    rng = np.random.default_rng(4)
    xs = st.X_train.min(0) + rng.random((16, 2)) * np.ptp(st.X_train, axis=0)
    a, grad, _, var = O.acquisition_value_and_grad(st, "lcb", xs, kappa=2.0)
    for q in range(2):
        e = np.zeros(2)
        e[q] = 1e-6
        fd = (O.acquisition_sweep(st, "lcb", xs + e, kappa=2.0)[2] - O.acquisition_sweep(st, "lcb", xs - e, kappa=2.0)[2]) / 2e-6
        np.testing.assert_allclose(grad[:, q], fd, rtol=1e-3, atol=1e-4 * np.abs(grad).max())


def test_partials_nan_rule():
    dm, dv = O.acquisition_partials("ei", np.array([0.0, 0.0]), np.array([0.0, -1.0]), eta=0.1)
    assert np.isnan(dm).all() and np.isnan(dv).all()


def test_multistart_step_rule_minimises_a_box_constrained_quadratic():
    """The lock-step projected-gradient rule (restated from multistart_step_kernel) on f(x) = 1/2 (x-c)' A (x-c):
    monotone, stays in the box, reaches the projected optimum; NaN trial values are never accepted."""
    rng = np.random.default_rng(0)
    S, d = 64, 4
    lo, hi = np.zeros(d), np.ones(d)
    A = np.diag([1.0, 10.0, 100.0, 3.0])
    c = np.array([0.3, 0.7, 1.4, -0.2])          # two coordinates of the optimum are pinned to the box

    def fg(x):
        r = x - c
        return 0.5 * np.einsum("sd,de,se->s", r, A, r), r @ A

    xt = rng.random((S, d))
    xc, gc, fc, alpha = np.empty_like(xt), np.empty_like(xt), np.empty(S), np.ones(S)
    history = []
    for k in range(80):
        ft, gt = fg(xt)
        if k == 5:
            ft = ft.copy()
            ft[::2] = np.nan                      # a NaN evaluation must be rejected, not adopted
        O.multistart_step(lo, hi, xc, fc, gc, xt, ft, gt, alpha, first=(k == 0))
        assert (xt >= lo).all() and (xt <= hi).all() and not np.isnan(fc).any()
        history.append(fc.copy())
    history = np.array(history)
    assert (np.diff(history, axis=0) <= 0).all()                       # monotone per start
    best = np.clip(c, lo, hi)
    assert np.abs(xc - best).max() < 1e-6


def test_multistart_optimizer_argument_validation():
    from bopy_b200.bounds import Bound, Bounds
    from bopy_b200.optimizer import MultiStartOptimizer
    b = Bounds([Bound(0.0, 1.0)])
    with pytest.raises(ValueError, match="method"):
        MultiStartOptimizer(None, b, method="newton")
    with pytest.raises(ValueError, match="non-negative"):
        MultiStartOptimizer(None, b, iterations=-1)
    with pytest.raises(ValueError, match="multiple of 128"):
        MultiStartOptimizer(None, b, points_per_start=100)
