"""Acquisition gradient on the device (grad_kernel: mirrored backward solve + kernel-derivative sums) against the
oracle's analytic gradient (itself pinned by finite differences, tests/test_oracle_gradient.py).  -m gpu."""
import numpy as np
import pytest

from conftest import golden_names
from oracle import gp_oracle as O
from parity_util import prior_var
from test_gpu_parity import cached_native

pytestmark = pytest.mark.gpu


def grad_bound(st, xs, acq, eta, kappa, mean, var):
    """Rounding-level bound: |d acq/d mean| y_std sum|alpha_i dk_i| + |d acq/d var| 2 y_var sum|w_i dk_i|, times eps-ish,
    plus the sensitivity of the partials to the moments' own tolerance."""
    from scipy.linalg import solve_triangular
    spec = st.kernel
    ls = np.broadcast_to(spec.length_scale, (st.X_train.shape[1],))
    Kt = O.kernel_cross(spec, xs, st.X_train)
    V = solve_triangular(st.L, Kt.T, lower=True, check_finite=False)
    W = solve_triangular(st.L.T, V, lower=False, check_finite=False)
    kd = np.abs(O.kernel_cross_grad_factor(spec, xs, st.X_train))
    diff = np.abs(xs[:, None, :] - st.X_train[None, :, :]) / (ls * ls)
    am = np.einsum("mn,mnd->md", kd * np.abs(st.alpha)[None, :], diff) * st.y_std
    av = np.einsum("mn,mnd->md", kd * np.abs(W.T), diff) * 2 * st.y_std ** 2
    dm, dv = O.acquisition_partials(acq, mean, var, eta, kappa)
    cond = np.linalg.cond(st.L)
    return (np.abs(dm)[:, None] * am + np.abs(dv)[:, None] * av) * 64 * np.finfo(np.float64).eps * max(cond, 1.0)


@pytest.mark.parametrize("acq", ["lcb", "ei", "poi"])
@pytest.mark.parametrize("name", [n for n in golden_names() if not n.startswith(("edge_alpha0", "c5"))])
def test_gradient_matches_the_oracle(name, acq):
    g, st, gp = cached_native(name, "f64")
    xs, eta = g["Xs"][:300], float(g["eta"])
    val, grad, mean, var = (t.cpu().numpy() for t in gp.value_and_grad(gp.candidates(xs), acq, eta=eta, kappa=2.0))
    o_val, o_grad, o_mean, o_var = O.acquisition_value_and_grad(st, acq, xs, eta=eta, kappa=2.0)
    resolved = (o_var > 1e-6 * prior_var(st)) & np.isfinite(o_grad).all(axis=1)
    assert resolved.any()
    bound = grad_bound(st, xs, acq, eta, 2.0, o_mean, o_var)
    scale = np.abs(o_grad[resolved]).max()
    err = np.abs(grad - o_grad)[resolved]
    assert (err <= bound[resolved] + 1e-7 * np.abs(o_grad[resolved]) + 1e-9 * scale).all(), \
        f"worst {np.max(err / (bound[resolved] + 1e-7 * np.abs(o_grad[resolved]) + 1e-9 * scale)):.3g}x"
    np.testing.assert_allclose(val[resolved], o_val[resolved], rtol=1e-7, atol=1e-7 * (np.ptp(o_val[resolved]) or 1.0))
    assert np.array_equal(np.isnan(grad).any(axis=1), ~(np.sqrt(np.where(var < 0, np.nan, var)) > 0))


@pytest.mark.parametrize("m", [1, 8, 9, 33, 150, 1185, 5000])
def test_every_batch_shape_and_chunking(m):
    """8/16/32-candidate batches, several batches per group, and (m = 5000) two chunks of the V/W workspace."""
    g, st, gp = cached_native("c4_hartmann6_n2048", "f64")
    rng = np.random.default_rng(m)
    xs = rng.random((m, 6))
    eta = float(g["eta"])
    val, grad, mean, var = (t.cpu().numpy() for t in gp.value_and_grad(gp.candidates(xs), "ei", eta=eta))
    sub = slice(0, min(m, 200))
    o_val, o_grad, o_mean, o_var = O.acquisition_value_and_grad(st, "ei", xs[sub], eta=eta)
    bound = grad_bound(st, xs[sub], "ei", eta, 2.0, o_mean, o_var)
    scale = np.abs(o_grad).max()
    assert (np.abs(grad[sub] - o_grad) <= bound + 1e-7 * np.abs(o_grad) + 1e-9 * scale).all()
    # a candidate's gradient does not depend on batch shape / position
    if m >= 33:
        one = gp.value_and_grad(gp.candidates(xs[17:18]), "ei", eta=eta)[1].cpu().numpy()
        assert np.array_equal(one[0], grad[17])
    if m == 5000:
        tail = gp.value_and_grad(gp.candidates(xs[4900:]), "ei", eta=eta)[1].cpu().numpy()
        assert np.array_equal(tail, grad[4900:])


def test_large_n():
    g, st, gp = cached_native("c5_rbf_d20_n8192", "f64")
    xs, eta = g["Xs"][:40], float(g["eta"])
    val, grad, mean, var = (t.cpu().numpy() for t in gp.value_and_grad(gp.candidates(xs), "lcb", kappa=2.0))
    o_val, o_grad, o_mean, o_var = O.acquisition_value_and_grad(st, "lcb", xs, kappa=2.0)
    bound = grad_bound(st, xs, "lcb", eta, 2.0, o_mean, o_var)
    assert (np.abs(grad - o_grad) <= bound + 1e-7 * np.abs(o_grad) + 1e-9 * np.abs(o_grad).max()).all()


def test_public_api_and_descent_direction():
    """A small step against the gradient lowers the acquisition (what the refinement relies on)."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel

    from bopy_b200.acquisition import LCB
    from bopy_b200.surrogate import B200GPSurrogate
    rng = np.random.default_rng(2)
    X = rng.random((300, 4))
    y = np.sin(3 * X[:, 0]) * X[:, 1] + X[:, 2] ** 2 - X[:, 3]
    sur = B200GPSurrogate(GaussianProcessRegressor(ConstantKernel(1.0) * RBF(0.4 * np.ones(4)), alpha=1e-6,
                                                    normalize_y=True, optimizer=None))
    sur.fit(X, y)
    acq = LCB(sur)
    acq.fit(X, y)
    xs = 0.1 + 0.8 * rng.random((64, 4))
    val, grad = acq.value_and_grad(xs)
    assert val.shape == (64,) and grad.shape == (64, 4)
    np.testing.assert_allclose(val, acq(xs), rtol=1e-12, atol=1e-12)
    step = 1e-4 / np.maximum(np.linalg.norm(grad, axis=1, keepdims=True), 1e-12)
    lower = acq(xs - step * grad)
    assert (lower < val).mean() > 0.95


def test_fp32_handle_is_refused():
    g, st, gp = cached_native("c3_branin_n256", "f32")
    with pytest.raises(Exception, match="fp64 handle"):
        gp.value_and_grad(gp.candidates(g["Xs"][:4]), "ei")


def test_step_kernel_is_the_oracle_rule_bit_for_bit():
    import torch

    from bopy_b200 import _native
    rng = np.random.default_rng(9)
    S, d = 257, 5
    lo, hi = -1.0 + np.zeros(d), np.array([1.0, 2.0, 0.5, 3.0, 1.5])
    for first in (True, False):
        xc = lo + rng.random((S, d)) * (hi - lo)
        xt = np.clip(xc + 0.05 * rng.standard_normal((S, d)), lo, hi)
        fc, ft = rng.standard_normal(S), rng.standard_normal(S)
        gc, gt = rng.standard_normal((S, d)), rng.standard_normal((S, d))
        alpha = np.abs(rng.standard_normal(S)) + 0.01
        ft[::17] = np.nan            # NaN trial values are never accepted
        fc[5::31] = np.nan           # ... unless the current value is NaN too
        gt[3::29] = np.nan           # NaN gradients freeze the start
        gt[7] = 0.0
        dev = [torch.as_tensor(a.copy(), device="cuda") for a in (xc, fc, gc, xt, ft, gt, alpha)]
        _native.multistart_step(lo, hi, *dev, first=first)
        host = [a.copy() for a in (xc, fc, gc, xt, ft, gt, alpha)]
        O.multistart_step(lo, hi, *host, first=first)
        for name, a, b in zip(("xc", "fc", "gc", "xt", "ft", "gt", "alpha"), dev, host):
            assert np.array_equal(a.cpu().numpy(), b, equal_nan=True), (name, first)


def test_gradient_multistart_finds_stationary_points():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel

    from bopy_b200.acquisition import LCB
    from bopy_b200.benchmark_functions import hartmann6
    from bopy_b200.bounds import Bound, Bounds
    from bopy_b200.optimizer import MultiStartOptimizer
    from bopy_b200.surrogate import B200GPSurrogate
    rng = np.random.default_rng(8)
    X = rng.random((500, 6))
    y = hartmann6(X)
    sur = B200GPSurrogate(GaussianProcessRegressor(ConstantKernel(1.0) * RBF(0.3 * np.ones(6)), alpha=1e-6,
                                                    normalize_y=True, optimizer=None))
    sur.fit(X, y)
    acq = LCB(sur, kappa=2.0)
    acq.fit(X, y)
    bounds = Bounds([Bound(0.0, 1.0)] * 6)
    grad_opt = MultiStartOptimizer(acq, bounds, n_starts=128, n_candidates=1 << 16, seed=3, method="gradient", iterations=60)
    cloud_opt = MultiStartOptimizer(acq, bounds, n_starts=128, n_candidates=1 << 16, seed=3, method="cloud", rounds=6)
    rg, rc = grad_opt.optimize(), cloud_opt.optimize()
    assert rg.x_min.shape == (1, 6) and (rg.x_min >= 0).all() and (rg.x_min <= 1).all()
    assert rg.f_min[0] <= rc.f_min[0] + 1e-3 * abs(rc.f_min[0])          # as good as the derivative-free clouds
    assert abs(acq(rg.x_min)[0] - rg.f_min[0]) <= 1e-10 * max(1.0, abs(rg.f_min[0]))
    # the refined starts are (projected-)stationary: the gradient vanishes in every coordinate that is not pinned to a bound
    xs, vs = grad_opt.local_minima()
    _, g = acq.value_and_grad(xs)
    free = (xs > 1e-9) & (xs < 1 - 1e-9)
    pg = np.where(free, g, np.where(xs <= 1e-9, np.minimum(g, 0.0), np.maximum(g, 0.0)))
    start_vals = None
    assert np.median(np.abs(pg).max(axis=1)) < 1e-3 * np.abs(g).max() + 1e-6
    # never worse than where it started: monotone method
    assert vs.min() == rg.f_min[0]
