"""Fixed-hyper-parameter fit on the device (bopy_gp_fit: Gram + blocked Cholesky + alpha_) against scikit-learn's
fit ($SK/_gpr.py:349-367) and the golden vectors of the unmodified reference.  -m gpu."""
import numpy as np
import pytest
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

from conftest import golden_names, golden_state
from parity_util import check_mean, check_var

from bopy_b200.acquisition import EI
from bopy_b200.surrogate import B200GPSurrogate

pytestmark = pytest.mark.gpu


def sklearn_kernel(st):
    k = st.kernel
    base = RBF(k.length_scale) if k.kind == "rbf" else Matern(k.length_scale, nu=k.nu)
    kern = ConstantKernel(k.amplitude) * base
    return kern + WhiteKernel(k.noise_level) if k.noise_level > 0 else kern


def device_fitted(name, dtype="f64"):
    g, st = golden_state(name)
    gp = GaussianProcessRegressor(kernel=sklearn_kernel(st), alpha=float(g["alpha_reg"]),
                                  normalize_y=bool(g["normalize_y"]), optimizer=None)
    sur = B200GPSurrogate(gp, dtype=dtype, device_fit=True)
    sur.fit(g["X"], g["y"])
    assert sur.fitted_on_device
    return g, st, sur


@pytest.mark.parametrize("name", [n for n in golden_names() if n != "edge_alpha0_nan"])
def test_device_fit_reproduces_the_reference_posterior(name):
    g, st, sur = device_fitted(name)
    mean, var = sur.predict_diag(g["Xs"])
    err, bound = check_mean(mean, g["mean"], st, "f64")
    # two backward-stable Cholesky factorisations of K agree in the weights only up to eps * cond(K): the fit
    # parity bound carries that term (it exceeds the 1e-9 convention only for cond(K) > 1e9, i.e. alpha_reg <= 1e-10
    # with candidates on training points)
    cond_term = 0.0
    if st.L.shape[0] <= 2048:
        s = np.linalg.svd(st.L, compute_uv=False)
        cond_term = 1e-2 * np.finfo(np.float64).eps * (s[0] / s[-1]) ** 2
    assert (err <= bound + cond_term * st.y_std).all(), f"mean: worst {np.max(err / bound):.3g}x the bound"
    err, bound = check_var(var, g["var"], st, "f64")
    from parity_util import prior_var
    assert (err <= bound + cond_term * prior_var(st)).all(), f"var: worst {np.max(err / bound):.3g}x the bound"
    ei = EI(sur)
    ei.fit(g["X"], g["y"])
    idx, _ = ei.argmin(g["Xs"])
    ref = g["ei"]
    assert idx == int(g["argmin_ei"]) or abs(ref[idx] - ref.min()) <= 1e-9 * max(np.ptp(ref), abs(ref.min()))
    # the scikit-learn object is left as its own fit would leave it (minus the factor)
    assert np.array_equal(sur.gp.X_train_, g["X"]) and sur.gp.alpha_.shape == (len(g["X"]),)
    assert float(np.ravel(sur.gp._y_train_std)[0]) == pytest.approx(st.y_std, rel=1e-15)


@pytest.mark.parametrize("name", ["c3_branin_n256", "ragged_n333_d4_opt", "c4_hartmann6_n2048", "ard_amp_white"])
def test_device_cholesky_factor_and_weights(name):
    g, st, sur = device_fitted(name)
    L = sur.export_factor()
    n = len(g["X"])
    assert L.shape == (n, n) and np.array_equal(L, np.tril(L))
    # same factor as LAPACK's up to rounding: compare through the matrix it factors (backward error) and directly
    K = st.L @ st.L.T
    assert np.max(np.abs(L @ L.T - K)) <= 1e-13 * np.max(np.abs(K)) * n ** 0.5
    assert np.max(np.abs(L - st.L)) <= 1e-9 * np.max(np.abs(st.L))
    # alpha_ is ill-conditioned (cond(K) ~ 1e8): compare its action on the kernel rows, which is what predict uses
    from oracle import gp_oracle as O
    Kt = O.kernel_cross(st.kernel, g["Xs"][:64], st.X_train)
    np.testing.assert_allclose(Kt @ sur.gp.alpha_, Kt @ st.alpha, rtol=1e-9, atol=1e-9)


def test_device_fit_is_used_by_default_only_for_fixed_hyper_parameters():
    x = np.linspace(0, 1, 12).reshape(-1, 1)
    y = np.sin(6 * x).ravel()
    fixed = B200GPSurrogate(GaussianProcessRegressor(kernel=RBF(0.3), alpha=1e-8, optimizer=None))
    fixed.fit(x, y)
    assert fixed.fitted_on_device
    tuned = B200GPSurrogate(GaussianProcessRegressor(kernel=RBF(0.3), alpha=1e-8))
    tuned.fit(x, y)
    assert not tuned.fitted_on_device
    with pytest.raises(ValueError, match="device_fit=True needs a scalar alpha"):
        B200GPSurrogate(GaussianProcessRegressor(kernel=RBF(0.3), alpha=np.full(12, 1e-6)), device_fit=True).fit(x, y)
    # same predictions from both routes when the hyper-parameters coincide
    host = B200GPSurrogate(GaussianProcessRegressor(kernel=RBF(0.3), alpha=1e-8, optimizer=None), device_fit=False)
    host.fit(x, y)
    grid = np.linspace(0, 1, 101).reshape(-1, 1)
    m1, v1 = fixed.predict_diag(grid)
    m2, v2 = host.predict_diag(grid)
    np.testing.assert_allclose(m1, m2, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(v1, v2, rtol=1e-6, atol=1e-11)


def test_not_positive_definite_raises_like_scipy():
    x = np.array([[0.1], [0.1], [0.7]])           # duplicate row, alpha = 0 -> singular Gram matrix
    y = np.array([1.0, 1.0, 2.0])
    sur = B200GPSurrogate(GaussianProcessRegressor(kernel=RBF(0.3), alpha=0.0, optimizer=None))
    with pytest.raises(np.linalg.LinAlgError, match="not positive definite"):
        sur.fit(x, y)


def test_refit_chain_like_kriging_believer():
    """Fantasise-and-refit (bopy/acquisition.py:188-197) with every refit on the device."""
    from bopy_b200.acquisition import LCB, KriggingBeliever
    from bopy_b200.bounds import Bound, Bounds
    from bopy_b200.optimizer import CandidateSweepOptimizer, SequentialBatchOptimizer
    g, st, sur = device_fitted("c3_branin_n256")
    bounds = Bounds([Bound(-5.0, 10.0), Bound(0.0, 15.0)])
    kb = KriggingBeliever(LCB(sur))
    kb.fit(g["X"], g["y"])
    opt = SequentialBatchOptimizer(kb, bounds, CandidateSweepOptimizer(kb, bounds, n_candidates=1 << 14, seed=5), batch_size=3)
    res = opt.optimize()
    assert res.x_min.shape == (3, 2) and len(sur.x) == 256 and sur.fitted_on_device
    assert len({tuple(np.round(p, 6)) for p in res.x_min}) == 3


LML_KERNELS = [
    ConstantKernel(1.7) * RBF(0.4),
    ConstantKernel(0.6) * RBF([0.3, 0.5, 0.9]),
    RBF([0.3, 0.5, 0.9]),
    ConstantKernel(2.0) * Matern(0.7, nu=0.5),
    ConstantKernel(2.0) * Matern([0.4, 0.8, 0.6], nu=1.5),
    Matern(0.5, nu=2.5) * ConstantKernel(1.3),
    ConstantKernel(1.1) * RBF([0.3, 0.5, 0.9]) + WhiteKernel(1e-2),
    ConstantKernel(1.5, constant_value_bounds="fixed") * RBF(0.4),
]


@pytest.mark.parametrize("kernel", LML_KERNELS, ids=[str(i) for i in range(len(LML_KERNELS))])
@pytest.mark.parametrize("n", [50, 333])
def test_device_lml_and_gradient_match_sklearn(kernel, n):
    """bopy_gp_lml against GaussianProcessRegressor.log_marginal_likelihood ($SK/_gpr.py:541-656)."""
    from bopy_b200 import _native
    from bopy_b200.kernel_spec import flatten_sklearn_kernel, theta_gradient
    rng = np.random.default_rng(n)
    X = rng.random((n, 3))
    y = np.sin(4 * X[:, 0]) + X[:, 1] ** 2 - X[:, 2]
    ref = GaussianProcessRegressor(kernel=kernel, alpha=1e-6, normalize_y=True, optimizer=None).fit(X, y)
    theta = ref.kernel_.theta
    lml_ref, grad_ref = ref.log_marginal_likelihood(theta, eval_gradient=True)
    flat = flatten_sklearn_kernel(ref.kernel_)
    gp = _native.NativeGP(n, 3, kernel=flat.kernel)
    lml, g = gp.lml(X, ref.y_train_, flat.length_scale, amplitude=flat.amplitude, noise_level=flat.noise_level,
                    alpha_reg=1e-6, want_grad=True)
    assert lml == pytest.approx(lml_ref, rel=1e-10, abs=1e-8)
    grad = theta_gradient(ref.kernel_, g)
    assert grad.shape == grad_ref.shape
    np.testing.assert_allclose(grad, grad_ref, rtol=1e-7, atol=1e-7 * max(1.0, np.max(np.abs(grad_ref))))
    value_only, none = gp.lml(X, ref.y_train_, flat.length_scale, amplitude=flat.amplitude,
                              noise_level=flat.noise_level, alpha_reg=1e-6, want_grad=False)
    assert none is None and value_only == lml


def test_hyper_parameter_optimisation_on_the_device():
    rng = np.random.default_rng(5)
    X = rng.random((300, 2))
    y = np.sin(5 * X[:, 0]) * np.cos(3 * X[:, 1]) + 0.05 * rng.standard_normal(300)
    kernel = ConstantKernel(1.0) * RBF([1.0, 1.0]) + WhiteKernel(1e-1)
    host = GaussianProcessRegressor(kernel=kernel, alpha=1e-8, normalize_y=True, random_state=0).fit(X, y)
    sur = B200GPSurrogate(GaussianProcessRegressor(kernel=kernel, alpha=1e-8, normalize_y=True, random_state=0),
                          device_fit=True)
    sur.fit(X, y)
    assert sur.fitted_on_device
    # same optimum as scikit-learn's host route (smooth, well-identified problem)
    np.testing.assert_allclose(sur.gp.kernel_.theta, host.kernel_.theta, rtol=1e-3, atol=1e-3)
    assert sur.gp.log_marginal_likelihood_value_ == pytest.approx(host.log_marginal_likelihood_value_, rel=1e-6)
    grid = rng.random((200, 2))
    mean, var = sur.predict_diag(grid)
    h_mean, h_std = host.predict(grid, return_std=True)
    np.testing.assert_allclose(mean, h_mean, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(np.sqrt(var), h_std, rtol=1e-3, atol=1e-5)


def test_one_point_growth_is_a_row_append_and_matches_a_fresh_fit():
    """A refit on (previous data + one point) with fixed hyper-parameters extends the factor by one row
    (bopy_gp_append); it must give what a fit from scratch gives, also through a chain of appends, across a
    128-row block edge (where it falls back to the full fit) and after the Kriging believer's shrink."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import Matern, RBF, ConstantKernel, WhiteKernel

    from bopy_b200.surrogate import B200GPSurrogate
    rng = np.random.default_rng(77)
    X = rng.random((400, 3))
    y = np.sin(4 * X[:, 0]) + X[:, 1] * X[:, 2] + 0.05 * rng.standard_normal(400)
    xs = rng.random((200, 3))
    kernels = [ConstantKernel(1.3) * RBF([0.3, 0.4, 0.5]), ConstantKernel(0.8) * Matern(0.5, nu=2.5) + WhiteKernel(1e-3)]
    for kernel in kernels:
        def make():
            return B200GPSurrogate(GaussianProcessRegressor(kernel=kernel, alpha=1e-6, normalize_y=True, optimizer=None))
        sur = make()
        n0 = 370
        sur.fit(X[:n0], y[:n0])
        assert getattr(sur, "appended_rows", 0) == 0
        for n in range(n0 + 1, n0 + 20):                         # 371 .. 389: crosses 384 = 3 * 128
            sur.fit(X[:n], y[:n])
            if n in (n0 + 1, 384, 385, n0 + 19):
                fresh = make()
                fresh.fit(X[:n], y[:n])
                m_a, v_a = sur.predict_diag(xs)
                m_f, v_f = fresh.predict_diag(xs)
                scale = np.abs(m_f).max()
                np.testing.assert_allclose(m_a, m_f, rtol=0, atol=1e-9 * scale)
                np.testing.assert_allclose(v_a, v_f, rtol=1e-8, atol=1e-10 * np.abs(v_f).max())
                np.testing.assert_allclose(sur.gp.alpha_, fresh.gp.alpha_, rtol=0, atol=1e-7 * np.abs(fresh.gp.alpha_).max())
        # 19 growth steps, all but the one that needed a fourth block (384 -> 385) were appends
        assert sur.appended_rows == 18
        # shrinking back to a leading subset (what KriggingBeliever.finish_batch does): 389 -> 370 needs one block row
        # less, a plain refit; 370 -> 366 keeps the blocks and truncates the kept factor
        sur.fit(X[:n0], y[:n0])
        assert getattr(sur, "truncations", 0) == 0
        sur.fit(X[:n0 + 3], y[:n0 + 3])
        sur.fit(X[:n0 - 4], y[:n0 - 4])
        assert sur.truncations == 1
        fresh = make()
        fresh.fit(X[:n0 - 4], y[:n0 - 4])
        for a, b in zip(sur.predict_diag(xs), fresh.predict_diag(xs)):
            np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-10 * np.abs(b).max())
        sur.fit(X[:n0], y[:n0])              # neither a one-point growth nor a subset: refit
        fresh = make()
        fresh.fit(X[:n0], y[:n0])
        assert np.array_equal(sur.predict_diag(xs)[0], fresh.predict_diag(xs)[0])
        # a changed point in the middle is not an append
        Xc = X[:n0 + 1].copy()
        Xc[5] += 0.01
        before = sur.appended_rows
        sur.fit(Xc, y[:n0 + 1])
        assert sur.appended_rows == before
