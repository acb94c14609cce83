"""Randomised shapes through every kernel family: throughput path, latency path, gradient, pruned arg-min -- against the
oracle.  n straddles the 128-row block edges, d from 1 to 9, all four base kernels, ragged m.  -m gpu."""
import numpy as np
import pytest

from oracle import gp_oracle as O
from parity_util import check_mean, check_var, is_stated_tie, prior_var
from test_gpu_parity import native_for, select_path

pytestmark = pytest.mark.gpu

KINDS = [("rbf", 1.5), ("matern", 0.5), ("matern", 1.5), ("matern", 2.5)]


def random_case(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([1, 2, 3, 17, 127, 128, 129, 255, 256, 257, 300, 511, 513, 640]))
    d = int(rng.integers(1, 10))
    kind, nu = KINDS[seed % 4]
    X = rng.random((n, d))
    y = np.sin(2.0 * X.sum(1)) + 0.3 * rng.standard_normal(n)
    ard = rng.random() < 0.5
    ls = (0.4 + rng.random(d)) if ard else np.array([0.4 + rng.random()])
    spec = O.KernelSpec(kind=kind, nu=nu, length_scale=ls, amplitude=float(0.5 + 2 * rng.random()),
                        noise_level=float(rng.choice([0.0, 1e-3])))
    st = O.fit_state(X, y, spec, float(rng.choice([1e-6, 1e-4])), normalize_y=bool(rng.integers(0, 2)))
    m = int(rng.choice([1, 7, 8, 9, 100, 129, 333, 600]))
    xs = np.concatenate([rng.random((m, d)), X[: min(n, 3)] + 1e-3])[:m] if m > 3 else rng.random((m, d))
    return st, xs, float(y.min()), rng


@pytest.mark.parametrize("seed", range(40))
def test_random_shape(seed):
    st, xs, eta, rng = random_case(seed)
    gp = native_for(st, "f64")
    acq = ["lcb", "ei", "poi"][seed % 3]
    o_mean, o_var, o_a, (o_idx, _) = O.acquisition_sweep(st, acq, xs, eta=eta, kappa=2.0)
    dot_cond = st.y_std * (np.abs(O.kernel_cross(st.kernel, xs, st.X_train)) @ np.abs(st.alpha))
    outs = {}
    for path in ("sweep", "latency"):
        select_path(gp, path)
        out = gp.sweep(gp.candidates(xs), acq=acq, eta=eta, kappa=2.0, want_mean=True, want_var=True, want_acq=True,
                       want_min=True, index_base=11)
        mean, var, a = (out[k].cpu().numpy() for k in ("mean", "var", "acq"))
        err, bound = check_mean(mean, o_mean, st, "f64")
        assert (err <= bound + 16 * np.finfo(np.float64).eps * dot_cond).all(), (path, "mean")
        err, bound = check_var(var, o_var, st, "f64")
        assert (err <= bound).all(), (path, "var", float(np.max(err / bound)))
        idx = int(out["min_idx"].item()) - 11
        assert idx == int(np.argmin(a)), path
        if not np.isnan(a).any() and not np.isnan(o_a).any():
            assert is_stated_tie(o_a, o_idx, idx, "f64"), path
        outs[path] = (mean, var, a)
    np.testing.assert_allclose(outs["latency"][1], outs["sweep"][1], rtol=0, atol=1e-12 * prior_var(st))
    # gradient against the oracle where the variance is resolved and the kernel is smooth at the candidate
    val, grad, _, _ = (t.cpu().numpy() for t in gp.value_and_grad(gp.candidates(xs), acq, eta=eta, kappa=2.0))
    o_val, o_grad, _, _ = O.acquisition_value_and_grad(st, acq, xs, eta=eta, kappa=2.0)
    ok = (o_var > 1e-4 * prior_var(st)) & np.isfinite(o_grad).all(axis=1)
    if ok.any():
        scale = max(np.abs(o_grad[ok]).max(), 1e-300)
        cond = max(np.linalg.cond(st.L), 1.0)
        tol = 1e-6 * np.abs(o_grad[ok]) + 1e-13 * cond * scale + 1e-9 * scale
        assert (np.abs(grad[ok] - o_grad[ok]) <= tol).all(), float(np.max(np.abs(grad[ok] - o_grad[ok]) / tol))


@pytest.mark.parametrize("seed", range(8))
def test_random_pruned_argmin(seed):
    from bopy_b200 import _native
    st, _, eta, rng = random_case(100 + seed)
    gp = select_path(native_for(st, "f64"), "sweep")
    d = st.X_train.shape[1]
    m = int(rng.choice([40_000, 100_003]))
    xs = _native.candidates_uniform(seed, 0, m, -0.5 * np.ones(d), 1.5 * np.ones(d))
    acq = ["lcb", "ei", "poi"][seed % 3]
    full = gp.sweep(xs, acq=acq, eta=eta, kappa=2.0, want_min=True)
    minv, mini, stats = gp.argmin_pruned(xs, acq, eta=eta, kappa=2.0)
    if not np.isnan(float(full["min_val"].item())):
        assert int(mini.item()) == int(full["min_idx"].item()) and float(minv.item()) == float(full["min_val"].item()), stats


@pytest.mark.parametrize("m", [600, 2000])
def test_largest_dimensionality_through_every_batch_shape(m):
    """d = 32 (the supported maximum) with 16- and 32-candidate batches: the shared-memory budgets of the latency path
    and the gradient kernel at their tightest."""
    rng = np.random.default_rng(m)
    n, d = 257, 32
    X = rng.random((n, d))
    y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(n)
    spec = O.KernelSpec(kind="rbf", length_scale=1.5 + rng.random(d), amplitude=1.2)
    st = O.fit_state(X, y, spec, 1e-6, normalize_y=True)
    gp = select_path(native_for(st, "f64"), "latency")
    xs = rng.random((m, d))
    out = gp.sweep(gp.candidates(xs), acq="lcb", kappa=2.0, want_mean=True, want_var=True, want_acq=True, want_min=True)
    o_mean, o_var, o_a, (o_idx, _) = O.acquisition_sweep(st, "lcb", xs, kappa=2.0)
    err, bound = check_mean(out["mean"].cpu().numpy(), o_mean, st, "f64")
    dot_cond = st.y_std * (np.abs(O.kernel_cross(st.kernel, xs, st.X_train)) @ np.abs(st.alpha))
    assert (err <= bound + 16 * np.finfo(np.float64).eps * dot_cond).all()
    err, bound = check_var(out["var"].cpu().numpy(), o_var, st, "f64")
    assert (err <= bound).all()
    assert is_stated_tie(o_a, o_idx, int(out["min_idx"].item()), "f64")
    val, grad, _, _ = (t.cpu().numpy() for t in gp.value_and_grad(gp.candidates(xs), "lcb", kappa=2.0))
    _, o_grad, _, _ = O.acquisition_value_and_grad(st, "lcb", xs, kappa=2.0)
    scale = np.abs(o_grad).max()
    assert (np.abs(grad - o_grad) <= 1e-6 * np.abs(o_grad) + 1e-9 * scale).all()
