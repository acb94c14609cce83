"""Branch-and-bound arg-min (bopy_acq_argmin_pruned): same index and value as the plain fused arg-min, with only a
fraction of the candidates going through the full posterior.  -m gpu."""
import numpy as np
import pytest

from test_gpu_parity import cached_native

pytestmark = pytest.mark.gpu

BOXES = {"c4_hartmann6_n2048": (np.zeros(6), np.ones(6)), "c3_branin_n256": (np.array([-5.0, 0.0]), np.array([10.0, 15.0])),
         "matern15_d2": None, "matern05_d2": None, "ard_amp_white": None, "ragged_n333_d4_opt": None}


def box_for(name, st):
    if BOXES[name] is not None:
        return BOXES[name]
    lo, hi = st.X_train.min(0), st.X_train.max(0)
    pad = 0.25 * (hi - lo)
    return lo - pad, hi + pad          # reaches into the far field, where the bound prunes everything


@pytest.mark.parametrize("acq", ["lcb", "ei", "poi"])
@pytest.mark.parametrize("name", list(BOXES))
def test_pruned_argmin_is_the_plain_argmin(name, acq):
    from bopy_b200 import _native
    g, st, gp = cached_native(name, "f64", "sweep")
    lo, hi = box_for(name, st)
    eta = float(g["eta"])
    for seed, m, base in ((3, 200_003, 0), (4, 1 << 19, 1 << 33)):
        xs = _native.candidates_uniform(seed, 0, m, lo, hi)
        full = gp.sweep(xs, acq=acq, eta=eta, kappa=2.0, want_min=True, index_base=base)
        minv, mini, stats = gp.argmin_pruned(xs, acq, eta=eta, kappa=2.0, index_base=base)
        f_val, f_idx = float(full["min_val"].item()), int(full["min_idx"].item())
        if np.isnan(f_val):
            continue                    # NaN acquisition values are the documented exception
        assert int(mini.item()) == f_idx and float(minv.item()) == f_val, (name, acq, stats)
        assert stats["candidates"] == m and stats["swept"] <= m


def test_prunes_most_candidates_on_the_headline_problem():
    from bopy_b200 import _native
    g, st, gp = cached_native("c4_hartmann6_n2048", "f64", "sweep")
    xs = _native.candidates_uniform(1235, 0, 1 << 20, np.zeros(6), np.ones(6))
    eta = float(g["eta"])
    for acq in ("ei", "lcb"):
        full = gp.sweep(xs, acq=acq, eta=eta, kappa=2.0, want_min=True)
        minv, mini, stats = gp.argmin_pruned(xs, acq, eta=eta, kappa=2.0)
        assert int(mini.item()) == int(full["min_idx"].item()) and float(minv.item()) == float(full["min_val"].item())
        assert stats["swept"] < 0.2 * stats["candidates"], stats


def test_small_sets_fall_back_to_the_plain_sweep():
    g, st, gp = cached_native("c3_branin_n256", "f64", "sweep")
    xs = gp.candidates(g["Xs"])
    minv, mini, stats = gp.argmin_pruned(xs, "ei", eta=float(g["eta"]), index_base=5)
    full = gp.sweep(xs, acq="ei", eta=float(g["eta"]), want_min=True, index_base=5)
    assert stats["swept"] == stats["candidates"] == len(g["Xs"]) and stats["sample"] == 0
    assert int(mini.item()) == int(full["min_idx"].item())


def test_fp32_handles_and_the_public_api():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel

    from bopy_b200.acquisition import EI
    from bopy_b200.benchmark_functions import hartmann6
    from bopy_b200.bounds import Bound, Bounds
    from bopy_b200.optimizer import CandidateSweepOptimizer
    from bopy_b200.surrogate import B200GPSurrogate
    rng = np.random.default_rng(0)
    X = rng.random((600, 6))
    y = hartmann6(X)
    for dtype in ("f64", "f32"):
        sur = B200GPSurrogate(GaussianProcessRegressor(ConstantKernel(1.0) * RBF(0.3 * np.ones(6)), alpha=1e-6,
                                                        normalize_y=True, optimizer=None), dtype=dtype)
        sur.fit(X, y)
        acq = EI(sur)
        acq.fit(X, y)
        bounds = Bounds([Bound(0.0, 1.0)] * 6)
        plain = CandidateSweepOptimizer(acq, bounds, n_candidates=1 << 18, seed=5).optimize()
        pruned = CandidateSweepOptimizer(acq, bounds, n_candidates=1 << 18, seed=5, prune=True).optimize()
        assert np.array_equal(plain.x_min, pruned.x_min) and np.array_equal(plain.f_min, pruned.f_min)
        assert sur.last_prune_stats["swept"] < sur.last_prune_stats["candidates"]


@pytest.mark.parametrize("acq", ["lcb", "ei", "poi"])
@pytest.mark.parametrize("name", ["c4_hartmann6_n2048", "c3_branin_n256", "matern15_d2"])
def test_pruned_segment_argmin_is_the_plain_segment_argmin(name, acq):
    from bopy_b200 import _native
    g, st, gp = cached_native(name, "f64", "sweep")
    lo, hi = box_for(name, st)
    eta = float(g["eta"])
    for seg, nseg in ((128, 300), (1024, 256)):
        xs = _native.candidates_uniform(seg, 0, seg * nseg, lo, hi)
        v0, i0 = gp.segment_argmin(xs, seg, acq, eta=eta, kappa=2.0, index_base=1000)
        v1, i1, stats = gp.segment_argmin_pruned(xs, seg, acq, eta=eta, kappa=2.0, index_base=1000)
        v0, i0, v1, i1 = (t.cpu().numpy() for t in (v0, i0, v1, i1))
        ok = ~np.isnan(v0)                      # NaN acquisition values are the documented exception
        assert np.array_equal(i0[ok], i1[ok]) and np.array_equal(v0[ok], v1[ok]), (name, acq, seg, stats)


def test_multistart_with_pruned_global_sweep_returns_the_same_result():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel

    from bopy_b200.acquisition import EI
    from bopy_b200.benchmark_functions import hartmann6
    from bopy_b200.bounds import Bound, Bounds
    from bopy_b200.optimizer import MultiStartOptimizer
    from bopy_b200.surrogate import B200GPSurrogate
    rng = np.random.default_rng(1)
    X = rng.random((700, 6))
    y = hartmann6(X)
    sur = B200GPSurrogate(GaussianProcessRegressor(ConstantKernel(1.0) * RBF(0.3 * np.ones(6)), alpha=1e-6,
                                                    normalize_y=True, optimizer=None))
    sur.fit(X, y)
    from bopy_b200.acquisition import LCB
    bounds = Bounds([Bound(0.0, 1.0)] * 6)
    for acq in (EI(sur), LCB(sur)):
        acq.fit(X, y)
        a = MultiStartOptimizer(acq, bounds, n_starts=128, n_candidates=1 << 17, seed=9, method="gradient").optimize()
        b = MultiStartOptimizer(acq, bounds, n_starts=128, n_candidates=1 << 17, seed=9, method="gradient",
                                prune=True).optimize()
        assert np.array_equal(a.x_min, b.x_min) and np.array_equal(a.f_min, b.f_min)
    # per-segment incumbents are weak for EI (most segments' best sample is ~0), strong for LCB
    assert sur.last_prune_stats["swept"] < sur.last_prune_stats["candidates"]
