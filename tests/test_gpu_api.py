"""The reference's API-conformance tests (tests/test_surrogate.py, test_acquisiton.py, test_optimizer.py,
test_bayes_opt.py, test_callback.py of /root/reference) restated against bopy_b200 on the GPU, plus numeric
parity of the public API against the oracle.  -m gpu."""
import numpy as np
import pytest
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern

from bopy_b200.acquisition import EI, LCB, POI, KriggingBeliever, OneShotBatchAcquisitionFunction
from bopy_b200.bayes_opt import BayesOpt
from bopy_b200.benchmark_functions import forrester
from bopy_b200.bounds import Bound, Bounds
from bopy_b200.callback import Callback
from bopy_b200.initial_design import UniformRandomInitialDesign
from bopy_b200.optimizer import (CandidateSweepOptimizer, DirectOptimizer, OneShotBatchOptimizer,
                                 OneShotBatchOptimizerRandomSamplingStrategy, OptimizationResult,
                                 SequentialBatchOptimizer)
from bopy_b200.surrogate import ScipyGPSurrogate, Surrogate
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

n_samples = 10


@pytest.fixture(scope="module")
def x():
    return np.linspace(0, 1, n_samples).reshape(-1, 1)


@pytest.fixture(scope="module")
def y(x):
    return forrester(x)


@pytest.fixture(scope="module")
def trained_surrogate(x, y):
    s = ScipyGPSurrogate(gp=GaussianProcessRegressor(kernel=Matern(nu=1.5), alpha=1e-5, normalize_y=True))
    s.fit(x, y)
    return s


class TestSurrogateAfterFitting:
    def test_shapes_and_references(self, trained_surrogate, x, y):
        mean, cov = trained_surrogate.predict(x)
        assert mean.shape == (n_samples,) and cov.shape == (n_samples, n_samples)
        assert isinstance(mean, np.ndarray) and isinstance(cov, np.ndarray)
        assert trained_surrogate.x is x and trained_surrogate.y is y

    def test_predict_matches_sklearn(self, trained_surrogate):
        grid = np.linspace(0, 1, 200).reshape(-1, 1)
        mean, cov = trained_surrogate.predict(grid)
        r_mean, r_cov = trained_surrogate.gp.predict(grid, return_cov=True)
        scale = float(np.ravel(trained_surrogate.gp._y_train_std)[0])
        np.testing.assert_allclose(mean, r_mean, rtol=1e-9, atol=1e-9 * scale)
        np.testing.assert_allclose(cov, r_cov, rtol=1e-9, atol=1e-11 * scale ** 2)
        d_mean, d_var = trained_surrogate.predict_diag(grid)
        np.testing.assert_allclose(d_var, np.diag(r_cov), rtol=1e-9, atol=1e-11 * scale ** 2)
        assert np.array_equal(d_mean, mean)

    def test_validation_after_fit(self, trained_surrogate):
        with pytest.raises(ValueError, match="`x` must contain at least one sample"):
            trained_surrogate.predict(x=np.array([]))
        with pytest.raises(ValueError, match="`x` must be 2D"):
            trained_surrogate.predict(x=np.array([1.0]))
        with pytest.raises(ValueError, match="`x` must have the same number of dimensions as the training data"):
            trained_surrogate.predict(x=np.array([[1.0, 1.0]]))

    def test_refit_with_more_data_rebuilds_the_device_state(self, x, y):
        s = ScipyGPSurrogate(gp=GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(0.2), alpha=1e-8,
                                                         normalize_y=True, optimizer=None))
        s.fit(x[:6], y[:6])
        m6, _ = s.predict(x)
        s.fit(x, y)
        m10, _ = s.predict(x)
        np.testing.assert_allclose(m10, s.gp.predict(x), rtol=1e-9, atol=1e-9)
        assert not np.allclose(m6, m10)


@pytest.fixture(scope="module", params=[LCB, EI, POI], ids=["LCB", "EI", "POI"])
def trained_acquisition(request):
    xx = np.linspace(-np.pi, np.pi, n_samples).reshape(-1, 1)
    yy = np.sin(xx).flatten()
    s = ScipyGPSurrogate(gp=GaussianProcessRegressor(kernel=Matern()))
    s.fit(xx, yy)
    a = request.param(surrogate=s)
    a.fit(xx, yy)
    return a, xx, yy


class TestAcquisition:
    def test_output_dimensions(self, trained_acquisition):
        a, xx, _ = trained_acquisition
        out = a(xx)
        assert out.shape == (n_samples,) and isinstance(out, np.ndarray)

    def test_values_match_the_reference_formulas(self, trained_acquisition):
        a, xx, yy = trained_acquisition
        grid = np.linspace(-np.pi, np.pi, 501).reshape(-1, 1)
        got = a(grid)
        st = O.state_from_sklearn(a.surrogate.gp)
        mean, var = O.posterior_diag(st, grid)
        ref = O.acquisition(a.kind, mean, var, eta=float(np.min(yy)), kappa=2.0)
        resolved = var > 1e-7
        np.testing.assert_allclose(got[resolved], ref[resolved], rtol=1e-6, atol=1e-9)
        idx, val = a.argmin(grid)
        assert idx == int(np.argmin(got)) and val == got[idx]

    def test_foreign_surrogate_uses_the_device_epilogue(self, trained_acquisition):
        a, xx, yy = trained_acquisition

        class HostSurrogate(Surrogate):
            def __init__(self, gp):
                super().__init__()
                self.gp = gp

            def _fit(self, x, y):
                pass

            def _predict(self, x):
                return self.gp.predict(x, return_cov=True)

        host = HostSurrogate(a.surrogate.gp)
        host.fit(xx, yy)
        b = type(a)(host)
        b.fit(xx, yy)
        grid = np.linspace(-np.pi, np.pi, 97).reshape(-1, 1)
        got, want = b(grid), a(grid)
        mean, cov = host.predict(grid)
        resolved = np.diag(cov) > 1e-7
        np.testing.assert_allclose(got[resolved], want[resolved], rtol=1e-6, atol=1e-9)
        assert b.argmin(grid)[0] == int(np.argmin(got))


def forrester_setup(surrogate_dtype="f64"):
    bounds = Bounds(bounds=[Bound(lower=0.0, upper=1.0)])
    sur = ScipyGPSurrogate(gp=GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(0.2), alpha=1e-8,
                                                       normalize_y=True, optimizer=None), dtype=surrogate_dtype)
    return bounds, sur


class TestOptimizers:
    @pytest.fixture(scope="class")
    def fitted(self):
        bounds, sur = forrester_setup()
        xx = np.linspace(0, 1, 8).reshape(-1, 1)
        yy = forrester(xx)
        sur.fit(xx, yy)
        acq = LCB(sur, kappa=2.0)
        acq.fit(xx, yy)
        return bounds, sur, acq, xx, yy

    def test_candidate_sweep_optimizer(self, fitted):
        bounds, sur, acq, _, _ = fitted
        res = CandidateSweepOptimizer(acq, bounds, n_candidates=1 << 16, seed=3).optimize()
        assert isinstance(res, OptimizationResult)
        assert isinstance(res.x_min, np.ndarray) and res.x_min.shape == (1, 1)
        assert isinstance(res.f_min, np.ndarray) and res.f_min.shape == (1,)
        grid = np.linspace(0, 1, 20001).reshape(-1, 1)
        truth = acq(grid)
        assert res.f_min[0] <= truth.min() + 1e-3 * np.ptp(truth)
        assert abs(acq(res.x_min)[0] - res.f_min[0]) <= 1e-12 * max(1.0, abs(res.f_min[0]))
        zoomed = CandidateSweepOptimizer(acq, bounds, n_candidates=1 << 12, zoom_rounds=3, seed=3).optimize()
        coarse = CandidateSweepOptimizer(acq, bounds, n_candidates=1 << 12, seed=3).optimize()
        assert zoomed.f_min[0] <= coarse.f_min[0]
        assert zoomed.f_min[0] <= truth.min() + 1e-6 * np.ptp(truth)

    def test_direct_optimizer(self, fitted):
        bounds, sur, acq, _, _ = fitted
        res = DirectOptimizer(acq, bounds, maxf=100).optimize()
        assert res.x_min.shape == (1, 1) and res.f_min.shape == (1,)
        assert 0.0 <= res.x_min[0, 0] <= 1.0

    def test_sequential_batch_with_kriging_believer(self, fitted):
        bounds, sur, _, xx, yy = fitted
        kb = KriggingBeliever(LCB(sur))
        kb.fit(xx, yy)
        base = CandidateSweepOptimizer(kb, bounds, n_candidates=1 << 12)
        res = SequentialBatchOptimizer(kb, bounds, base_optimizer=base, batch_size=2).optimize()
        assert res.x_min.shape == (2, 1) and res.f_min.shape == (2,)
        assert len(sur.x) == len(xx)
        assert abs(res.x_min[0, 0] - res.x_min[1, 0]) > 1e-6      # the believer moved the second pick

    def test_one_shot_batch(self, fitted):
        bounds, sur, _, xx, yy = fitted
        acq = OneShotBatchAcquisitionFunction(LCB(sur))
        acq.fit(xx, yy)
        res = OneShotBatchOptimizer(acq, bounds, base_optimizer=DirectOptimizer(acq, bounds, maxf=100), batch_size=2,
                                    strategy=OneShotBatchOptimizerRandomSamplingStrategy()).optimize()
        assert res.x_min.shape == (2, 1) and res.f_min.shape == (2,)


class Recorder(Callback):
    def __init__(self):
        self.events = []

    def on_initial_design_end(self, bo):
        self.events.append("on_initial_design_end")

    def on_acquisition_optimized(self, bo, opt_result):
        assert isinstance(opt_result, OptimizationResult)
        self.events.append("on_acquisition_optimized")

    def on_surrogate_updated(self, bo):
        self.events.append("on_surrogate_updated")

    def on_acquisition_updated(self, bo):
        self.events.append("on_acquisition_updated")

    def on_trial_end(self, bo):
        self.events.append("on_trial_end")

    def on_bo_end(self, bo):
        self.events.append("on_bo_end")


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_bayes_opt_finds_the_forrester_minimum(dtype):
    np.random.seed(0)
    bounds, sur = forrester_setup(dtype)
    acq = LCB(sur, kappa=2.0)
    rec = Recorder()
    bo = BayesOpt(objective_function=forrester, surrogate=sur, acquisition_function=acq,
                  optimizer=CandidateSweepOptimizer(acq, bounds, n_candidates=1 << 14, zoom_rounds=2, seed=1),
                  initial_design=UniformRandomInitialDesign(), bounds=bounds, callbacks=[rec])
    res = bo.run(n_trials=12, n_initial_design=5)
    assert res.x_opt.shape == (1, 1) and len(res.trial_results) == 12
    assert res.f_opt < -5.9 and abs(res.x_opt[0, 0] - 0.757249) < 0.02       # global minimum -6.0207 at 0.7572
    assert rec.events[0] == "on_initial_design_end" and rec.events[-1] == "on_bo_end"
    assert rec.events.count("on_trial_end") == 12 and rec.events.count("on_acquisition_optimized") == 12


class TestMultiStart:
    def test_segment_argmin_matches_numpy(self):
        from bopy_b200 import _native
        from conftest import golden_state
        from test_gpu_parity import native_for
        g, st = golden_state("c3_branin_n256")
        gp = native_for(st, "f64")
        xs = _native.candidates_uniform(5, 0, 128 * 37 + 11, [-5.0, 0.0], [10.0, 15.0])
        eta = float(g["eta"])
        a = gp.sweep(xs, acq="ei", eta=eta, want_acq=True)["acq"].cpu().numpy()
        for seg in (128, 384, 1280):
            vals, idxs = gp.segment_argmin(xs, seg, "ei", eta=eta, index_base=7000)
            vals, idxs = vals.cpu().numpy(), idxs.cpu().numpy()
            nseg = -(-len(a) // seg)
            assert vals.shape == (nseg,)
            for s in range(nseg):
                chunk = a[s * seg:(s + 1) * seg]
                assert idxs[s] - 7000 == s * seg + int(np.argmin(chunk)) and vals[s] == chunk.min()
        rows = _native.gather_rows(xs, torch_as(idxs, xs.device), index_base=7000).cpu().numpy()
        assert np.array_equal(rows, xs.cpu().numpy()[idxs - 7000])

    def test_candidate_clouds(self):
        from bopy_b200 import _native
        import torch
        starts = torch.tensor([[0.1, 0.9], [0.5, 0.5], [1.0, 0.0]], dtype=torch.float64, device="cuda")
        cloud = _native.candidates_around(3, starts, 256, [0.05, 0.2], [0.0, 0.0], [1.0, 1.0]).cpu().numpy()
        assert cloud.shape == (768, 2)
        for s in range(3):
            block = cloud[s * 256:(s + 1) * 256]
            assert np.array_equal(block[0], starts[s].cpu().numpy())          # the incumbent is point 0
            assert (block >= 0).all() and (block <= 1).all()
            assert (np.abs(block - starts[s].cpu().numpy()) <= np.array([0.05, 0.2]) + 1e-15).all()
            assert block[:, 0].std() > 0.01

    @pytest.mark.parametrize("method", ["gradient", "cloud"])
    def test_multi_start_optimizer_beats_the_plain_sweep(self, method):
        from bopy_b200.benchmark_functions import branin
        from bopy_b200.optimizer import MultiStartOptimizer
        lo, hi = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
        X = lo + np.random.default_rng(3).random((40, 2)) * (hi - lo)
        y = branin(X)
        sur = ScipyGPSurrogate(GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF([3.0, 3.0]), alpha=1e-6,
                                                        normalize_y=True, optimizer=None))
        sur.fit(X, y)
        acq = LCB(sur, kappa=1.0)
        acq.fit(X, y)
        bounds = Bounds([Bound(-5.0, 10.0), Bound(0.0, 15.0)])
        plain = CandidateSweepOptimizer(acq, bounds, n_candidates=1 << 14, seed=11).optimize()
        ms = MultiStartOptimizer(acq, bounds, n_starts=64, n_candidates=1 << 14, rounds=8, seed=11, method=method)
        res = ms.optimize()
        assert res.x_min.shape == (1, 2) and res.f_min.shape == (1,)
        assert res.f_min[0] <= plain.f_min[0] + 1e-12
        assert abs(acq(res.x_min)[0] - res.f_min[0]) <= 1e-10 * max(1.0, abs(res.f_min[0]))
        xs, vs = ms.local_minima()
        assert xs.shape == (64, 2) and vs.shape == (64,) and vs.min() == res.f_min[0]
        dense = acq(np.stack(np.meshgrid(np.linspace(-5, 10, 301), np.linspace(0, 15, 301)), -1).reshape(-1, 2))
        assert res.f_min[0] <= dense.min() + 1e-6 * np.ptp(dense)            # at least as good as a 300 x 300 grid


def torch_as(a, device):
    import torch
    return torch.as_tensor(a, dtype=torch.int64, device=device)


class FakeGPyModel:
    """Duck-typed stand-in for GPy.models.GPRegression (GPy is not installed): exposes exactly the attributes
    GPyGPSurrogate reads, computed with the oracle from the same definitions GPy uses."""

    class _Kern:
        name = "rbf"

        def __init__(self, variance, lengthscale):
            self.variance, self.lengthscale = np.array([variance]), np.array([lengthscale])

    class _Norm:
        def __init__(self, y):
            self.mean, self.std = y.mean(), y.std()

    class _Posterior:
        pass

    def __init__(self, x, y, variance=1.3, lengthscale=0.25, noise_var=1e-5):
        self.kern = self._Kern(variance, lengthscale)
        self.noise_var = noise_var
        self.optimized = 0
        self.set_XY(x, y)

    def set_XY(self, x, y):
        self.X, self.Y = x, y
        self.normalizer = self._Norm(y)
        spec = O.KernelSpec(kind="rbf", length_scale=self.kern.lengthscale, amplitude=float(self.kern.variance[0]))
        st = O.fit_state(x, y.ravel(), spec, self.noise_var, normalize_y=True)
        self.posterior = self._Posterior()
        self.posterior.woodbury_chol, self.posterior.woodbury_vector = st.L, st.alpha[:, None]
        self._state = st

    def optimize_restarts(self, n):
        self.optimized += n

    def predict_noiseless(self, x, full_cov=True):
        mean, cov = O.posterior_full(self._state, x)
        return mean[:, None], cov


def test_gpy_surrogate_with_a_duck_typed_model():
    from bopy_b200.surrogate import GPyGPSurrogate
    xx = np.linspace(0, 1, 10).reshape(-1, 1)
    yy = forrester(xx)
    sur = GPyGPSurrogate(gp_initializer=lambda x, y: FakeGPyModel(x, y), n_restarts=2)
    with pytest.raises(Exception, match="must be fitted first"):
        sur.predict(xx)
    sur.fit(xx, yy)
    assert sur.gp.optimized == 2 and sur.x is xx
    grid = np.linspace(0, 1, 77).reshape(-1, 1)
    mean, cov = sur.predict(grid)
    ref_mean, ref_cov = sur.gp.predict_noiseless(grid, full_cov=True)
    np.testing.assert_allclose(mean, ref_mean.ravel(), rtol=1e-9, atol=1e-9 * yy.std())
    np.testing.assert_allclose(cov, ref_cov, rtol=1e-8, atol=1e-11 * yy.var() * 1.3)
    sur.fit(xx[:7], yy[:7])                      # refit goes through set_XY and a new device state
    assert sur.gp.optimized == 4 and len(sur.gp.X) == 7
    lcb = LCB(sur)
    lcb.fit(xx[:7], yy[:7])
    assert lcb(grid).shape == (77,)


def test_examples_run_end_to_end():
    """BASELINE configs C1 / C2 (examples/example_1d.py, examples/example_batch_1d.py)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = []
    for name in ("example_1d", "example_batch_1d"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, "examples", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        np.random.seed(0)
        res = mod.main(out=lines.append)
        assert res.x_opt.shape == (1, 1) and 0.0 <= res.x_opt[0, 0] <= 1.0
        assert res.f_opt <= -0.9                      # both find at least the Forrester local basin
    one_d = importlib.util.spec_from_file_location("e1", os.path.join(root, "examples", "example_1d.py"))
    assert any("optimum found" in l for l in lines) and any("batch proposed" in l for l in lines)


def test_one_shot_batch_from_a_device_sweep():
    """the one-shot strategies choose from the evaluations one global pass logged
    (bopy/optimizer.py:271-276) -- here the pass is a fused device sweep of 2^17 candidates and the log stays on
    the device; the top-k strategy returns the k best evaluations that keep their distance."""
    from bopy_b200.optimizer import OneShotBatchOptimizerTopKStrategy
    rng = np.random.default_rng(12)
    X = rng.random((200, 2))
    y = np.sin(5 * X[:, 0]) * np.cos(4 * X[:, 1])
    sur = ScipyGPSurrogate(GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF([0.2, 0.2]), alpha=1e-6,
                                                    normalize_y=True, optimizer=None))
    sur.fit(X, y)
    acq = OneShotBatchAcquisitionFunction(LCB(sur))
    acq.fit(X, y)
    bounds = Bounds([Bound(0.0, 1.0), Bound(0.0, 1.0)])
    base = CandidateSweepOptimizer(acq, bounds, n_candidates=1 << 17, seed=2)
    strategy = OneShotBatchOptimizerTopKStrategy(min_distance=0.15)
    res = OneShotBatchOptimizer(acq, bounds, base_optimizer=base, batch_size=5, strategy=strategy).optimize()
    assert res.x_min.shape == (5, 2) and res.f_min.shape == (5,)
    xs, a_xs = acq.get_evaluations()
    assert xs.shape == (1 << 17, 2) and a_xs.shape == (1 << 17,)
    assert res.f_min[0] == a_xs.min() and (np.diff(res.f_min) >= 0).all()
    d = np.linalg.norm(res.x_min[:, None, :] - res.x_min[None, :, :], axis=2)
    assert (d[np.triu_indices(5, 1)] >= 0.15).all()
    hx, hf = strategy.select(xs, a_xs, 5)                       # the host rule picks the same batch
    assert np.array_equal(hx, res.x_min) and np.array_equal(hf, res.f_min)
    rnd = OneShotBatchOptimizer(acq, bounds, base_optimizer=base, batch_size=3,
                                strategy=OneShotBatchOptimizerRandomSamplingStrategy()).optimize()
    assert rnd.x_min.shape == (3, 2)


def test_hartmann6_example_runs_end_to_end():
    """BASELINE config C4 in miniature (examples/example_hartmann6.py): pruned sweeps, one-row appends of the factor,
    a Kriging-believer batch with its truncation, the gradient multi-start."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("example_hartmann6", os.path.join(root, "examples", "example_hartmann6.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lines = []
    result, batch, ms = mod.main(n_initial=300, n_trials=6, out=lines.append)
    assert result.x_opt.shape == (1, 6) and len(result.trial_results) == 6
    assert result.f_opt <= min(r.f_opt_so_far for r in [result.initial_design_result]) + 1e-12
    assert batch.x_min.shape == (4, 6) and batch.f_min.shape == (4,)
    assert len({tuple(np.round(x, 9)) for x in batch.x_min}) == 4          # the believer moves on after each pick
    assert ms.x_min.shape == (1, 6) and ms.f_min[0] <= 0.0
    text = "\n".join(lines)
    import re
    assert int(re.search(r"one-row appends: (\d+)", text).group(1)) >= 5       # every trial but a block-edge one
    assert "truncations 1" in text
