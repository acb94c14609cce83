"""Host-side behaviour of the drop-in API that needs no GPU: argument validation (messages verbatim from the
reference, bopy/mixin.py:34-64 / tests/test_surrogate.py:55-101), bounds, designs, objectives, the BayesOpt loop
and its callbacks driven with CPU test doubles, kernel flattening, loud failure without CUDA."""
import numpy as np
import pytest
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, DotProduct, Matern, WhiteKernel

import bopy_b200
from bopy_b200.acquisition import EI, LCB, POI, AcquisitionFunction, KriggingBeliever, OneShotBatchAcquisitionFunction
from bopy_b200.bayes_opt import BayesOpt, BOResult
from bopy_b200.benchmark_functions import bohachevsky, branin, forrester, hartmann6
from bopy_b200.bounds import Bound, Bounds
from bopy_b200.callback import EVENTS, Callback
from bopy_b200.exceptions import NativeLibraryError, NotFittedError
from bopy_b200.initial_design import (LatinHypercubeInitialDesign, SobolSequenceInitialDesign,
                                      UniformRandomInitialDesign)
from bopy_b200.kernel_spec import UnsupportedKernelError, flatten_sklearn_kernel
from bopy_b200.optimizer import (OneShotBatchOptimizer, OneShotBatchOptimizerRandomSamplingStrategy,
                                 OptimizationResult, Optimizer, SequentialBatchOptimizer)
from bopy_b200.surrogate import B200GPSurrogate, ScipyGPSurrogate, Surrogate


def make_surrogate():
    return ScipyGPSurrogate(gp=GaussianProcessRegressor(kernel=Matern(nu=1.5), alpha=1e-5, normalize_y=True))


class TestArgumentsToFit:
    @pytest.mark.parametrize("x, y, message", [
        (np.array([]), np.array([1.0]), "`x` must contain at least one sample"),
        (np.array([[1.0]]), np.array([]), "`y` must contain at least one sample"),
        (np.array([[1.0]]), np.array([1.0, 1.0]), "`x` and `y` must have the same number of samples"),
        (np.array([[[1.0]]]), np.array([1.0]), "`x` must be 2D"),
        (np.array([[1.0]]), np.array([[1.0]]), "`y` must be 1D"),
    ])
    def test_bad_fit_arguments_raise_before_anything_runs(self, x, y, message):
        with pytest.raises(ValueError, match=message):
            make_surrogate().fit(x=x, y=y)
        for cls in (LCB, EI, POI):
            with pytest.raises(ValueError, match=message):
                cls(make_surrogate()).fit(x, y)


def test_predict_before_fit_raises_not_fitted():
    x = np.linspace(0, 1, 10).reshape(-1, 1)
    with pytest.raises(NotFittedError, match="must be fitted first"):
        make_surrogate().predict(x)
    for cls in (LCB, EI, POI):
        with pytest.raises(NotFittedError, match="must be fitted first"):
            cls(make_surrogate())(x)


def test_predict_argument_validation_after_fit():
    sur = make_surrogate()
    sur.has_been_fitted, sur.n_dimensions = True, 1   # validation is host-side and precedes any device work
    with pytest.raises(ValueError, match="`x` must contain at least one sample"):
        sur.predict(np.array([]))
    with pytest.raises(ValueError, match="`x` must be 2D"):
        sur.predict(np.array([1.0]))
    with pytest.raises(ValueError, match="`x` must have the same number of dimensions as the training data"):
        sur.predict(np.array([[1.0, 1.0]]))


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    x = np.linspace(0, 1, 10).reshape(-1, 1)
    with pytest.raises(NativeLibraryError, match="no CPU fallback"):
        make_surrogate().fit(x, forrester(x))


def test_drop_in_names_and_constructors():
    assert issubclass(ScipyGPSurrogate, B200GPSurrogate) and issubclass(B200GPSurrogate, Surrogate)
    sur = make_surrogate()
    assert LCB(sur).kappa == 2.0 and LCB(sur, kappa=0.5).kappa == 0.5
    assert EI(sur)._eta == np.inf and POI(sur)._eta == np.inf
    for mod in ("acquisition", "bayes_opt", "benchmark_functions", "bounds", "callback", "exceptions",
                "initial_design", "mixin", "optimizer", "surrogate"):
        assert hasattr(bopy_b200, mod)


def test_bounds():
    with pytest.raises(ValueError, match="`lower` must be less than `upper`"):
        Bound(lower=1.0, upper=0.0)
    with pytest.raises(ValueError, match="`lower` must be less than `upper`"):
        Bound(lower=1.0, upper=1.0)
    with pytest.raises(ValueError, match="`bounds` must contain at least one bound."):
        Bounds(bounds=[])
    b = Bounds(bounds=[Bound(0.0, 1.0), Bound(-2.0, 3.0)])
    assert b.n_dimensions == 2 and b.lowers == [0.0, -2.0] and b.uppers == [1.0, 3.0]


@pytest.mark.parametrize("design", [UniformRandomInitialDesign(), SobolSequenceInitialDesign(),
                                    LatinHypercubeInitialDesign()])
def test_initial_designs(design):
    bounds = Bounds(bounds=[Bound(-1.0, 2.0), Bound(10.0, 11.0)])
    with pytest.raises(ValueError, match="`n_points` must be positive."):
        design.generate(bounds, 0)
    pts = design.generate(bounds, 10)
    assert pts.shape == (10, 2)
    assert (pts >= np.array(bounds.lowers)).all() and (pts <= np.array(bounds.uppers)).all()


def test_benchmark_functions():
    assert forrester(np.array([[0.757249]]))[0] == pytest.approx(-6.02074, abs=1e-5)
    assert bohachevsky(np.zeros((1, 2)))[0] == pytest.approx(0.0, abs=1e-12)
    assert branin(np.array([[np.pi, 2.275], [-np.pi, 12.275], [9.42478, 2.475]])) == pytest.approx(0.397887, abs=1e-5)
    assert hartmann6(np.array([[0.20169, 0.150011, 0.476874, 0.275332, 0.311652, 0.6573]]))[0] == \
        pytest.approx(-3.32237, abs=1e-5)
    assert forrester(np.zeros((4, 1))).shape == (4,) and hartmann6(np.zeros((3, 6))).shape == (3,)


def test_kernel_flattening():
    f = flatten_sklearn_kernel(ConstantKernel(2.5) * RBF([0.2, 0.7, 1.3]) + WhiteKernel(1e-3))
    assert (f.kernel, f.amplitude, f.noise_level) == ("rbf", 2.5, 1e-3) and f.length_scale.tolist() == [0.2, 0.7, 1.3]
    f = flatten_sklearn_kernel(Matern(length_scale=0.4, nu=2.5) * ConstantKernel(3.0) * ConstantKernel(0.5))
    assert (f.kernel, f.amplitude, f.noise_level) == ("matern52", 1.5, 0.0) and f.length_scale.tolist() == [0.4]
    assert flatten_sklearn_kernel(Matern(nu=0.5)).kernel == "matern12"
    assert flatten_sklearn_kernel(RBF()).kernel == "rbf"
    for bad in (DotProduct(), RBF() * RBF(), Matern(nu=3.5), RBF() + RBF(), ConstantKernel(1.0)):
        with pytest.raises(UnsupportedKernelError):
            flatten_sklearn_kernel(bad)


# ---- the loop, driven with CPU test doubles (host logic only) --------------------------------------------------
class NearestSurrogate(Surrogate):
    def _fit(self, x, y):
        pass

    def _predict(self, x):
        d = np.abs(x[:, None, 0] - self.x[None, :, 0])
        return self.y[np.argmin(d, axis=1)], np.diag(np.min(d, axis=1) + 1e-3)


class MeanAcquisition(AcquisitionFunction):
    def _f(self, x):
        mean, _ = self.surrogate.predict(x)
        return mean


class GridOptimizer(Optimizer):
    def _optimize(self):
        grid = np.linspace(self.bounds.lowers[0], self.bounds.uppers[0], 101).reshape(-1, 1)
        values = self.acquisition_function(grid)
        i = int(np.argmin(values))
        return grid[i:i + 1], values[i:i + 1]


class Recorder(Callback):
    def __init__(self):
        self.seen = []

    def __getattribute__(self, name):
        if name in EVENTS:
            return lambda *args: object.__getattribute__(self, "seen").append((name, len(args)))
        return object.__getattribute__(self, name)


def make_loop(callbacks=None, optimizer_factory=None):
    bounds = Bounds(bounds=[Bound(0.0, 1.0)])
    sur = NearestSurrogate()
    acq = MeanAcquisition(sur)
    opt = (optimizer_factory or GridOptimizer)(acq, bounds)
    return BayesOpt(objective_function=forrester, surrogate=sur, acquisition_function=acq, optimizer=opt,
                    initial_design=UniformRandomInitialDesign(), bounds=bounds, callbacks=callbacks)


def test_bayes_opt_loop_results_and_shapes():
    np.random.seed(0)
    bo = make_loop()
    res = bo.run(n_trials=3, n_initial_design=5)
    assert isinstance(res, BOResult)
    assert res.x_opt.shape == (1, 1) and np.isscalar(res.f_opt + 0.0)
    assert res.initial_design_result.x_selected.shape == (5, 1)
    assert res.initial_design_result.f_selected.shape == (5,)
    assert len(res.trial_results) == 3
    for t in res.trial_results:
        assert t.x_selected.shape == (1, 1) and t.f_selected.shape == (1,) and t.x_opt_so_far.shape == (1, 1)
    assert bo.x.shape == (8, 1) and bo.y.shape == (8,)
    assert res.f_opt == bo.y.min()
    assert bo.surrogate.x is bo.x and bo.surrogate.y is bo.y          # references, not copies


def test_every_callback_fires_in_order():
    np.random.seed(0)
    rec = Recorder()
    make_loop(callbacks=[rec]).run(n_trials=1, n_initial_design=5)
    assert rec.seen == [("on_initial_design_end", 1), ("on_acquisition_optimized", 2), ("on_surrogate_updated", 1),
                        ("on_acquisition_updated", 1), ("on_trial_end", 1), ("on_bo_end", 1)]


def test_sequential_batch_optimizer_with_kriging_believer():
    np.random.seed(0)
    bounds = Bounds(bounds=[Bound(0.0, 1.0)])
    sur = NearestSurrogate()
    x = np.random.rand(6, 1)
    sur.fit(x, forrester(x))
    kb = KriggingBeliever(MeanAcquisition(sur))
    kb.fit(x, forrester(x))
    opt = SequentialBatchOptimizer(kb, bounds, base_optimizer=GridOptimizer(kb, bounds), batch_size=3)
    res = opt.optimize()
    assert isinstance(res, OptimizationResult) and res.x_min.shape == (3, 1) and res.f_min.shape == (3,)
    assert len(sur.x) == 6                                           # fantasies removed by finish_batch


def test_one_shot_batch_optimizer_logs_and_selects():
    np.random.seed(0)
    bounds = Bounds(bounds=[Bound(0.0, 1.0)])
    sur = NearestSurrogate()
    x = np.random.rand(6, 1)
    sur.fit(x, forrester(x))
    acq = OneShotBatchAcquisitionFunction(MeanAcquisition(sur))
    acq.fit(x, forrester(x))
    opt = OneShotBatchOptimizer(acq, bounds, base_optimizer=GridOptimizer(acq, bounds), batch_size=4,
                                strategy=OneShotBatchOptimizerRandomSamplingStrategy())
    res = opt.optimize()
    assert res.x_min.shape == (4, 1) and res.f_min.shape == (4,)
    xs, a_xs = acq.get_evaluations()
    assert xs.shape == (101, 1) and a_xs.shape == (101,)


def test_theta_gradient_follows_sklearn_hyperparameter_order():
    from bopy_b200.kernel_spec import theta_gradient
    flat = np.array([10.0, 1.0, 2.0, 3.0, 99.0])          # [amplitude, l0, l1, l2, noise]
    k = ConstantKernel(2.0) * RBF([0.1, 0.2, 0.3]) + WhiteKernel(1e-2)
    assert [h.name for h in k.hyperparameters] == ["k1__k1__constant_value", "k1__k2__length_scale", "k2__noise_level"]
    assert theta_gradient(k, flat).tolist() == [10.0, 1.0, 2.0, 3.0, 99.0]
    k = WhiteKernel(1e-2) + RBF([0.1, 0.2, 0.3]) * ConstantKernel(2.0)
    assert theta_gradient(k, flat).tolist() == [99.0, 1.0, 2.0, 3.0, 10.0]
    k = ConstantKernel(2.0, constant_value_bounds="fixed") * RBF([0.1, 0.2, 0.3])
    assert theta_gradient(k, flat).tolist() == [1.0, 2.0, 3.0]                 # fixed hyper-parameters drop out
    k = ConstantKernel(2.0) * ConstantKernel(3.0) * Matern(0.5, nu=1.5)
    assert theta_gradient(k, np.array([10.0, 7.0, 0.0])).tolist() == [10.0, 10.0, 7.0]
    with pytest.raises(UnsupportedKernelError):
        theta_gradient(ConstantKernel(2.0) * RBF([0.1, 0.2]), flat)            # layout mismatch


def test_bounds_helpers():
    from bopy_b200.bounds import unit_box
    b = Bounds([Bound(-1.0, 2.0), Bound(10.0, 11.0)])
    lo, hi = b.as_arrays()
    assert lo.tolist() == [-1.0, 10.0] and hi.tolist() == [2.0, 11.0] and len(b) == 2 and b[1] == Bound(10.0, 11.0)
    x = np.array([[0.0, 10.5], [-2.0, 10.5], [0.0, 12.0]])
    assert b.contains(x).tolist() == [True, False, False]
    assert np.array_equal(b.clip(x), [[0.0, 10.5], [-1.0, 10.5], [0.0, 11.0]])
    assert b == Bounds([Bound(-1.0, 2.0), Bound(10.0, 11.0)]) and b != unit_box(2)
    assert unit_box(3).lowers == [0.0, 0.0, 0.0] and Bound(0.0, 2.5).width == 2.5
    with pytest.raises(ValueError, match="`lower` must be less than `upper`"):
        Bound(float("nan"), 1.0)


def test_top_k_strategy_host_rule():
    from bopy_b200.optimizer import OneShotBatchOptimizerTopKStrategy
    x = np.array([[0.0, 0.0], [0.01, 0.0], [0.5, 0.5], [0.52, 0.5], [1.0, 1.0], [0.2, 0.9]])
    a = np.array([-5.0, -4.9, -4.0, -4.5, np.nan, -1.0])
    xs, fs = OneShotBatchOptimizerTopKStrategy(min_distance=0.1).select(x, a, 3)
    assert fs.tolist() == [-5.0, -4.5, -1.0]            # -4.9 and -4.0 sit too close to better picks; NaN never chosen
    assert np.array_equal(xs, x[[0, 3, 5]])
    xs, fs = OneShotBatchOptimizerTopKStrategy().select(x, a, 2)
    assert fs.tolist() == [-5.0, -4.9]
    xs, fs = OneShotBatchOptimizerTopKStrategy(min_distance=10.0).select(x, a, 4)
    assert fs.tolist() == [-5.0]                        # fewer than asked for when the distance rule leaves no more


def test_direct_optimizer_maps_scipydirect_arguments_and_defaults():
    """scipydirect's defaults (original DIRECT, maxf 20000, maxT 6000) and its keyword names."""
    import scipy.optimize
    from bopy_b200.optimizer import DirectOptimizer

    class Quadratic:
        def __call__(self, x):
            return np.array([float(((x - 0.3) ** 2).sum())])

    seen = {}
    real = scipy.optimize.direct

    def spy(func, bounds, **kwargs):
        seen.update(kwargs)
        return real(func, bounds, **kwargs)

    scipy.optimize.direct = spy
    try:
        res = DirectOptimizer(Quadratic(), Bounds([Bound(0.0, 1.0)]), maxf=80, fglobal=0.0, fglper=1.0, volper=-1.0,
                              disp=True, logfilename="x.log").optimize()
    finally:
        scipy.optimize.direct = real
    assert seen["maxfun"] == 80 and seen["maxiter"] == 6000 and seen["locally_biased"] is False
    assert seen["f_min"] == 0.0 and seen["f_min_rtol"] == pytest.approx(0.01) and "vol_tol" not in seen
    assert "disp" not in seen and "logfilename" not in seen
    assert res.x_min.shape == (1, 1) and res.f_min.shape == (1,) and abs(res.x_min[0, 0] - 0.3) < 0.05


def test_latin_hypercube_follows_numpy_global_seed_and_one_shot_rejects_sharded_base():
    from bopy_b200.initial_design import LatinHypercubeInitialDesign
    from bopy_b200.optimizer import CandidateSweepOptimizer, OneShotBatchOptimizer, OneShotBatchOptimizerRandomSamplingStrategy
    b = Bounds([Bound(0.0, 1.0), Bound(-2.0, 2.0)])
    np.random.seed(5)
    first = LatinHypercubeInitialDesign().generate(b, 7)
    np.random.seed(5)
    assert np.array_equal(first, LatinHypercubeInitialDesign().generate(b, 7))

    class Acq:
        surrogate = None

        def start_optimization(self):
            pass

    base = CandidateSweepOptimizer(Acq(), b, n_candidates=128, distributed=True)
    with pytest.raises(ValueError, match="whole evaluation log"):
        OneShotBatchOptimizer(Acq(), b, base_optimizer=base, batch_size=2,
                              strategy=OneShotBatchOptimizerRandomSamplingStrategy()).optimize()


def test_matern_inf_is_flattened_to_rbf_and_unsupported_kernels_say_so():
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, RationalQuadratic
    from bopy_b200.kernel_spec import UnsupportedKernelError, flatten_sklearn_kernel
    flat = flatten_sklearn_kernel(ConstantKernel(2.0) * Matern([0.5, 0.25], nu=np.inf))
    assert flat.kernel == "rbf" and flat.amplitude == 2.0 and np.array_equal(flat.length_scale, [0.5, 0.25])
    with pytest.raises(UnsupportedKernelError, match="no CPU fallback"):
        flatten_sklearn_kernel(RationalQuadratic())


def test_direct_optimizer_asks_for_the_inverse_path_only_with_a_large_budget():
    """DirectOptimizer knows how often it will probe one fitted state (bopy/optimizer.py:95-107): from maxf = 64 on it switches
    the surrogate's handle to W = L^-1 at the first probe and hands the mode back afterwards -- unless the user fixed the mode."""
    from bopy_b200.optimizer import DirectOptimizer

    class Handle:
        n = 2048

        def __init__(self):
            self.calls = []

        def set_inverse_path(self, mode):
            self.calls.append(mode)
            return 8

    class Sur:
        def __init__(self, mode):
            self.native, self.inverse_path = Handle(), mode

    class Acq:
        def __init__(self, sur):
            self.surrogate = sur

        def __call__(self, x):
            return np.array([float(((x - 0.3) ** 2).sum())])

    b = Bounds([Bound(0.0, 1.0)])
    for mode, maxf, want in (("auto", 200, [1, -1]), ("auto", 63, []), (False, 200, []), (True, 200, [])):
        acq = Acq(Sur(mode))
        DirectOptimizer(acq, b, maxf=maxf).optimize()
        assert acq.surrogate.native.calls == want, (mode, maxf)

    class Failing(Acq):
        def __call__(self, x):
            raise RuntimeError("objective failed")
    acq = Failing(Sur("auto"))
    with pytest.raises(RuntimeError):
        DirectOptimizer(acq, b, maxf=200).optimize()
    assert acq.surrogate.native.calls == [1, -1]        # the mode is handed back even when the run dies


def test_surrogate_rejects_an_unknown_inverse_path_setting():
    from bopy_b200.surrogate import B200GPSurrogate
    with pytest.raises(ValueError, match="inverse_path must be"):
        B200GPSurrogate(object(), inverse_path="sometimes")
    for ok in ("auto", True, False):
        assert B200GPSurrogate(object(), inverse_path=ok).inverse_path is ok or ok == "auto"
