"""BASELINE config C1: 1-D Bayesian optimisation of the Forrester function with LCB (kappa = 2), 5 Sobol points +
5 trials -- the setting of the reference's examples/example_1d.py, with the posterior and the acquisition sweep on
the B200.  The reference drives a GPy RBF model with DIRECT (maxf=100); GPy and scipydirect are not installed here,
so the surrogate is scikit-learn's GP with the same kernel family (variance * RBF, noise 1e-10, normalised targets,
hyper-parameters optimised) and the optimiser is the fused candidate sweep + zoom.

    python examples/example_1d.py
"""
import os
import sys

import numpy as np
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from bopy_b200.acquisition import LCB  # noqa: E402
from bopy_b200.bayes_opt import BayesOpt  # noqa: E402
from bopy_b200.benchmark_functions import forrester  # noqa: E402
from bopy_b200.bounds import Bound, Bounds  # noqa: E402
from bopy_b200.callback import Callback  # noqa: E402
from bopy_b200.initial_design import SobolSequenceInitialDesign  # noqa: E402
from bopy_b200.optimizer import CandidateSweepOptimizer  # noqa: E402
from bopy_b200.surrogate import ScipyGPSurrogate  # noqa: E402


class ProgressCallback(Callback):
    """Text stand-in for the reference example's plotting callback: one line per event of interest."""

    def __init__(self, out=print):
        self.out, self.trial = out, 0

    def on_acquisition_optimized(self, bo, opt_result):
        grid = np.linspace(bo.bounds.lowers[0], bo.bounds.uppers[0], 5).reshape(-1, 1)
        mean, var = bo.surrogate.predict_diag(grid)
        self.out(f"trial {self.trial}: propose x = {opt_result.x_min[0, 0]:.5f}  LCB = {opt_result.f_min[0]:.4f}   "
                 f"posterior mean on a 5-point grid {np.round(mean, 3)}  std {np.round(np.sqrt(np.abs(var)), 3)}")

    def on_trial_end(self, bo):
        x_best, f_best = bo.get_opt_so_far()
        self.out(f"trial {self.trial}: best so far f({x_best[0, 0]:.5f}) = {f_best:.5f}  ({len(bo.y)} evaluations)")
        self.trial += 1


def main(n_trials=5, n_initial_design=5, n_candidates=1 << 16, out=print):
    bounds = Bounds(bounds=[Bound(lower=0.0, upper=1.0)])
    gp = GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(0.2), alpha=1e-10, normalize_y=True)
    surrogate = ScipyGPSurrogate(gp=gp)
    acquisition = LCB(surrogate=surrogate)                      # kappa = 2.0, as in the reference example
    optimizer = CandidateSweepOptimizer(acquisition, bounds, n_candidates=n_candidates, zoom_rounds=3, seed=1)
    bo = BayesOpt(objective_function=forrester, surrogate=surrogate, acquisition_function=acquisition,
                  optimizer=optimizer, initial_design=SobolSequenceInitialDesign(), bounds=bounds,
                  callbacks=[ProgressCallback(out)])
    result = bo.run(n_trials=n_trials, n_initial_design=n_initial_design)
    out(f"optimum found: f({result.x_opt[0, 0]:.5f}) = {result.f_opt:.5f}   (true minimum f(0.75725) = -6.02074)")
    return result


if __name__ == "__main__":
    main()
