"""BASELINE config C2: batch 1-D Bayesian optimisation -- q = 2 points per iteration chosen sequentially with the
Kriging-believer strategy around LCB, 3 trials: the setting of the reference's examples/example_batch_1d.py.
Every member of a batch costs one device refit (fixed hyper-parameters: Gram + Cholesky on the GPU) and one fused
sweep; the fantasies are removed again by `finish_batch`.

    python examples/example_batch_1d.py
"""
import os
import sys

import numpy as np
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from bopy_b200.acquisition import LCB, KriggingBeliever  # noqa: E402
from bopy_b200.bayes_opt import BayesOpt  # noqa: E402
from bopy_b200.benchmark_functions import forrester  # noqa: E402
from bopy_b200.bounds import Bound, Bounds  # noqa: E402
from bopy_b200.callback import Callback  # noqa: E402
from bopy_b200.initial_design import SobolSequenceInitialDesign  # noqa: E402
from bopy_b200.optimizer import CandidateSweepOptimizer, SequentialBatchOptimizer  # noqa: E402
from bopy_b200.surrogate import ScipyGPSurrogate  # noqa: E402


class BatchLog(Callback):
    def __init__(self, out=print):
        self.out = out

    def on_acquisition_optimized(self, bo, opt_result):
        pts = ", ".join(f"{v:.4f}" for v in opt_result.x_min[:, 0])
        self.out(f"batch proposed: x = [{pts}]  LCB = {np.round(opt_result.f_min, 4)}")

    def on_trial_end(self, bo):
        x_best, f_best = bo.get_opt_so_far()
        self.out(f"  after {len(bo.y)} evaluations: best f({x_best[0, 0]:.5f}) = {f_best:.5f}")


def main(n_trials=3, n_initial_design=5, batch_size=2, out=print):
    bounds = Bounds(bounds=[Bound(lower=0.0, upper=1.0)])
    gp = GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(0.15), alpha=1e-8, normalize_y=True, optimizer=None)
    surrogate = ScipyGPSurrogate(gp=gp)
    acquisition = KriggingBeliever(base_acquisition=LCB(surrogate=surrogate))
    base = CandidateSweepOptimizer(acquisition, bounds, n_candidates=1 << 15, zoom_rounds=2, seed=2)
    optimizer = SequentialBatchOptimizer(acquisition, bounds, base_optimizer=base, batch_size=batch_size)
    bo = BayesOpt(objective_function=forrester, surrogate=surrogate, acquisition_function=acquisition,
                  optimizer=optimizer, initial_design=SobolSequenceInitialDesign(), bounds=bounds,
                  callbacks=[BatchLog(out)])
    result = bo.run(n_trials=n_trials, n_initial_design=n_initial_design)
    out(f"optimum found: f({result.x_opt[0, 0]:.5f}) = {result.f_opt:.5f}")
    return result


if __name__ == "__main__":
    main()
