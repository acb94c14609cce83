"""BASELINE config C4 in miniature: Bayesian optimisation of Hartmann-6 (d = 6) with a fixed-hyper-parameter GP on
the B200 path -- the pieces the reference's loop calls every trial (bopy/bayes_opt.py:203-247), at sizes the reference
cannot reach:

  * acquisition optimisation = a fused sweep of 2^18 candidates with branch and bound (same winner as the plain sweep),
    or the gradient-based multi-start;
  * surrogate update = a one-row append of the Cholesky factor (the data grow by one point per trial);
  * a Kriging-believer batch of 4 (bopy/acquisition.py:172-197): four sweeps, four one-row appends, one truncation.

    python examples/example_hartmann6.py [n_initial] [n_trials]
"""
import os
import sys
import time

import numpy as np
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from bopy_b200.acquisition import EI, KriggingBeliever  # noqa: E402
from bopy_b200.bayes_opt import BayesOpt  # noqa: E402
from bopy_b200.benchmark_functions import hartmann6  # noqa: E402
from bopy_b200.bounds import unit_box  # noqa: E402
from bopy_b200.callback import Callback  # noqa: E402
from bopy_b200.initial_design import UniformRandomInitialDesign  # noqa: E402
from bopy_b200.optimizer import CandidateSweepOptimizer, MultiStartOptimizer, SequentialBatchOptimizer  # noqa: E402
from bopy_b200.surrogate import B200GPSurrogate  # noqa: E402


class Timing(Callback):
    def __init__(self):
        self.t = time.perf_counter()
        self.rows = []

    def _lap(self):
        now = time.perf_counter()
        dt, self.t = now - self.t, now
        return 1e3 * dt

    def on_acquisition_optimized(self, bo, opt_result):
        self.opt_ms = self._lap()

    def on_surrogate_updated(self, bo):
        self.fit_ms = self._lap()

    def on_trial_end(self, bo):
        self.rows.append((len(bo.y), float(np.min(bo.y)), self.opt_ms, self.fit_ms))
        self.t = time.perf_counter()


def main(n_initial=512, n_trials=8, out=print):
    np.random.seed(0)
    bounds = unit_box(6)
    gp = GaussianProcessRegressor(ConstantKernel(1.0) * RBF(0.3 * np.ones(6)), alpha=1e-6, normalize_y=True, optimizer=None)
    surrogate = B200GPSurrogate(gp)
    acquisition = EI(surrogate)
    timing = Timing()
    sweep = CandidateSweepOptimizer(acquisition, bounds, n_candidates=1 << 18, prune=True, seed=1)
    bo = BayesOpt(hartmann6, surrogate, acquisition, sweep, UniformRandomInitialDesign(), bounds, callbacks=[timing])
    result = bo.run(n_trials=n_trials, n_initial_design=n_initial)
    for n, best, opt_ms, fit_ms in timing.rows:
        out(f"n = {n:5d}  best f = {best:8.4f}   acquisition sweep (2^18, pruned) {opt_ms:7.2f} ms   surrogate update {fit_ms:6.2f} ms")
    out(f"incumbent after {n_trials} trials: f = {result.f_opt:.4f} at {np.round(result.x_opt[0], 3)}  "
        f"(global minimum -3.3224); one-row appends: {getattr(surrogate, 'appended_rows', 0)}")

    # a Kriging-believer batch of 4 on the same model
    believer = KriggingBeliever(EI(surrogate))
    believer.fit(bo.x, bo.y)
    base = CandidateSweepOptimizer(believer, bounds, n_candidates=1 << 18, prune=True, seed=2)
    t0 = time.perf_counter()
    batch = SequentialBatchOptimizer(believer, bounds, base_optimizer=base, batch_size=4).optimize()
    out(f"Kriging-believer batch of 4 in {1e3 * (time.perf_counter() - t0):.1f} ms: EI values {np.round(batch.f_min, 5)}; "
        f"appends so far {surrogate.appended_rows}, truncations {getattr(surrogate, 'truncations', 0)}")

    # gradient-based multi-start on the final model
    acquisition.fit(bo.x, bo.y)
    t0 = time.perf_counter()
    ms = MultiStartOptimizer(acquisition, bounds, n_starts=256, n_candidates=1 << 17, seed=3).optimize()
    out(f"multi-start (256 starts, gradient refinement) in {1e3 * (time.perf_counter() - t0):.1f} ms: EI = {ms.f_min[0]:.6f}")
    return result, batch, ms


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:3]))
