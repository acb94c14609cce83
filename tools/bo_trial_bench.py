"""Wall time of the pieces of ONE BayesOpt trial (bopy/bayes_opt.py:203-230: optimise the acquisition, evaluate, refit
the surrogate, refit the acquisition) on the B200 path, through the public API, fixed hyper-parameters."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bopy_b200.acquisition import EI  # noqa: E402
from bopy_b200.bounds import Bound, Bounds  # noqa: E402
from bopy_b200.optimizer import CandidateSweepOptimizer, DirectOptimizer, MultiStartOptimizer  # noqa: E402
from bopy_b200.surrogate import B200GPSurrogate  # noqa: E402


def torch_sync():
    import torch
    torch.cuda.synchronize()


def wall(fn, reps=5):
    import torch
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps, out


def main():
    for n, d in ((2048, 6), (8192, 20)) if "--large" in sys.argv else ((2048, 6),):
        X, y, gp = bench.make_problem(n, d)
        sur = B200GPSurrogate(gp)
        t_fit, _ = wall(lambda: sur.fit(X, y))
        # the per-trial refit: one more observation, same hyper-parameters -> one-row extension of the factor
        Xg, yg = np.vstack([X, np.random.default_rng(1).random((8, d))]), np.concatenate([y, y[:8]])
        t_app = []
        for k in range(1, 7):
            t0 = time.perf_counter()
            sur.fit(Xg[:n + k], yg[:n + k])
            torch_sync()
            t_app.append(1e3 * (time.perf_counter() - t0))
        sur.fit(X, y)
        ei = EI(sur)
        ei.fit(X, y)
        bounds = Bounds([Bound(0.0, 1.0)] * d)
        row = dict(n=n, d=d, surrogate_fit_ms=t_fit, surrogate_refit_one_more_point_ms=float(np.median(t_app[1:])),
                   appended_rows=int(getattr(sur, "appended_rows", 0)))
        ncand = 1 << 20 if n <= 2048 else 1 << 18
        row["sweep_ms"], r0 = wall(lambda: CandidateSweepOptimizer(ei, bounds, n_candidates=ncand, seed=1).optimize(), 3)
        row["sweep_pruned_ms"], r1 = wall(lambda: CandidateSweepOptimizer(ei, bounds, n_candidates=ncand, seed=1, prune=True).optimize(), 3)
        row["same_winner"] = bool(np.array_equal(r0.x_min, r1.x_min) and np.array_equal(r0.f_min, r1.f_min))
        row["multistart_gradient_ms"], r2 = wall(lambda: MultiStartOptimizer(ei, bounds, n_starts=256, n_candidates=ncand, seed=1).optimize(), 3)
        row["direct_maxf100_ms"], r3 = wall(lambda: DirectOptimizer(ei, bounds, maxf=100).optimize(), 2)
        # DIRECT at a realistic budget (the reference's default is maxf = 20000): W = L^-1 is built at the first probe
        row["direct_maxf2000_ms"], r4 = wall(lambda: DirectOptimizer(ei, bounds, maxf=2000).optimize(), 2)
        sur.native.set_inverse_path(0)
        sur.inverse_path = False
        row["direct_maxf2000_chained_path_ms"], r5 = wall(lambda: DirectOptimizer(ei, bounds, maxf=2000).optimize(), 2)
        sur.inverse_path = "auto"
        sur.native.set_inverse_path(-1)
        row["direct_same_optimum"] = bool(np.allclose(r4.x_min, r5.x_min, atol=1e-6))
        row["candidates"] = ncand
        row["f_min"] = dict(sweep=float(r0.f_min[0]), multistart=float(r2.f_min[0]), direct=float(r3.f_min[0]))
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
