"""Batched multi-start acquisition minimisation (BASELINE config C5: 1024 starts): derivative-free clouds against
the gradient method, wall time and value reached.  Prints one JSON line per run."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bopy_b200.acquisition import EI, LCB  # noqa: E402
from bopy_b200.bounds import Bound, Bounds  # noqa: E402
from bopy_b200.optimizer import MultiStartOptimizer  # noqa: E402
from bopy_b200.surrogate import B200GPSurrogate  # noqa: E402


def main():
    import torch
    shapes = [(2048, 6, 1 << 20), (8192, 20, 1 << 18)] if "--large" in sys.argv else [(2048, 6, 1 << 20)]
    for n, d, ncand in shapes:
        X, y, gp = bench.make_problem(n, d)
        sur = B200GPSurrogate(gp)
        sur.fit(X, y)
        for name, acq in (("lcb", LCB(sur, kappa=2.0)), ("ei", EI(sur))):
            acq.fit(X, y)
            bounds = Bounds([Bound(0.0, 1.0)] * d)
            for method, kw in (("cloud", dict(rounds=6)), ("gradient", dict(iterations=40)),
                               ("gradient", dict(iterations=40, prune=True))):
                opt = MultiStartOptimizer(acq, bounds, n_starts=1024, n_candidates=ncand, seed=1, method=method, **kw)
                opt.optimize()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res = opt.optimize()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                xs, vs = opt.local_minima()
                print(json.dumps(dict(n=n, d=d, acq=name, method=method, n_starts=1024, global_candidates=opt.n_candidates,
                                      seconds=dt, f_min=float(res.f_min[0]), median_local=float(np.nanmedian(vs)),
                                      **kw)), flush=True)


if __name__ == "__main__":
    main()
