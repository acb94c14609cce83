"""Arg-min over 2^21 candidates at the headline shape (n=2048, d=6): plain fused sweep vs branch and bound."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bopy_b200 import _native  # noqa: E402
from bopy_b200.surrogate import B200GPSurrogate  # noqa: E402


def timed(fn, reps):
    import torch
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    for n, d, m in ((2048, 6, 1 << 21), (8192, 20, 1 << 19)) if "--large" in sys.argv else ((2048, 6, 1 << 21),):
        X, y, gp = bench.make_problem(n, d)
        sur = B200GPSurrogate(gp)
        sur.fit(X, y)
        eta = float(y.min())
        xs = _native.candidates_uniform(1235, 0, m, np.zeros(d), np.ones(d))
        for acq in ("ei", "lcb", "poi"):
            t_full, full = timed(lambda: sur.native.sweep(xs, acq=acq, eta=eta, kappa=2.0, want_min=True), 3)
            t_pr, (minv, mini, stats) = timed(lambda: sur.native.argmin_pruned(xs, acq, eta=eta, kappa=2.0), 3)
            same = int(mini.item()) == int(full["min_idx"].item()) and float(minv.item()) == float(full["min_val"].item())
            print(json.dumps(dict(n=n, d=d, m=m, acq=acq, plain_ms=t_full, pruned_ms=t_pr, speedup=t_full / t_pr,
                                  swept=stats["swept"], same_argmin=same)), flush=True)


if __name__ == "__main__":
    main()
