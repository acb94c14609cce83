"""Per-call latency of the public API for the reference's own calling pattern (few candidates per call:
DIRECT probes with m = 1, bopy/optimizer.py:96-97; plotting grids with m = 100..2500), beside the reference path."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bopy_b200.acquisition import EI  # noqa: E402
from bopy_b200.surrogate import B200GPSurrogate  # noqa: E402
from oracle import reference_path as R  # noqa: E402


def main():
    import torch
    for n, d in ((256, 2), (2048, 6)):
        X, y, gp = bench.make_problem(n, d)
        sur = B200GPSurrogate(gp)
        sur.fit(X, y)
        ei = EI(sur)
        ei.fit(X, y)
        host = bench.make_problem(n, d)[2].fit(X, y)
        eta = float(y.min())
        rng = np.random.default_rng(0)
        for m in (1, 8, 64, 128, 1024):
            xs = rng.random((m, d))
            ei(xs)
            torch.cuda.synchronize()
            reps = 20
            t0 = time.perf_counter()
            for _ in range(reps):
                ei(xs)
            t_gpu = (time.perf_counter() - t0) / reps
            with bench.all_host_threads():
                R.ei(host, xs[:min(m, 256)], eta)
                t0 = time.perf_counter()
                for _ in range(3):
                    for s in range(0, m, 64):
                        R.ei(host, xs[s:s + 64], eta)
                t_ref = (time.perf_counter() - t0) / 3
            print(f"n={n:5d} d={d} m={m:5d}: EI(x) through the API {1e3 * t_gpu:8.3f} ms/call   reference path {1e3 * t_ref:8.3f} ms/call")


if __name__ == "__main__":
    main()
