"""A few launches of the latency-path kernels (probe_kernel, grad_kernel, multistart_step_kernel) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_latency.csv \
        python tools/profile_latency.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bopy_b200 import _native  # noqa: E402
from bopy_b200.surrogate import B200GPSurrogate  # noqa: E402


def main():
    import torch
    n, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 6)
    X, y, gp = bench.make_problem(n, d)
    sur = B200GPSurrogate(gp)
    sur.fit(X, y)
    native, eta = sur.native, float(y.min())
    rng = np.random.default_rng(0)
    for rep in range(2):                      # the second pass is the warm one
        for m in (1, 64, 1024, 4096):
            xs = native.candidates(rng.random((m, d)))
            native.sweep(xs, acq="ei", eta=eta, want_acq=True, want_min=True)        # probe_kernel + minloc_finalize
        for m in (1, 1024):
            xs = native.candidates(rng.random((m, d)))
            val, grad, _, _ = native.value_and_grad(xs, "ei", eta=eta)               # probe_kernel + grad_kernel
        xc, gc = torch.empty_like(xs), torch.empty_like(xs)
        fc, alpha = torch.empty_like(val), torch.ones_like(val)
        _native.multistart_step(np.zeros(d), np.ones(d), xc, fc, gc, xs, val, grad, alpha, first=True)
    torch.cuda.synchronize()
    print("ok", float(val[0]), float(grad[0, 0]))


if __name__ == "__main__":
    main()
