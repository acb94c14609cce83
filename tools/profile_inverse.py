"""A few launches of the inverse path (W = L^-1 by tile_gemm_async_kernel, then probe_inv_kernel) for ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_inverse.csv \
        python tools/profile_inverse.py [n d]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bopy_b200.surrogate import B200GPSurrogate  # noqa: E402


def main():
    import torch
    n, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 6)
    X, y, gp = bench.make_problem(n, d)
    sur = B200GPSurrogate(gp, inverse_path=True)
    sur.fit(X, y)
    native, eta = sur.native, float(y.min())
    rng = np.random.default_rng(0)
    served = native.set_inverse_path(1)
    for rep in range(3):                      # the first launch builds W; the later ones are warm
        for m in (1, served):
            xs = native.candidates(rng.random((m, d)))
            out = native.sweep(xs, acq="ei", eta=eta, want_acq=True, want_min=True)
    torch.cuda.synchronize()
    print("served", served, "acq", out["acq"].cpu().numpy()[:2])


if __name__ == "__main__":
    main()
