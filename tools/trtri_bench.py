"""Time of W = L^-1 (the inverse path's one-time cost per fitted state) and of bopy_gp_lml with its gradient, for the
recursive 2 x 2 block inversion on tile_gemm_async_kernel (default) and for round-2's first version, one block diagonal
at a time on tile_gemm_kernel (BOPY_B200_TRTRI=diagonal).  Build time = (device fit + first probe on the inverse path)
- (device fit + first probe on the chained path), host clock around a synchronised call, best of 7."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bopy_b200 import _native  # noqa: E402


def best(f, reps=7):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    return min(ts)


def main():
    rng = np.random.default_rng(0)
    for n, d, ls in ((256, 2, 0.3), (700, 3, 0.3), (2048, 6, 0.3), (4096, 10, 0.6), (8192, 20, 1.0)):
        X = rng.random((n, d))
        y = np.sin(X.sum(1))
        yn = (y - y.mean()) / y.std()
        gp = _native.NativeGP(n, d, kernel="rbf", dtype="f64")
        gp.set_latency_path(1 << 30)
        x1 = gp.candidates(rng.random((1, d)))
        Xd, yd = gp._dev64(X, (n, d)), gp._dev64(yn, (n,))

        def fit_and_probe():
            gp.fit(Xd, yd, [ls], alpha_reg=1e-6)
            gp.sweep(x1, acq="ei", eta=0.0, want_acq=True)
        row = {}
        for variant in ("recursive", "diagonal"):
            os.environ["BOPY_B200_TRTRI"] = variant
            gp.set_inverse_path(0)
            fit_and_probe()
            base = best(fit_and_probe)
            gp.set_inverse_path(1)
            fit_and_probe()
            row[variant] = best(fit_and_probe) - base
            row[variant + "_lml"] = best(lambda: gp.lml(Xd, yd, [ls], alpha_reg=1e-6, want_grad=True), 5)
        row["lml_value_only"] = best(lambda: gp.lml(Xd, yd, [ls], alpha_reg=1e-6, want_grad=False), 5)
        print(f"n={n:5d} d={d:2d}: W = L^-1 build {row['recursive']:8.3f} ms (one block diagonal at a time: {row['diagonal']:8.3f})   "
              f"lml + gradient {row['recursive_lml']:8.3f} ms ({row['diagonal_lml']:8.3f}; value only {row['lml_value_only']:.3f})", flush=True)
        gp.close()


if __name__ == "__main__":
    main()
