"""Time of the device fit (bopy_gp_fit: Gram matrix, blocked Cholesky, alpha_, packing) with the panel / trailing-update
tiles on gemm_nt_async_kernel (cp.async ring, default) and on the register-staged gemm_nt_kernel
(BOPY_B200_CHOL_GEMM=registers).  Host clock around a synchronised call, best of 9."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bopy_b200 import _native  # noqa: E402


def best(f, reps=9):
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
    for n, d, ls in ((256, 2, 0.3), (1000, 3, 0.3), (2048, 6, 0.3), (4096, 10, 0.6), (8192, 20, 1.0)):
        X = rng.random((n, d))
        y = np.sin(X.sum(1))
        yn = (y - y.mean()) / y.std()
        gp = _native.NativeGP(n, d, kernel="rbf", dtype="f64")
        Xd, yd = gp._dev64(X, (n, d)), gp._dev64(yn, (n,))
        row, factor = {}, {}
        for variant in ("cp_async", "registers"):
            os.environ["BOPY_B200_CHOL_GEMM"] = variant
            _, L = gp.fit(Xd, yd, [ls], alpha_reg=1e-6, want_factor=True)
            factor[variant] = L.cpu().numpy()
            row[variant] = best(lambda: gp.fit(Xd, yd, [ls], alpha_reg=1e-6))
        same = np.array_equal(np.tril(factor["cp_async"]), np.tril(factor["registers"]))
        print(f"n={n:5d} d={d:2d}: fit {row['cp_async']:8.3f} ms (register-staged tiles: {row['registers']:8.3f})   "
              f"factors bit-identical: {same}", flush=True)
        gp.close()


if __name__ == "__main__":
    main()
