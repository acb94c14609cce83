"""Small end-to-end exercise of every kernel for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
Ragged sizes on purpose (n = 333 -> 3 block rows with padding, m = 517 -> 5 tiles with a tail)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sklearn.gaussian_process import GaussianProcessRegressor  # noqa: E402
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern  # noqa: E402

from bopy_b200 import _native  # noqa: E402
from bopy_b200.acquisition import EI, LCB  # noqa: E402
from bopy_b200.surrogate import B200GPSurrogate  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    X = rng.random((333, 4))
    y = np.sin(3 * X.sum(1))
    xs = rng.random((517, 4))
    for dtype in ("f64", "f32"):
        for kernel, device_fit in ((ConstantKernel(1.3) * RBF([0.4, 0.5, 0.6, 0.7]), True), (Matern(0.5, nu=2.5), False)):
            sur = B200GPSurrogate(GaussianProcessRegressor(kernel=kernel, alpha=1e-6, normalize_y=True, optimizer=None),
                                  dtype=dtype, device_fit=device_fit)
            sur.fit(X, y)
            mean, var = sur.predict_diag(xs)
            m2, cov = sur.predict(xs[:130])
            ei = EI(sur)
            ei.fit(X, y)
            a = ei(xs)
            idx, val = ei.argmin(xs)
            assert idx == int(np.argmin(a)), (idx, int(np.argmin(a)))
            vals, idxs = sur.acquisition_segment_argmin("lcb", xs, 128, kappa=2.0)
            cloud = _native.candidates_around(1, _native.gather_rows(sur.native.candidates(xs), idxs), 128,
                                              [0.1] * 4, [0.0] * 4, [1.0] * 4)
            LCB(sur).fit(X, y)
            print(dtype, type(kernel).__name__, "ok", float(mean[0]), float(var[0]), float(cov[0, 0]), idx, val,
                  tuple(cloud.shape))
    print("peaks", _native.measure_peak("fp64_mma"))


if __name__ == "__main__":
    main()
