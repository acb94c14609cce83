"""One pass through the non-sweep kernels (device fit, LML + gradient, full covariance) for ncu:
    ncu --set full -k regex:"gemm_nt|tile_gemm|chol_block|solve_alpha|gram_kernel|lml_|cov_kernel" python tools/profile_aux.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bopy_b200.surrogate import B200GPSurrogate  # noqa: E402


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    X, y, gp = bench.make_problem(n, 6)
    sur = B200GPSurrogate(gp)
    sur.fit(X, y)                                   # gram, chol_block, gemm_nt, solve_alpha, pack_*
    lml, grad = sur.native.lml(X, sur.gp.y_train_, 0.3 * np.ones(6), alpha_reg=1e-6, want_grad=True)   # + tile_gemm, lml_*
    xs = np.random.default_rng(0).random((256, 6))
    mean, cov = sur.predict(xs)                     # sweep + cov_kernel
    torch.cuda.synchronize()
    print("lml", lml, "grad", grad[:3], "cov", cov[0, 0])


if __name__ == "__main__":
    main()
