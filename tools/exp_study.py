"""Accuracy of the branch-free exp used by the kernel-tile step (bopy_b200/csrc/sweep_kernel.cuh: exp_nonpos),
restated operation by operation in numpy: <= 1.01 ulp against np.exp on [-708, 0].
    python tools/exp_study.py"""
import math

import numpy as np


def main():
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.random(2_000_000) * 50, -rng.random(500_000) * 708, -np.logspace(-18, 2.8, 200000)])
    x = x[x >= -708.0]
    magic = 6755399441055744.0
    t = x * 1.4426950408889634 + magic
    kd = t - magic
    r = (x - kd * 6.93147180369123816490e-01) - kd * 1.90821492927058770002e-10
    p = np.full_like(r, 1.0 / math.factorial(13))
    for n in range(12, -1, -1):
        p = p * r + 1.0 / math.factorial(n)
    res = np.ldexp(p, kd.astype(np.int64))
    ref = np.exp(x)
    rel = np.abs(res - ref) / ref
    print(f"exp_nonpos: max rel err {rel.max():.3e} = {rel.max() / 2.220446049250313e-16:.2f} ulp, max |r| {np.abs(r).max():.4f}")


if __name__ == "__main__":
    main()
