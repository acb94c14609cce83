"""Accuracy of the branch-free exp used by the kernel-tile step (bopy_b200/csrc/sweep_kernel.cuh: exp_nonpos),
restated operation by operation in numpy: <= 2 ulp against np.exp on [-708, 0].
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
    c = [1.0 / math.factorial(n) for n in range(14)]
    a = [c[2 * i] + c[2 * i + 1] * r for i in range(7)]          # Estrin's scheme, as in the kernel
    r2 = r * r
    r4 = r2 * r2
    r8 = r4 * r4
    b0, b1, b2 = a[0] + a[1] * r2, a[2] + a[3] * r2, a[4] + a[5] * r2
    p = (b0 + b1 * r4) + (b2 + a[6] * r4) * r8
    res = np.ldexp(p, kd.astype(np.int64))
    ref = np.exp(x)
    rel = np.abs(res - ref) / ref
    print(f"exp_nonpos: max rel err {rel.max():.3e} = {rel.max() / 2.220446049250313e-16:.2f} ulp, max |r| {np.abs(r).max():.4f}")


if __name__ == "__main__":
    main()
