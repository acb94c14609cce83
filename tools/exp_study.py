"""Accuracy of the branch-free exp used by the kernel-tile step (bopy_b200/csrc/sweep_kernel.cuh: exp_nonpos),
restated operation by operation in numpy (fused multiply-adds emulated in extended precision): <= 1 ulp against the
exponential on [-708, 0], exact at 0.
    python tools/exp_study.py"""
import numpy as np


def fma(a, b, c):
    return np.float64(np.asarray(a, dtype=np.longdouble) * np.asarray(b, dtype=np.longdouble) + np.asarray(c, dtype=np.longdouble))


def exp_nonpos(x):
    table = np.exp2(np.arange(64) / 64.0)            # 2^(j/64); the kernel holds the correctly rounded values
    magic = 6755399441055744.0
    xc = np.maximum(x, -708.0)
    t = fma(xc, 92.33248261689366, magic)
    nd = t - magic
    n = nd.astype(np.int64)
    r = fma(nd, -0.01083042469326756, xc)
    r = fma(nd, -2.9815858269852933e-12, r)
    r2 = r * r
    a1 = fma(1.6666666666666666e-01, r, 0.5)
    a2 = fma(8.333333333333333e-03, r, 4.1666666666666664e-02)
    q = fma(fma(a2, r2, a1), r2, r)
    T = table[n & 63]
    res = np.ldexp(fma(T, q, T), n >> 6)
    return np.where(x < -708.0, 0.0, res), r


def main():
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.random(2_000_000) * 50, -rng.random(500_000) * 708, -np.logspace(-18, 2.8, 200000), np.zeros(1)])
    x = x[x >= -708.0]
    res, r = exp_nonpos(x)
    ref = np.exp(x)
    rel = np.abs(res - ref) / ref
    print(f"exp_nonpos: max rel err against np.exp {rel.max():.3e} = {rel.max() / 2.220446049250313e-16:.2f} ulp, "
          f"max |r| {np.abs(r).max():.5f}, exp(0) == 1: {bool(res[-1] == 1.0)}")
    try:
        import mpmath
        mpmath.mp.prec = 200
        worst = 0.0
        for i in np.argsort(rel)[-2000:]:
            true = mpmath.exp(mpmath.mpf(float(x[i])))
            worst = max(worst, float(abs((mpmath.mpf(float(res[i])) - true) / true)))
        print(f"worst true error among the 2000 largest: {worst / 2.220446049250313e-16:.2f} ulp")
    except ImportError:
        pass


if __name__ == "__main__":
    main()
