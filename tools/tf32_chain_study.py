"""Design study (CPU, numpy) for the tcgen05 fp32 engine: how long may an fp32 accumulation chain in TMEM be?

The blocked forward substitution R_I = K*_I - sum_J L_IJ V_J with L and V stored as TF32 pairs (hi = top 19 bits of
the fp32 value, lo = fp32(x - hi), read by the tensor core as its top 19 bits) and the three products
hi.hi + hi.lo + lo.hi accumulated in fp32, one rounding per k=8 MMA step.  Two accumulator models:
  rn   round to nearest after every MMA step (what an IEEE FADD chain would do)
  rz   truncation towards zero after every MMA step (what round 1 measured on mma.sync TF32: a biased drift)
`chain` = number of 128-column blocks of L accumulated in one fp32 chain before it is folded into the fp64 residual.
Run: python tools/tf32_chain_study.py [n] [d] [ls]
"""
import sys

import numpy as np
from scipy.linalg import solve_triangular
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
d = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ls = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
m, B = 128, 128
rng = np.random.default_rng(1234)
X = rng.random((n, d))
y = np.sin(X.sum(1) * 3)
gp = GaussianProcessRegressor(kernel=C(1.0) * RBF(ls * np.ones(d)), alpha=1e-6, optimizer=None, normalize_y=True)
gp.fit(X, y)
L = gp.L_
Xs = np.random.default_rng(1235).random((m, d))
Kt = gp.kernel_(Xs, gp.X_train_).T          # (n, m)
V = solve_triangular(L, Kt, lower=True, check_finite=False)
var_ref = 1.0 - np.einsum("ij,ij->j", V, V)
print(f"n={n} d={d} ls={ls} var quantiles {np.quantile(var_ref, [0, .1, .5, 1])}")


def tf32_trunc(x32):
    return (x32.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)


def tf32_rn(x32):
    """cvt.rna.tf32.f32: round to nearest (ties away) on the 13 dropped bits."""
    u = x32.view(np.uint32).astype(np.uint64) + np.uint64(0x1000)
    return (u.astype(np.uint32) & np.uint32(0xffffe000)).view(np.float32)


SPLIT = "rn"   # how the stored pair is formed: "trunc" (round 1's in-register split) or "rn" (pre-split at packing time)


def split(x):
    x32 = x.astype(np.float32)
    if SPLIT == "trunc":
        hi = tf32_trunc(x32)
        lo = tf32_trunc((x32 - hi).astype(np.float32))
    else:
        hi = tf32_rn(x32)
        lo = tf32_rn((x.astype(np.float64) - hi.astype(np.float64)).astype(np.float32))   # remainder of the fp64 value
    return hi.astype(np.float64), lo.astype(np.float64)


def to_f32(x, mode):
    r = x.astype(np.float32)
    if mode == "rz":
        over = np.abs(r.astype(np.float64)) > np.abs(x)
        r = np.where(over, np.nextafter(r, np.float32(0)), r)
    return r


def solve(chain, mode, kstep=8):
    nb = n // B
    Vh = np.zeros((n, m))
    Vl = np.zeros((n, m))
    Vout = np.zeros((n, m))
    Lh, Ll = split(-L)
    for I in range(nb):
        r0, r1 = I * B, (I + 1) * B
        Dinv = solve_triangular(L[r0:r1, r0:r1], np.eye(B), lower=True)
        R = Kt[r0:r1].copy()
        acc = np.zeros((B, m), dtype=np.float32)
        for J in range(I):
            for k0 in range(J * B, (J + 1) * B, kstep):
                ks = slice(k0, k0 + kstep)
                p = Ll[r0:r1, ks] @ Vh[ks] + Lh[r0:r1, ks] @ Vl[ks] + Lh[r0:r1, ks] @ Vh[ks]
                acc = to_f32(acc.astype(np.float64) + p, mode)
            if (J + 1) % chain == 0 or J + 1 == I:
                R += acc.astype(np.float64)
                acc[:] = 0
        Vi = Dinv @ R
        Vout[r0:r1] = Vi
        Vh[r0:r1], Vl[r0:r1] = split(Vi)
    return Vout


for SPLIT, mode, chain in (("trunc", "rn", 1), ("rn", "rn", 1), ("rn", "rn", 16), ("rn", "rz", 1), ("rn", "rz", 2), ("rn", "rz", 4),
                           ("rn", "rz", 16)):
    if True:
        Vb = solve(chain, mode)
        var = 1.0 - np.einsum("ij,ij->j", Vb, Vb)
        err = np.abs(var - var_ref)
        print(f"split={SPLIT} acc={mode} chain={chain:2d} J-blocks ({chain * 48:4d} MMAs): max |dvar|/prior {err.max():.3e}  median {np.median(err):.3e}  "
              f"max rel {np.max(err / np.abs(var_ref)):.3e}")
