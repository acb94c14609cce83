"""Design study (CPU, numpy): accuracy of the blocked forward substitution used by the
CUDA path -- V_I = inv(L_II) (K_I - sum_{J<I} L_IJ V_J) -- against LAPACK dtrtrs,
in fp64 and in emulated fp32, on the C4-shaped problem (n=2048, d=6, RBF l=0.3, alpha=1e-6).
Run: python tools/numerics_study.py [n] [d] [ls]
"""
import sys, time
import numpy as np
from scipy.linalg import solve_triangular
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
d = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ls = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
alpha = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-6
m = 512
rng = np.random.default_rng(1234)
X = rng.random((n, d))
y = np.sin(X.sum(1) * 3) + 0.1 * rng.standard_normal(n)
gp = GaussianProcessRegressor(kernel=C(1.0) * RBF(ls * np.ones(d)), alpha=alpha, optimizer=None, normalize_y=True)
t = time.time(); gp.fit(X, y); print("fit s", time.time() - t)
L = gp.L_
Xs = np.random.default_rng(1235).random((m, d))
Kt = gp.kernel_(Xs, gp.X_train_)          # (m,n)
V = solve_triangular(L, Kt.T, lower=True, check_finite=False)
ss = np.einsum("ij,ij->j", V, V)
var_ref = 1.0 - ss
print("cond(L) ~", np.linalg.cond(L), " var quantiles", np.quantile(var_ref, [0, .01, .1, .5, .9, 1]))
# longdouble reference for the truth
def blocked(Lm, K, b, dtype, inv_dtype=np.float64, acc64=False):
    nn = Lm.shape[0]
    nb = (nn + b - 1) // b
    Ld = Lm.astype(dtype)
    Vout = np.zeros((nn, K.shape[1]), dtype=dtype)
    for I in range(nb):
        r0, r1 = I * b, min(nn, (I + 1) * b)
        Dinv = solve_triangular(Lm[r0:r1, r0:r1].astype(inv_dtype), np.eye(r1 - r0, dtype=inv_dtype), lower=True).astype(dtype)
        R = K[r0:r1].astype(dtype)
        if I > 0:
            # emulate kc-chunked accumulation in dtype
            for c0 in range(0, r0, 16):
                R = (R - Ld[r0:r1, c0:c0 + 16] @ Vout[c0:c0 + 16]).astype(dtype)
        Vout[r0:r1] = (Dinv @ R).astype(dtype)
    return Vout

for b in (16, 32, 64, 128, 256):
    Vb = blocked(L, Kt.T, b, np.float64)
    ssb = np.einsum("ij,ij->j", Vb, Vb)
    print(f"fp64 b={b:4d}: max rel dvar {np.max(np.abs(ssb - ss) / var_ref):.3e}  max |dV|/|V| {np.max(np.abs(Vb - V)) / np.max(np.abs(V)):.3e}")
# full explicit inverse
Linv = solve_triangular(L, np.eye(n), lower=True)
Vi = Linv @ Kt.T
ssi = np.einsum("ij,ij->j", Vi, Vi)
print(f"fp64 full inverse: max rel dvar {np.max(np.abs(ssi - ss) / var_ref):.3e}")
# fp32
K32 = gp.kernel_(Xs.astype(np.float32).astype(np.float64), gp.X_train_)  # ignore fp32 K error first
for b in (16, 32, 128):
    Vb = blocked(L, Kt.T, b, np.float32)
    ssb = np.einsum("ij,ij->j", Vb.astype(np.float64), Vb.astype(np.float64))
    rel = np.abs(ssb - ss) / var_ref
    print(f"fp32 b={b:4d}: rel dvar max {rel.max():.3e} median {np.median(rel):.3e}; abs max {np.max(np.abs(ssb-ss)):.3e}")
# plain fp32 forward substitution (sequential, fp32 everything) via scipy float32 LAPACK strtrs
V32 = solve_triangular(L.astype(np.float32), Kt.T.astype(np.float32), lower=True, check_finite=False)
ss32 = np.einsum("ij,ij->j", V32.astype(np.float64), V32.astype(np.float64))
rel = np.abs(ss32 - ss) / var_ref
print(f"fp32 strtrs: rel dvar max {rel.max():.3e} median {np.median(rel):.3e}; abs max {np.max(np.abs(ss32-ss)):.3e}")
mean_ref = Kt @ gp.alpha_
mean32 = (Kt.astype(np.float32) @ gp.alpha_.astype(np.float32)).astype(np.float64)
print("fp32 mean: max abs err", np.max(np.abs(mean32 - mean_ref)), " |alpha|max", np.abs(gp.alpha_).max(), "mean range", mean_ref.min(), mean_ref.max())

