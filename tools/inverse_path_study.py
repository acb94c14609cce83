"""How accurate is var = k(x,x) - |W k*|^2 with an EXPLICIT W = L^-1 (the inverse path, probe_inv_kernel.cuh) against
forward substitution, on every golden set of the unmodified reference?  CPU only (numpy / scipy, fp64): the yardstick
is the parity bound of tests/parity_util.py, 1e-9 |var_ref| + 1e-11 prior.  Output: profiles/r02/inverse_path_numerics.log"""
import os
import sys

import numpy as np
import scipy.linalg as sl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_names, golden_state  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402

for name in golden_names():
    g, st = golden_state(name)
    Xs = g["Xs"][:2048]
    Ks = O.kernel_cross(st.kernel, Xs, st.X_train).T
    n = st.L.shape[0]
    W = sl.solve_triangular(st.L, np.eye(n), lower=True)
    v1 = sl.solve_triangular(st.L, Ks, lower=True)
    v2 = W @ Ks
    prior = st.kernel.amplitude + st.kernel.noise_level
    var1, var2 = prior - (v1 * v1).sum(0), prior - (v2 * v2).sum(0)
    ref = g["var"][:Xs.shape[0]] / st.y_std ** 2
    bound = 1e-9 * np.abs(ref) + 1e-11 * prior
    print(f"{name:32s} n={n:5d} cond(L)={np.linalg.cond(st.L):9.2e}  substitution: worst err/bound "
          f"{np.nanmax(np.abs(var1 - ref) / bound):8.2e}   explicit inverse: {np.nanmax(np.abs(var2 - ref) / bound):8.2e}")

# ill-conditioned fits on purpose (long length scales, jitter down to 1e-10: cond(K) up to 1e13), against a substitution in
# extended precision (np.longdouble) as the reference: is there a conditioning at which the explicit inverse leaves the bound?
print()
rng = np.random.default_rng(0)
for n, d, ls, alpha in [(300, 2, 0.3, 1e-6), (300, 2, 0.6, 1e-6), (300, 2, 0.3, 1e-8), (300, 2, 0.6, 1e-8), (300, 2, 0.3, 1e-10),
                        (300, 2, 1.0, 1e-10), (600, 3, 0.5, 1e-8), (600, 3, 1.0, 1e-10), (1000, 6, 1.0, 1e-8), (1000, 6, 2.0, 1e-10)]:
    X = rng.random((n, d))
    y = np.sin(3 * X.sum(1))
    spec = O.KernelSpec(kind="rbf", length_scale=np.full(d, ls), amplitude=1.0)
    st = O.fit_state(X, y, spec, alpha, True)
    Xs = rng.random((512, d))
    Ks = O.kernel_cross(spec, Xs, X).T
    W = sl.solve_triangular(st.L, np.eye(n), lower=True)
    Lq, Kq = st.L.astype(np.longdouble), Ks.astype(np.longdouble)
    v = np.zeros_like(Kq)
    for i in range(n):
        v[i] = (Kq[i] - Lq[i, :i] @ v[:i]) / Lq[i, i]
    ref = (1.0 - (v * v).sum(0)).astype(np.float64)
    v1, v2 = sl.solve_triangular(st.L, Ks, lower=True), W @ Ks
    e1, e2 = np.abs(1 - (v1 * v1).sum(0) - ref), np.abs(1 - (v2 * v2).sum(0) - ref)
    bound = 1e-9 * np.abs(ref) + 1e-11
    print(f"synthetic n={n:4d} d={d} length scale {ls:3.1f} jitter {alpha:7.0e}: cond(L)={np.linalg.cond(st.L):8.2e}  "
          f"substitution: worst err/bound {np.max(e1 / bound):8.2e}   explicit inverse: {np.max(e2 / bound):8.2e}")
