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
