"""The C ABI used from plain C (no Python, no torch): examples/c_abi_example.c is compiled with gcc against
include/bopy_b200.h + libbopy_b200.so + cudart, run, and its numbers compared with the Python path.  -m gpu."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_program_against_the_python_path(tmp_path):
    gcc = shutil.which("gcc")
    cuda = "/usr/local/cuda"
    if gcc is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("gcc / CUDA headers not available")
    from bopy_b200 import build
    lib_dir = os.path.dirname(build.build())
    exe = str(tmp_path / "c_abi_example")
    subprocess.run([gcc, "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
                    os.path.join(ROOT, "examples", "c_abi_example.c"), "-o", exe, "-L" + lib_dir, "-lbopy_b200",
                    "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + lib_dir], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    probes = np.array([[float(t) for t in ln.split()[3::2]] for ln in lines if ln.startswith("probe")])   # mean var ei
    idx, val = lines[-2].split()[1:3]
    assert lines[-1].split()[1:3] == [idx, val]                      # branch and bound: same winner

    # the same through Python
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel

    from bopy_b200 import _native
    from bopy_b200.acquisition import EI
    from bopy_b200.surrogate import B200GPSurrogate
    N, D, M = 300, 3, 5
    i = np.arange(N)[:, None]
    q = np.arange(D)[None, :]
    X = np.fmod(0.5 + (i + 1) * (0.6180339887498949 + 0.1 * q * q + 0.07 * q), 1.0)
    y = np.sin(4.0 * X[:, 0]) + X[:, 1] * X[:, 2]
    xs = 0.1 + 0.17 * np.arange(M)[:, None] + 0.05 * q
    sur = B200GPSurrogate(GaussianProcessRegressor(ConstantKernel(1.3) * RBF([0.3, 0.4, 0.5]), alpha=1e-6,
                                                    normalize_y=True, optimizer=None))
    sur.fit(X, y)
    ei = EI(sur)
    ei.fit(X, y)
    mean, var = sur.predict_diag(xs)
    np.testing.assert_allclose(probes[:, 0], mean, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(probes[:, 1], var, rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(probes[:, 2], ei(xs), rtol=1e-6, atol=1e-12)
    cands = _native.candidates_uniform(42, 0, 200000, np.zeros(D), np.ones(D))
    p_idx, p_val = ei.argmin(cands)
    assert p_idx == int(idx) and abs(p_val - float(val)) <= 1e-9 * abs(p_val)
