import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    """Load one fixture made by tools/make_golden.py; regenerate X from its seed when not stored."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    if "X" not in g:
        shape = tuple(int(v) for v in g["X_shape"])
        g["X"] = np.random.default_rng(int(g["X_seed"])).random(shape)
    import hashlib
    assert hashlib.sha256(np.ascontiguousarray(g["X"]).tobytes()).hexdigest() == str(g["X_sha256"])
    return g


_STATE_CACHE = {}


def golden_state(name):
    """Oracle GPState for a golden case (fixed-hyper-parameter refit with the stored kernel)."""
    from oracle import gp_oracle as O
    if name not in _STATE_CACHE:
        g = load_golden(name)
        spec = O.KernelSpec(kind=str(g["kernel_kind"]), nu=float(g["kernel_nu"]),
                            length_scale=np.asarray(g["length_scale"], dtype=np.float64),
                            amplitude=float(g["amplitude"]), noise_level=float(g["noise_level"]))
        st = O.fit_state(g["X"], g["y"], spec, float(g["alpha_reg"]), bool(g["normalize_y"]))
        _STATE_CACHE[name] = (g, st)
    return _STATE_CACHE[name]


@pytest.fixture(scope="session")
def golden_loader():
    return golden_state
