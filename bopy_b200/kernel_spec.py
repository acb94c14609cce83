"""Translate a fitted scikit-learn kernel into the flat description the C ABI takes.

Supported (what bopy's users build, bopy/surrogate.py:72-91 + tests/test_surrogate.py:24-27):
``RBF``, ``Matern(nu in {0.5, 1.5, 2.5})``, optionally multiplied by ``ConstantKernel``s and
optionally summed with one ``WhiteKernel``.  Anything else raises: the B200 path never silently
approximates a kernel it does not implement.
"""
from dataclasses import dataclass

import numpy as np

_MATERN_IDS = {0.5: "matern12", 1.5: "matern32", 2.5: "matern52"}


@dataclass
class FlatKernel:
    kernel: str               # 'rbf' | 'matern12' | 'matern32' | 'matern52'  (bopy_kernel enum names)
    length_scale: np.ndarray  # (1,) isotropic or (d,) ARD
    amplitude: float          # product of the ConstantKernel factors
    noise_level: float        # WhiteKernel level; enters k(x, x) only


class UnsupportedKernelError(ValueError):
    pass


def flatten_sklearn_kernel(kernel) -> FlatKernel:
    from sklearn.gaussian_process import kernels as sk

    amplitude, noise, base = 1.0, 0.0, None

    def visit_factor(node):
        nonlocal amplitude, base
        if isinstance(node, sk.Product):
            visit_factor(node.k1)
            visit_factor(node.k2)
        elif isinstance(node, sk.ConstantKernel):
            amplitude *= float(node.constant_value)
        elif isinstance(node, (sk.Matern, sk.RBF)):
            if base is not None:
                raise UnsupportedKernelError("products of two stationary kernels are not supported")
            base = node
        else:
            raise UnsupportedKernelError(
                f"unsupported kernel component {node!r}: the B200 path implements Constant x RBF / Matern (+ White) "
                f"and has no CPU fallback")

    if isinstance(kernel, sk.Sum):
        terms = (kernel.k1, kernel.k2)
        whites = [t for t in terms if isinstance(t, sk.WhiteKernel)]
        others = [t for t in terms if not isinstance(t, sk.WhiteKernel)]
        if len(whites) != 1 or len(others) != 1:
            raise UnsupportedKernelError(f"only `stationary + WhiteKernel` sums are supported, got {kernel!r}")
        noise = float(whites[0].noise_level)
        visit_factor(others[0])
    else:
        visit_factor(kernel)
    if base is None:
        raise UnsupportedKernelError(f"no RBF / Matern component in {kernel!r}")
    if isinstance(base, sk.Matern):
        nu = float(base.nu)
        if nu == float("inf"):            # scikit-learn evaluates Matern(nu=inf) with the RBF formula ($SK/kernels.py:1730-1731)
            ls = np.atleast_1d(np.asarray(base.length_scale, dtype=np.float64)).copy()
            return FlatKernel(kernel="rbf", length_scale=ls, amplitude=amplitude, noise_level=noise)
        if nu not in _MATERN_IDS:
            raise UnsupportedKernelError(f"Matern nu={nu} is not supported on the device (0.5, 1.5, 2.5 and inf are)")
        name = _MATERN_IDS[nu]
    else:
        name = "rbf"
    ls = np.atleast_1d(np.asarray(base.length_scale, dtype=np.float64)).copy()
    return FlatKernel(kernel=name, length_scale=ls, amplitude=amplitude, noise_level=noise)


def theta_gradient(kernel, flat_grad: np.ndarray) -> np.ndarray:
    """Re-order a device gradient [d/dlog amplitude, d/dlog length_scale..., d/dlog noise_level] into the layout of
    `kernel.theta` (scikit-learn: the log of every NON-fixed hyper-parameter, in `kernel.hyperparameters` order)."""
    n_ls = len(flat_grad) - 2
    out = []
    for hp in kernel.hyperparameters:
        if hp.fixed:
            continue
        if hp.name.endswith("constant_value"):
            out.append(flat_grad[0])             # every constant factor c_i: dK/dlog c_i = K
        elif hp.name.endswith("length_scale"):
            if hp.n_elements != n_ls:
                raise UnsupportedKernelError("length_scale layout does not match the flattened kernel")
            out.extend(flat_grad[1:1 + n_ls])
        elif hp.name.endswith("noise_level"):
            out.append(flat_grad[1 + n_ls])
        else:
            raise UnsupportedKernelError(f"no device gradient for hyper-parameter {hp.name}")
    return np.asarray(out, dtype=np.float64)
