"""ctypes binding of the C ABI in include/bopy_b200.h (bopy_b200/lib/libbopy_b200.so).

This is the only module that touches the shared library.  It fails loudly: a missing library or
a non-zero status raises NativeLibraryError -- there is no CPU fallback behind it.
PyTorch is used for device buffers and streams only.
"""
import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int64, c_uint64, c_void_p

from .exceptions import NativeLibraryError

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libbopy_b200.so")
ABI_VERSION = 2

OK, ERR_BAD_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOT_READY, ERR_NOT_POSITIVE_DEFINITE, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6
F64, F32 = 0, 1
KERNEL_IDS = {"rbf": 0, "matern12": 1, "matern32": 2, "matern52": 3}
ACQ_IDS = {None: -1, "none": -1, "lcb": 0, "ei": 1, "poi": 2}
PEAK_IDS = {"fp64_fma": 0, "fp32_fma": 1, "fp64_mma": 2, "tf32_mma_sync": 3, "tf32_tcgen05": 4}
NAN_POLICY_IDS = {"first": 0, "skip": 1}

# name -> (restype, argtypes); mirrors include/bopy_b200.h line by line
_SIGNATURES = {
    "bopy_abi_version": (c_int, []),
    "bopy_last_error": (c_char_p, []),
    "bopy_gp_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int64, c_int]),
    "bopy_gp_destroy": (None, [c_void_p]),
    "bopy_gp_resize": (c_int, [c_void_p, c_int64]),
    "bopy_gp_set_state": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_double), c_int, c_double,
                                  c_double, c_double, c_double, c_void_p]),
    "bopy_gp_fit": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_double), c_int, c_double, c_double, c_double,
                            c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "bopy_gp_append": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_double, c_void_p, c_void_p]),
    "bopy_gp_truncate": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_double, c_double, c_void_p, c_void_p]),
    "bopy_gp_lml": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_double), c_int, c_double, c_double, c_double,
                            POINTER(c_double), POINTER(c_double), c_void_p]),
    "bopy_gp_posterior_acq": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_double, c_double, c_void_p, c_void_p,
                                      c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "bopy_gp_probe_trace": (c_int, [c_void_p, POINTER(c_int64)]),
    "bopy_gp_set_latency_path": (c_int, [c_void_p, c_int64, POINTER(c_int64)]),
    "bopy_gp_set_inverse_path": (c_int, [c_void_p, c_int, POINTER(c_int64)]),
    "bopy_gp_set_group_mode": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int)]),
    "bopy_group_schedule": (c_int, [c_int64, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "bopy_acq_value_and_grad": (c_int, [c_void_p, c_int, c_double, c_double, c_void_p, c_int64, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "bopy_acq_eval_host": (c_int, [c_void_p, c_int, c_double, c_double, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "bopy_gp_predict_diag": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "bopy_acq_eval": (c_int, [c_void_p, c_int, c_double, c_double, c_void_p, c_int64, c_void_p, c_void_p]),
    "bopy_acq_argmin": (c_int, [c_void_p, c_int, c_double, c_double, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                                c_void_p]),
    "bopy_acq_argmin_pruned": (c_int, [c_void_p, c_int, c_double, c_double, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                                       POINTER(c_int64), c_void_p]),
    "bopy_acq_segment_argmin": (c_int, [c_void_p, c_int, c_double, c_double, c_void_p, c_int64, c_int64, c_int64,
                                        c_void_p, c_void_p, c_void_p]),
    "bopy_acq_segment_argmin_pruned": (c_int, [c_void_p, c_int, c_double, c_double, c_void_p, c_int64, c_int64, c_int64,
                                               c_void_p, c_void_p, POINTER(c_int64), c_void_p]),
    "bopy_candidates_around": (c_int, [c_uint64, c_void_p, c_int64, c_int, c_int, POINTER(c_double), POINTER(c_double),
                                       POINTER(c_double), c_void_p, c_void_p]),
    "bopy_multistart_step": (c_int, [c_int64, c_int, POINTER(c_double), POINTER(c_double), c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "bopy_gather_rows": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "bopy_acq_from_moments": (c_int, [c_int, c_double, c_double, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                      c_void_p, c_void_p, c_void_p]),
    "bopy_gp_predict_cov": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "bopy_candidates_uniform": (c_int, [c_uint64, c_int64, c_int64, c_int, POINTER(c_double), POINTER(c_double),
                                        c_void_p, c_void_p]),
    "bopy_measure_peak": (c_int, [c_int, POINTER(c_double)]),
    "bopy_gp_set_nan_policy": (c_int, [c_void_p, c_int]),
    "bopy_multistart_refine": (c_int, [c_void_p, c_int, c_double, c_double, c_int64, POINTER(c_double), POINTER(c_double),
                                       c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "bopy_topk_min_distance": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_double, POINTER(c_double), c_void_p,
                                       c_void_p, c_void_p]),
    "bopy_comm_unique_id": (c_int, [c_void_p, c_int64]),
    "bopy_comm_create": (c_int, [POINTER(c_void_p), c_void_p, c_int, c_int, c_int]),
    "bopy_comm_destroy": (None, [c_void_p]),
    "bopy_minloc_allreduce": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "bopy_gp_launch_info": (c_int, [c_void_p, c_int64, POINTER(c_int), POINTER(c_int), POINTER(c_int64)]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def group_schedule(job, n_blocks, lead):
    """(tile sequence number, block row) of a group's job-th unit of work in group mode (host-only helper, for tests)."""
    k, i = c_int(), c_int()
    check(load().bopy_group_schedule(int(job), int(n_blocks), int(lead), byref(k), byref(i)), "bopy_group_schedule")
    return k.value, i.value


def load():
    """Load the shared library once; raise NativeLibraryError if it is missing or mismatched."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m bopy_b200.build` (nvcc, sm_100a). "
            "bopy_b200 has no CPU fallback for the posterior/acquisition path.")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (restype, argtypes) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from exc
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.bopy_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"ABI version mismatch: library {lib.bopy_abi_version()}, binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(status, what):
    if status != OK:
        msg = load().bopy_last_error().decode("utf-8", "replace")
        if status == ERR_NOT_POSITIVE_DEFINITE:
            import numpy as np
            raise np.linalg.LinAlgError(msg)   # what scipy's cholesky raises inside sklearn's fit
        raise NativeLibraryError(f"{what} failed with status {status}: {msg}")


def _ptr(t):
    """Raw device pointer of a torch tensor, or NULL."""
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


def _stream(device):
    """torch's current stream on `device` as a raw cudaStream_t."""
    import torch
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None and getattr(device, "index", None) is not None:
        return c_void_p(raw(device.index))
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def resolve_device(device=None):
    """torch.device('cuda', i) for None | int | str | torch.device."""
    torch = require_cuda()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    if isinstance(device, int):
        return torch.device("cuda", device)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise NativeLibraryError(f"bopy_b200 runs on CUDA devices only (got {dev})")
    return torch.device("cuda", torch.cuda.current_device() if dev.index is None else dev.index)


_CUDA_SEEN = False


def require_cuda():
    """torch, after checking ONCE per process that a CUDA device is visible (the check costs microseconds, and a DIRECT
    probe is 35 of them)."""
    global _CUDA_SEEN
    import torch
    if not _CUDA_SEEN:
        if not torch.cuda.is_available():
            raise NativeLibraryError("no CUDA device is visible; bopy_b200 runs the posterior/acquisition path on a "
                                     "B200 only (no CPU fallback)")
        _CUDA_SEEN = True
    return torch


class NativeGP:
    """Owner of one `bopy_gp*` handle (one per device; not thread-safe)."""

    def __init__(self, n, d, kernel="rbf", dtype="f64", device=None):
        torch = require_cuda()
        self.lib = load()
        self.device = resolve_device(device)
        self.n, self.d = int(n), int(d)
        self.dtype = {"f64": F64, "f32": F32}[dtype]
        self.dtype_name = dtype
        self.kernel = kernel
        handle = c_void_p()
        check(self.lib.bopy_gp_create(byref(handle), self.device.index, self.dtype, KERNEL_IDS[kernel], self.n, self.d),
              "bopy_gp_create")
        self._handle = handle

    def close(self):
        """Free the handle's device memory (idempotent)."""
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            try:
                self.lib.bopy_gp_destroy(handle)
            except Exception:      # interpreter shutdown: the library may already be gone
                pass

    def __del__(self):
        self.close()

    def _dev64(self, a, shape=None):
        import numpy as np
        torch = require_cuda()
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=torch.float64, non_blocking=True).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {shape}, got {tuple(t.shape)}")
        return t

    def set_state(self, X, L, alpha, length_scale, amplitude=1.0, noise_level=0.0, y_mean=0.0, y_std=1.0):
        """Upload (X_train_, L_, alpha_) and the kernel hyper-parameters; packs L on the device."""
        import numpy as np
        torch = require_cuda()
        Xd = self._dev64(X, (self.n, self.d))
        Ld = self._dev64(L, (self.n, self.n))
        ad = self._dev64(alpha, (self.n,))
        ls = np.atleast_1d(np.asarray(length_scale, dtype=np.float64))
        ls_c = (c_double * len(ls))(*ls.tolist())
        with torch.cuda.device(self.device):
            check(self.lib.bopy_gp_set_state(self._handle, _ptr(Xd), _ptr(Ld), _ptr(ad), ls_c, len(ls),
                                             float(amplitude), float(noise_level), float(y_mean), float(y_std),
                                             _stream(self.device)), "bopy_gp_set_state")
        # the library packed private copies and synchronised: nothing needs to stay alive

    def fit(self, X, y_normalised, length_scale, amplitude=1.0, noise_level=0.0, alpha_reg=1e-10, y_mean=0.0,
            y_std=1.0, want_factor=False):
        """Fixed-hyper-parameter fit on the device (Gram + Cholesky + alpha) and state installation.

        Returns (alpha (n,) device tensor, L (n, n) device tensor or None)."""
        import numpy as np
        torch = require_cuda()
        Xd = self._dev64(X, (self.n, self.d))
        yd = self._dev64(y_normalised, (self.n,))
        ls = np.atleast_1d(np.asarray(length_scale, dtype=np.float64))
        ls_c = (c_double * len(ls))(*ls.tolist())
        alpha = torch.empty(self.n, dtype=torch.float64, device=self.device)
        L = torch.empty((self.n, self.n), dtype=torch.float64, device=self.device) if want_factor else None
        with torch.cuda.device(self.device):
            check(self.lib.bopy_gp_fit(self._handle, _ptr(Xd), _ptr(yd), ls_c, len(ls), float(amplitude),
                                       float(noise_level), float(alpha_reg), float(y_mean), float(y_std), _ptr(L),
                                       _ptr(alpha), _stream(self.device)), "bopy_gp_fit")
        return alpha, L

    def append(self, X, y_normalised, y_mean=0.0, y_std=1.0):
        """Grow the fitted data set by one point (the last row of X) without refactorising: bopy_gp_append.
        X (n+1, d), y_normalised (n+1,).  Returns alpha (n+1,) as a device tensor."""
        torch = require_cuda()
        Xd = self._dev64(X, (self.n + 1, self.d))
        yd = self._dev64(y_normalised, (self.n + 1,))
        alpha = torch.empty(self.n + 1, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.bopy_gp_append(self._handle, _ptr(Xd), _ptr(yd), float(y_mean), float(y_std), _ptr(alpha),
                                          _stream(self.device)), "bopy_gp_append")
        self.n += 1
        return alpha

    def truncate(self, X, y_normalised, y_mean=0.0, y_std=1.0):
        """Keep only the first len(X) points of the fitted data set (a leading subset): bopy_gp_truncate.
        Returns alpha (n_new,) as a device tensor."""
        torch = require_cuda()
        n_new = len(X)
        Xd = self._dev64(X, (n_new, self.d))
        yd = self._dev64(y_normalised, (n_new,))
        alpha = torch.empty(n_new, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.bopy_gp_truncate(self._handle, n_new, _ptr(Xd), _ptr(yd), float(y_mean), float(y_std),
                                            _ptr(alpha), _stream(self.device)), "bopy_gp_truncate")
        self.n = n_new
        return alpha

    def lml(self, X, y_normalised, length_scale, amplitude=1.0, noise_level=0.0, alpha_reg=1e-10, want_grad=True):
        """Log marginal likelihood (and gradient w.r.t. [log amplitude, log length_scale..., log noise_level]) on the
        device.  X / y_normalised may be device tensors (no copy).  Raises numpy.linalg.LinAlgError if K is not PD."""
        import numpy as np
        torch = require_cuda()
        Xd = self._dev64(X, (self.n, self.d))
        yd = self._dev64(y_normalised, (self.n,))
        ls = np.atleast_1d(np.asarray(length_scale, dtype=np.float64))
        ls_c = (c_double * len(ls))(*ls.tolist())
        value = c_double()
        grad = (c_double * (len(ls) + 2))() if want_grad else None
        with torch.cuda.device(self.device):
            check(self.lib.bopy_gp_lml(self._handle, _ptr(Xd), _ptr(yd), ls_c, len(ls), float(amplitude),
                                       float(noise_level), float(alpha_reg), byref(value), grad, _stream(self.device)),
                  "bopy_gp_lml")
        return value.value, (np.array(grad[:]) if want_grad else None)

    def resize(self, n):
        """Reuse the handle (and its workspaces) for another n with the same number of 128-row blocks; returns False
        if a new handle is needed.  The state has to be installed again."""
        n = int(n)
        if n < 1 or (n + 127) // 128 != (self.n + 127) // 128:
            return False
        check(self.lib.bopy_gp_resize(self._handle, n), "bopy_gp_resize")
        self.n = n
        return True

    def probe_trace(self, arm=False):
        """arm=True: record time stamps in the following latency-path launches; arm=False: fetch them as an
        (n_blocks, 8) int64 array of nanoseconds and disarm."""
        import numpy as np
        if arm:
            check(self.lib.bopy_gp_probe_trace(self._handle, None), "bopy_gp_probe_trace")
            return None
        nb = (self.n + 127) // 128
        buf = (c_int64 * (nb * 8))()
        check(self.lib.bopy_gp_probe_trace(self._handle, buf), "bopy_gp_probe_trace")
        return np.array(buf[:], dtype=np.int64).reshape(nb, 8)

    def set_latency_path(self, max_m):
        """Candidate sets of up to `max_m` rows take the latency path (probe_kernel); 0 switches it off.
        Returns the limit in force (0 for fp32 handles)."""
        eff = c_int64()
        check(self.lib.bopy_gp_set_latency_path(self._handle, int(max_m), byref(eff)), "bopy_gp_set_latency_path")
        return eff.value

    def set_inverse_path(self, mode=-1):
        """Calls of a handful of candidates as one product with W = L^-1 (probe_inv_kernel): -1 = W is built at the 16th
        such call on one state (default), 1 = at the first, 0 = never.  Returns the number of candidates per call that
        path serves on the current state (0 = none)."""
        eff = c_int64()
        check(self.lib.bopy_gp_set_inverse_path(self._handle, int(mode), byref(eff)), "bopy_gp_set_inverse_path")
        return eff.value

    def set_group_mode(self, group_size=-1, slots=-1, lead=-1):
        """Thread blocks per candidate tile of the fp64 throughput sweep (bopy_gp_set_group_mode): -1 = the default chosen
        from n, 0 = one tile per block, >= 2 = group mode, -2 = report only.  Returns (group_size, slots, lead) in force."""
        eff = (c_int * 3)()
        check(self.lib.bopy_gp_set_group_mode(self._handle, int(group_size), int(slots), int(lead), eff), "bopy_gp_set_group_mode")
        return tuple(eff[:])

    def gradient_capable(self):
        """True if bopy_acq_value_and_grad serves this handle (fp64 solve, block rows fit one wave of thread blocks)."""
        return self.dtype == F64 and (self.n + 127) // 128 <= 148

    def candidates(self, x):
        """(m, d) fp64 device tensor from numpy / torch input."""
        t = self._dev64(x)
        if t.dim() != 2 or t.shape[1] != self.d:
            raise ValueError(f"candidates must be (m, {self.d})")
        return t

    def sweep(self, Xs, acq=None, eta=0.0, kappa=2.0, want_mean=False, want_var=False, want_acq=False,
              want_min=False, index_base=0):
        """One fused launch over device candidates Xs (m, d).  Returns a dict of device tensors."""
        torch = require_cuda()
        m = Xs.shape[0]
        dev = self.device
        out = {}
        mean = torch.empty(m, dtype=torch.float64, device=dev) if want_mean else None
        var = torch.empty(m, dtype=torch.float64, device=dev) if want_var else None
        acqv = torch.empty(m, dtype=torch.float64, device=dev) if want_acq else None
        minv = torch.empty(1, dtype=torch.float64, device=dev) if want_min else None
        mini = torch.empty(1, dtype=torch.int64, device=dev) if want_min else None
        with torch.cuda.device(dev):
            check(self.lib.bopy_gp_posterior_acq(self._handle, _ptr(Xs), m, ACQ_IDS[acq], float(eta), float(kappa),
                                                 _ptr(mean), _ptr(var), _ptr(acqv), int(index_base), _ptr(minv),
                                                 _ptr(mini), _stream(dev)), "bopy_gp_posterior_acq")
        out.update(mean=mean, var=var, acq=acqv, min_val=minv, min_idx=mini)
        return out

    HOST_CALL_MAX_M = 4096

    def eval_host(self, x, acq=None, eta=0.0, kappa=2.0, want_acq=True, want_mean=False, want_var=False):
        """Small numpy candidate set in, numpy out, one native call (H2D, fused sweep, D2H, stream sync inside):
        the per-probe path of DIRECT-style callers.  x: C-contiguous float64 (m, d), m <= HOST_CALL_MAX_M."""
        import numpy as np
        torch = require_cuda()
        m = x.shape[0]
        out = [np.empty(m) if w else None for w in (want_acq, want_mean, want_var)]
        ptrs = [c_void_p(o.ctypes.data) if o is not None else c_void_p(0) for o in out]
        if torch.cuda.current_device() == self.device.index:      # (the library selects the handle's device itself)
            check(self.lib.bopy_acq_eval_host(self._handle, ACQ_IDS[acq], float(eta), float(kappa),
                                              c_void_p(x.ctypes.data), m, ptrs[0], ptrs[1], ptrs[2],
                                              _stream(self.device)), "bopy_acq_eval_host")
            return out
        with torch.cuda.device(self.device):
            check(self.lib.bopy_acq_eval_host(self._handle, ACQ_IDS[acq], float(eta), float(kappa),
                                              c_void_p(x.ctypes.data), m, ptrs[0], ptrs[1], ptrs[2],
                                              _stream(self.device)), "bopy_acq_eval_host")
        return out

    def value_and_grad(self, Xs, acq, eta=0.0, kappa=2.0):
        """Acquisition values (m,), their gradient w.r.t. the candidates (m, d), posterior mean and variance (m,):
        device tensors.  fp64 handles only."""
        torch = require_cuda()
        m = Xs.shape[0]
        dev = self.device
        val = torch.empty(m, dtype=torch.float64, device=dev)
        grad = torch.empty((m, self.d), dtype=torch.float64, device=dev)
        mean = torch.empty(m, dtype=torch.float64, device=dev)
        var = torch.empty(m, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.bopy_acq_value_and_grad(self._handle, ACQ_IDS[acq], float(eta), float(kappa), _ptr(Xs), m,
                                                   _ptr(val), _ptr(grad), _ptr(mean), _ptr(var), _stream(dev)),
                  "bopy_acq_value_and_grad")
        return val, grad, mean, var

    def argmin_pruned(self, Xs, acq, eta=0.0, kappa=2.0, index_base=0):
        """Branch-and-bound arg-min over device candidates Xs (m, d): (min_val, min_idx) device tensors and a dict
        with the number of candidates that went through the full sweep."""
        torch = require_cuda()
        m = Xs.shape[0]
        minv = torch.empty(1, dtype=torch.float64, device=self.device)
        mini = torch.empty(1, dtype=torch.int64, device=self.device)
        stats = (c_int64 * 3)()
        with torch.cuda.device(self.device):
            check(self.lib.bopy_acq_argmin_pruned(self._handle, ACQ_IDS[acq], float(eta), float(kappa), _ptr(Xs), m,
                                                  int(index_base), _ptr(minv), _ptr(mini), stats, _stream(self.device)),
                  "bopy_acq_argmin_pruned")
        return minv, mini, {"candidates": stats[0], "sample": stats[1], "swept": stats[2]}

    def segment_argmin(self, Xs, seg_len, acq, eta=0.0, kappa=2.0, index_base=0):
        """Per-segment arg-min of the acquisition over consecutive segments of `seg_len` rows of Xs (one launch).
        Returns (values (nseg,), indices (nseg,)) device tensors."""
        torch = require_cuda()
        m = Xs.shape[0]
        nseg = (m + seg_len - 1) // seg_len
        vals = torch.empty(nseg, dtype=torch.float64, device=self.device)
        idxs = torch.empty(nseg, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.bopy_acq_segment_argmin(self._handle, ACQ_IDS[acq], float(eta), float(kappa), _ptr(Xs), m,
                                                   int(seg_len), int(index_base), _ptr(vals), _ptr(idxs),
                                                   _stream(self.device)), "bopy_acq_segment_argmin")
        return vals, idxs

    def segment_argmin_pruned(self, Xs, seg_len, acq, eta=0.0, kappa=2.0, index_base=0):
        """segment_argmin by branch and bound: (values, indices, stats).  Xs.shape[0] must be a multiple of seg_len."""
        torch = require_cuda()
        m = Xs.shape[0]
        nseg = m // seg_len
        vals = torch.empty(nseg, dtype=torch.float64, device=self.device)
        idxs = torch.empty(nseg, dtype=torch.int64, device=self.device)
        stats = (c_int64 * 3)()
        with torch.cuda.device(self.device):
            check(self.lib.bopy_acq_segment_argmin_pruned(self._handle, ACQ_IDS[acq], float(eta), float(kappa), _ptr(Xs), m,
                                                          int(seg_len), int(index_base), _ptr(vals), _ptr(idxs), stats,
                                                          _stream(self.device)), "bopy_acq_segment_argmin_pruned")
        return vals, idxs, {"candidates": stats[0], "sample": stats[1], "swept": stats[2]}

    def predict_cov(self, Xs):
        torch = require_cuda()
        m = Xs.shape[0]
        mean = torch.empty(m, dtype=torch.float64, device=self.device)
        cov = torch.empty((m, m), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.bopy_gp_predict_cov(self._handle, _ptr(Xs), m, _ptr(mean), _ptr(cov), _stream(self.device)),
                  "bopy_gp_predict_cov")
        return mean, cov

    def set_nan_policy(self, policy):
        """'first': np.argmin's rule (a NaN acquisition value wins, the first one) -- the parity rule and the default;
        'skip': np.nanargmin's (a NaN never wins; index -1 when every value is NaN) -- what the optimisers use."""
        if getattr(self, "_nan_policy", "first") != policy:
            check(self.lib.bopy_gp_set_nan_policy(self._handle, NAN_POLICY_IDS[policy]), "bopy_gp_set_nan_policy")
            self._nan_policy = policy

    def multistart_refine(self, starts, acq, lowers, uppers, iterations, eta=0.0, kappa=2.0):
        """All `iterations` + 1 rounds of (value + gradient at the trial points, projected-gradient step) in ONE native call
        (`bopy_multistart_refine`): returns (x (S, d), f (S,)) device tensors.  fp64 handles only."""
        torch = require_cuda()
        S, d = starts.shape
        xt = starts.clone()
        xc = torch.empty_like(xt)
        fc = torch.empty(S, dtype=torch.float64, device=self.device)
        work = torch.empty(2 * S * d + 2 * S, dtype=torch.float64, device=self.device)
        lo = (c_double * d)(*[float(v) for v in lowers])
        hi = (c_double * d)(*[float(v) for v in uppers])
        with torch.cuda.device(self.device):
            check(self.lib.bopy_multistart_refine(self._handle, ACQ_IDS[acq], float(eta), float(kappa), S, lo, hi,
                                                  int(iterations), _ptr(xt), _ptr(xc), _ptr(fc), _ptr(work),
                                                  _stream(self.device)), "bopy_multistart_refine")
        return xc, fc

    def launch_info(self, m):
        g, l, w = c_int(), c_int(), c_int64()
        check(self.lib.bopy_gp_launch_info(self._handle, int(m), byref(g), byref(l), byref(w)), "bopy_gp_launch_info")
        return {"grid": g.value, "launches": l.value, "workspace_bytes": w.value}


def acquisition_from_moments(acq, mean, var, eta=0.0, kappa=2.0, want_min=False, index_base=0):
    """LCB/EI/POI epilogue (+ optional argmin) on device tensors mean, var (m,) fp64 -- the same device
    code as the fused sweep's epilogue, for surrogates that are not B200-native."""
    torch = require_cuda()
    lib = load()
    m = mean.shape[0]
    out = torch.empty(m, dtype=torch.float64, device=mean.device)
    minv = torch.empty(1, dtype=torch.float64, device=mean.device) if want_min else None
    mini = torch.empty(1, dtype=torch.int64, device=mean.device) if want_min else None
    with torch.cuda.device(mean.device):
        check(lib.bopy_acq_from_moments(ACQ_IDS[acq], float(eta), float(kappa), _ptr(mean), _ptr(var), m, _ptr(out),
                                        int(index_base), _ptr(minv), _ptr(mini), _stream(mean.device)),
              "bopy_acq_from_moments")
    return out, minv, mini


def candidates_uniform(seed, index_base, m, lowers, uppers, device=None):
    """Counter-based uniform candidates in a box, generated on the device: (m, d) fp64 tensor."""
    torch = require_cuda()
    lib = load()
    d = len(lowers)
    dev = resolve_device(device)
    out = torch.empty((int(m), d), dtype=torch.float64, device=dev)
    lo = (c_double * d)(*[float(v) for v in lowers])
    hi = (c_double * d)(*[float(v) for v in uppers])
    with torch.cuda.device(dev):
        check(lib.bopy_candidates_uniform(int(seed), int(index_base), int(m), d, lo, hi, _ptr(out), _stream(dev)),
              "bopy_candidates_uniform")
    return out


def candidates_around(seed, starts, P, halfwidth, lowers, uppers):
    """(S*P, d) device tensor: P points around each of the S starts (row s*P is the start itself)."""
    torch = require_cuda()
    lib = load()
    S, d = starts.shape
    out = torch.empty((S * int(P), d), dtype=torch.float64, device=starts.device)
    hw = (c_double * d)(*[float(v) for v in halfwidth])
    lo = (c_double * d)(*[float(v) for v in lowers])
    hi = (c_double * d)(*[float(v) for v in uppers])
    with torch.cuda.device(starts.device):
        check(lib.bopy_candidates_around(int(seed), _ptr(starts), S, int(P), d, hw, lo, hi, _ptr(out),
                                         _stream(starts.device)), "bopy_candidates_around")
    return out


def multistart_step(lowers, uppers, xc, fc, gc, xt, ft, gt, alpha, first):
    """One lock-step projected-gradient step of all starts (device tensors, updated in place)."""
    lib = load()
    S, d = xt.shape
    lo = (c_double * d)(*[float(v) for v in lowers])
    hi = (c_double * d)(*[float(v) for v in uppers])
    import torch
    with torch.cuda.device(xt.device):
        check(lib.bopy_multistart_step(S, d, lo, hi, _ptr(xc), _ptr(fc), _ptr(gc), _ptr(xt), _ptr(ft), _ptr(gt),
                                       _ptr(alpha), 1 if first else 0, _stream(xt.device)), "bopy_multistart_step")


def gather_rows(xs, idx, index_base=0):
    """xs[idx - index_base] as a new (S, d) device tensor."""
    torch = require_cuda()
    lib = load()
    S, d = idx.shape[0], xs.shape[1]
    out = torch.empty((S, d), dtype=torch.float64, device=xs.device)
    with torch.cuda.device(xs.device):
        check(lib.bopy_gather_rows(_ptr(xs), xs.shape[0], d, _ptr(idx), S, int(index_base), _ptr(out),
                                   _stream(xs.device)), "bopy_gather_rows")
    return out


def topk_min_distance(x, a, k, min_distance=0.0, scale=None):
    """Indices (k,) int64 and values (k,) of the k best evaluations a (N,) at points x (N, d) that keep `min_distance`
    from each other (coordinates divided by `scale`), greedy best first, on the device (`bopy_topk_min_distance`).
    Picks that do not exist (fewer than k qualify) have index -1."""
    torch = require_cuda()
    lib = load()
    N, d = x.shape
    idx = torch.empty(int(k), dtype=torch.int64, device=x.device)
    val = torch.empty(int(k), dtype=torch.float64, device=x.device)
    sc = None if scale is None else (c_double * d)(*[float(v) for v in scale])
    with torch.cuda.device(x.device):
        check(lib.bopy_topk_min_distance(_ptr(x), _ptr(a), N, d, int(k), float(min_distance), sc, _ptr(idx), _ptr(val),
                                         _stream(x.device)), "bopy_topk_min_distance")
    return idx, val


def measure_peak(what):
    """TFLOP/s of a register-resident FMA/MMA loop on the current device ('fp64_fma' | 'fp32_fma' | 'fp64_mma')."""
    require_cuda()
    v = c_double()
    check(load().bopy_measure_peak(PEAK_IDS[what], byref(v)), "bopy_measure_peak")
    return v.value
