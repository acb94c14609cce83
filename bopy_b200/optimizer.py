"""Acquisition optimisers: `optimize() -> OptimizationResult(x_min (b, d), f_min (b,))`.

The interface is the reference's (bopy/optimizer.py:17-67).  The reference's only base optimiser
probes the acquisition ONE point per call through Fortran DIRECT (bopy/optimizer.py:95-107); here
the base optimiser is `CandidateSweepOptimizer`: a large counter-based candidate set is generated
on the device, the fused posterior -> acquisition -> argmin kernel sweeps it, and optional zoom
rounds re-sweep shrinking boxes around the incumbent.  `DirectOptimizer` is kept for drop-in use
(scipy's DIRECT instead of the unmaintained `scipydirect`); the batch wrappers are unchanged logic.
"""
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Callable, Dict, Optional, Tuple

import numpy as np

from . import _native
from .acquisition import (AcquisitionFunction, OneShotBatchAcquisitionFunction,
                          SequentialBatchAcquisitionFunction)
from .bounds import Bounds


@dataclass
class OptimizationResult:
    """x_min: (batch_size, n_dimensions) argmin; f_min: (batch_size,) acquisition value there."""

    x_min: np.ndarray
    f_min: np.ndarray


class Optimizer(ABC):
    """Finds the minimiser of `acquisition_function` inside `bounds`."""

    def __init__(self, acquisition_function: AcquisitionFunction, bounds: Bounds):
        self.acquisition_function = acquisition_function
        self.bounds = bounds

    def optimize(self) -> OptimizationResult:
        x_min, f_min = self._optimize()
        return OptimizationResult(x_min=x_min, f_min=f_min)

    @abstractmethod
    def _optimize(self) -> Tuple[np.ndarray, np.ndarray]:
        ...


class CandidateSweepOptimizer(Optimizer):
    """Batched argmin of the acquisition over `n_candidates` uniform candidates (+ zoom rounds).

    Parameters
    ----------
    n_candidates : int
        Size of the global sweep.  Candidates come from the counter-based generator
        (`bopy_candidates_uniform`): candidate i depends only on (seed, i), so any sharding of the
        index range over GPUs sees the same set.
    zoom_rounds, zoom_candidates, zoom_shrink :
        After the global sweep, `zoom_rounds` local sweeps of `zoom_candidates` points in a box of
        half-width `zoom_shrink**k * (upper - lower) / 2` around the incumbent (clipped to the bounds);
        the incumbent itself is candidate 0 of every round, so the value never gets worse.
    seed : int
        Base seed; the k-th call to `optimize()` uses seed + k.
    process_group : torch.distributed group or None
        With a group, each rank sweeps its contiguous slice of the index range and ONE min-loc
        all-gather picks the winner (see bopy_b200/distributed.py).
    prune : bool
        Branch and bound inside every sweep (`bopy_acq_argmin_pruned`): candidates whose mean-only lower bound cannot
        beat the best of a strided sample skip the full posterior.  Same winner; see B200GPSurrogate.acquisition_argmin
        for the one caveat (NaN acquisition values).
    """

    def __init__(self, acquisition_function: AcquisitionFunction, bounds: Bounds, n_candidates: int = 1 << 20,
                 zoom_rounds: int = 0, zoom_candidates: int = 1 << 14, zoom_shrink: float = 0.25, seed: int = 0,
                 process_group=None, distributed: bool = False, prune: bool = False):
        super().__init__(acquisition_function, bounds)
        if n_candidates < 1:
            raise ValueError("`n_candidates` must be positive.")
        self.n_candidates = int(n_candidates)
        self.zoom_rounds = int(zoom_rounds)
        self.zoom_candidates = int(zoom_candidates)
        self.zoom_shrink = float(zoom_shrink)
        self.seed = int(seed)
        self.process_group = process_group
        self.distributed = distributed or process_group is not None
        self.prune = bool(prune)
        self._calls = 0

    def _sweep_box(self, seed, lowers, uppers, m, incumbent=None):
        """argmin over m candidates in [lowers, uppers); returns (x (d,), value)."""
        from .distributed import sharded_argmin
        acq = self.acquisition_function
        if self.distributed:
            return sharded_argmin(acq, seed, lowers, uppers, m, group=self.process_group, incumbent=incumbent,
                                  prune=self.prune)
        torch = _native.require_cuda()
        xs = _native.candidates_uniform(seed, 0, m, lowers, uppers)
        if incumbent is not None:
            xs[0] = torch.as_tensor(incumbent, dtype=torch.float64, device=xs.device)
        # an optimiser never proposes a candidate whose acquisition value is NaN (posterior variance rounded to <= 0 on
        # top of a training point): np.nanargmin's rule; np.argmin's (first NaN) only if every candidate is NaN
        idx, val = acq.argmin(xs, prune=self.prune, nan_policy="skip")
        if idx < 0:
            idx, val = acq.argmin(xs, nan_policy="first")
        return xs[idx].cpu().numpy(), val

    def _optimize(self) -> Tuple[np.ndarray, np.ndarray]:
        lo = np.asarray(self.bounds.lowers, dtype=np.float64)
        hi = np.asarray(self.bounds.uppers, dtype=np.float64)
        seed = self.seed + 7919 * self._calls
        self._calls += 1
        x, val = self._sweep_box(seed, lo, hi, self.n_candidates)
        half = (hi - lo) / 2.0
        for k in range(1, self.zoom_rounds + 1):
            half = half * self.zoom_shrink
            zl, zu = np.maximum(lo, x - half), np.minimum(hi, x + half)
            x, val = self._sweep_box(seed + k, zl, zu, self.zoom_candidates, incumbent=x)
        return np.array([x]), np.array([val])


class MultiStartOptimizer(Optimizer):
    """Batched multi-start minimisation of the acquisition: all starts advance together, one launch per round.

    1. A global sweep of `n_candidates` counter-based candidates is cut into `n_starts` segments; the fused
       segmented argmin returns the best point of every segment: `n_starts` stratified starts.
    2. Refinement, `method`:
       'gradient'  `iterations` lock-step steps of a monotone projected-gradient method with Barzilai-Borwein step
                   lengths (`bopy_multistart_step`); value and gradient of the acquisition at all trial points come
                   from ONE device call per step (`bopy_acq_value_and_grad`: forward + mirrored backward solve).
       'cloud'     `rounds` times: a cloud of `points_per_start` points is drawn around every start in a box of
                   half-width `shrink**k * initial_halfwidth * (upper - lower)` (clipped to the bounds, the start
                   itself is point 0), all clouds are swept in ONE launch, and each start moves to the best point of
                   its own cloud.  Derivative-free; serves fp32 surrogates too.
       'auto'      'gradient' where the surrogate offers it (fp64 handle), else 'cloud'.
    3. The best start wins (`np.argmin` ordering).  `local_minima()` returns all refined starts of the last call.

    Everything between the first candidate and the final (x, value) stays on the device.  With a process group the
    starts are sharded over the ranks and one min-loc all-gather picks the winner.
    Needs a B200-native surrogate behind an LCB / EI / POI acquisition.
    """

    def __init__(self, acquisition_function: AcquisitionFunction, bounds: Bounds, n_starts: int = 256,
                 n_candidates: int = 1 << 20, rounds: int = 6, points_per_start: int = 256, shrink: float = 0.5,
                 initial_halfwidth: Optional[float] = None, seed: int = 0, process_group=None,
                 distributed: bool = False, method: str = "auto", iterations: int = 40, prune: bool = False):
        super().__init__(acquisition_function, bounds)
        if n_starts < 1 or rounds < 0 or iterations < 0:
            raise ValueError("`n_starts` must be positive and `rounds` / `iterations` non-negative.")
        if method not in ("auto", "gradient", "cloud"):
            raise ValueError("`method` must be 'auto', 'gradient' or 'cloud'.")
        self.method = method
        self.iterations = int(iterations)
        self.prune = bool(prune)    # branch and bound inside the global sweep (same starts; see acquisition_argmin)
        if points_per_start % 128 != 0 or points_per_start < 128:
            raise ValueError("`points_per_start` must be a positive multiple of 128.")
        self.n_starts = int(n_starts)
        seg = max(128, (int(n_candidates) // self.n_starts) // 128 * 128)
        self.segment = seg                       # candidates per start in the global sweep (multiple of 128)
        self.n_candidates = seg * self.n_starts
        self.rounds = int(rounds)
        self.points_per_start = int(points_per_start)
        self.shrink = float(shrink)
        self.initial_halfwidth = initial_halfwidth
        self.seed = int(seed)
        self.process_group = process_group
        self.distributed = distributed or process_group is not None
        self._calls = 0
        self._last = None

    def _inner(self):
        acq = self.acquisition_function
        while hasattr(acq, "base_acquisition"):
            acq = acq.base_acquisition
        sur = acq.surrogate
        if not (hasattr(acq, "kind") and hasattr(sur, "acquisition_segment_argmin")):
            raise TypeError("MultiStartOptimizer needs an LCB / EI / POI acquisition on a B200GPSurrogate")
        return acq, sur

    def local_minima(self) -> Tuple[np.ndarray, np.ndarray]:
        """(x (n_starts_local, d), values (n_starts_local,)) of the refined starts of the last `optimize()`."""
        if self._last is None:
            raise RuntimeError("call optimize() first")
        return self._last

    def _optimize(self) -> Tuple[np.ndarray, np.ndarray]:
        import torch

        from .distributed import all_reduce_minloc, shard_range
        acq, sur = self._inner()
        args = acq.native_args()
        lo = np.asarray(self.bounds.lowers, dtype=np.float64)
        hi = np.asarray(self.bounds.uppers, dtype=np.float64)
        d = len(lo)
        seed = self.seed + 104729 * self._calls
        self._calls += 1
        rank, world = 0, 1
        if self.distributed:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(self.process_group), dist.get_world_size(self.process_group)
        s0, s1 = shard_range(self.n_starts, rank, world)      # this rank's starts
        value, index, x_best = 0.0, -1, None
        if s1 > s0:
            base = s0 * self.segment
            xs = _native.candidates_uniform(seed, base, (s1 - s0) * self.segment, lo, hi)
            vals, idxs = sur.acquisition_segment_argmin(acq.kind, xs, self.segment, index_base=base, prune=self.prune,
                                                        nan_policy="skip", **args)
            # a segment whose candidates are all NaN has no arg-min (index -1): start from its first candidate
            seg_first = base + torch.arange(idxs.shape[0], dtype=torch.int64, device=idxs.device) * self.segment
            idxs = torch.where(idxs < 0, seg_first, idxs)
            starts = _native.gather_rows(xs, idxs, index_base=base)
            del xs
            method = self.method
            if method == "auto":
                method = "gradient" if getattr(sur, "supports_gradient", lambda: False)() else "cloud"
            if method == "gradient":
                # all starts advance together, every iteration = one value+gradient pass and one step kernel; the whole
                # loop is ONE native call (bopy_multistart_refine): its launches queue back to back, no host round trips
                native = sur.native
                if hasattr(native, "multistart_refine"):
                    starts, vals = native.multistart_refine(starts, acq.kind, lo, hi, self.iterations, **args)
                else:       # a surrogate that only offers value_and_grad: the same loop, one call pair per iteration
                    xt = starts.clone()
                    xc, gc = torch.empty_like(xt), torch.empty_like(xt)
                    fc = torch.empty(xt.shape[0], dtype=torch.float64, device=xt.device)
                    alpha = torch.ones_like(fc)
                    for k in range(self.iterations + 1):
                        ft, gt, _, _ = native.value_and_grad(xt, acq.kind, **args)
                        _native.multistart_step(lo, hi, xc, fc, gc, xt, ft, gt, alpha, first=(k == 0))
                    starts, vals = xc, fc
            else:
                # a start owns ~1/n_starts of the box: begin with a cloud of that size
                frac = self.initial_halfwidth if self.initial_halfwidth is not None else \
                    min(0.5, 0.5 * (1.0 / self.n_starts) ** (1.0 / d) * 2.0)
                half = (hi - lo) * frac
                P = self.points_per_start
                for k in range(self.rounds):
                    cloud = _native.candidates_around(seed + 1 + k + 1000 * rank, starts, P, half, lo, hi)
                    vals, idxs = sur.acquisition_segment_argmin(acq.kind, cloud, P, nan_policy="skip", **args)
                    own = torch.arange(idxs.shape[0], dtype=torch.int64, device=idxs.device) * P    # row s*P is start s
                    idxs = torch.where(idxs < 0, own, idxs)
                    starts = _native.gather_rows(cloud, idxs)
                    half = half * self.shrink
            v = vals.cpu().numpy()
            self._last = (starts.cpu().numpy(), v)
            j = int(np.argmin(v)) if np.isnan(v).all() else int(np.nanargmin(v))     # a NaN start never wins
            value, index, x_best = float(v[j]), s0 + j, self._last[0][j]
        if world > 1:
            value, index = all_reduce_minloc(value, index, group=self.process_group)
            owner = next(r for r in range(world) if shard_range(self.n_starts, r, world)[0] <= index
                         < shard_range(self.n_starts, r, world)[1])
            import torch.distributed as dist
            buf = torch.as_tensor(x_best if rank == owner else np.zeros(d), dtype=torch.float64)
            if dist.get_backend(self.process_group) != "gloo":
                buf = buf.cuda()
            dist.broadcast(buf, src=dist.get_global_rank(self.process_group, owner) if self.process_group is not None
                           else owner, group=self.process_group)
            x_best = buf.cpu().numpy()
        return np.array([x_best]), np.array([value])


class DirectOptimizer(Optimizer):
    """DIRECT global optimiser, one acquisition probe per call (bopy/optimizer.py:70-107).

    The reference binds the Fortran `scipydirect.minimize`; this uses `scipy.optimize.direct`
    (the same algorithm).  `direct_kwargs` accepts scipydirect's names and maps them onto scipy's:
    `maxf` -> `maxfun`, `maxT` -> `maxiter`, `algmethod` -> `locally_biased`, `fglobal` -> `f_min`,
    `fglper` (per cent) -> `f_min_rtol`, `volper` -> `vol_tol`, `sigmaper` -> `len_tol`; `disp` / `logfilename` have no
    counterpart and are ignored.  Keys that are absent take scipydirect's defaults (original DIRECT, maxf = 20000,
    maxT = 6000), not scipy's (DIRECT_L, 1000 d evaluations, 1000 iterations).
    """

    _RENAMED = {"maxf": "maxfun", "maxT": "maxiter", "algmethod": "locally_biased", "fglobal": "f_min",
                "fglper": "f_min_rtol", "volper": "vol_tol", "sigmaper": "len_tol"}
    _IGNORED = ("disp", "logfilename")
    _DEFAULTS = {"locally_biased": False, "maxfun": 20000, "maxiter": 6000}

    def __init__(self, acquisition_function: AcquisitionFunction, bounds: Bounds, **direct_kwargs: Dict[str, Any]):
        super().__init__(acquisition_function, bounds)
        self.direct_kwargs = direct_kwargs

    def _optimize(self) -> Tuple[np.ndarray, np.ndarray]:
        from scipy.optimize import direct

        kwargs = dict(self._DEFAULTS)
        for key, value in self.direct_kwargs.items():
            if key in self._IGNORED:
                continue
            name = self._RENAMED.get(key, key)
            if name == "locally_biased":
                value = bool(value)
            elif name == "f_min_rtol":
                value = float(value) / 100.0          # scipydirect: per cent
            elif name in ("vol_tol", "len_tol") and float(value) < 0:
                continue                              # scipydirect's "-1 = off": scipy's own (tiny) default
            kwargs[name] = value

        def objective(point):
            return float(self.acquisition_function(np.asarray(point, dtype=np.float64).reshape(1, -1))[0])

        # DIRECT probes ONE point per call, `maxfun` times on one fitted state: with a budget well beyond the point where
        # the library would switch by itself, build W = L^-1 at the first probe (bopy_gp_set_inverse_path)
        surrogate = getattr(self.acquisition_function, "surrogate", None)
        native = getattr(surrogate, "native", None)
        eager = (native is not None and getattr(surrogate, "inverse_path", None) == "auto"
                 and kwargs["maxfun"] >= 64)
        if eager:
            native.set_inverse_path(1)
        try:
            res = direct(objective, bounds=list(zip(self.bounds.lowers, self.bounds.uppers)), **kwargs)
        finally:
            if eager:
                native.set_inverse_path(-1)
        return np.array([res.x]), np.array([res.fun])


class SequentialBatchOptimizer(Optimizer):
    """Builds a batch by optimising / updating a SequentialBatchAcquisitionFunction `batch_size` times
    (bopy/optimizer.py:110-167)."""

    def __init__(self, acquisition_function: SequentialBatchAcquisitionFunction, bounds: Bounds,
                 base_optimizer: Optimizer, batch_size: int):
        super().__init__(acquisition_function, bounds)
        self.base_optimizer = base_optimizer
        self.batch_size = batch_size
        self.x_mins = []
        self.f_mins = []

    def start_batch(self) -> None:
        self.x_mins, self.f_mins = [], []

    def add_to_batch(self, optimization_result: OptimizationResult) -> None:
        self.x_mins.append(optimization_result.x_min)
        self.f_mins.append(optimization_result.f_min)

    def get_batch(self) -> Tuple[np.ndarray, np.ndarray]:
        return np.concatenate(self.x_mins), np.concatenate(self.f_mins)

    def _optimize(self) -> Tuple[np.ndarray, np.ndarray]:
        acq = self.acquisition_function
        self.start_batch()
        acq.start_batch()
        for _ in range(self.batch_size):
            picked = self.base_optimizer.optimize()
            self.add_to_batch(picked)
            acq.add_to_batch(picked)
        acq.finish_batch()
        return self.get_batch()


class OneShotBatchOptimizerStrategy(ABC):
    """Chooses `batch_size` of the points one global optimisation pass evaluated."""

    @abstractmethod
    def select(self, x: np.ndarray, a_x: np.ndarray, batch_size: int) -> Tuple[np.ndarray, np.ndarray]:
        raise NotImplementedError


class OneShotBatchOptimizerRandomSamplingStrategy(OneShotBatchOptimizerStrategy):
    """Uniformly random subset (with replacement, like the reference: bopy/optimizer.py:186-197)."""

    def select(self, x, a_x, batch_size):
        chosen = np.random.choice(len(x), size=batch_size)
        return x[chosen], a_x[chosen]


class OneShotBatchOptimizerTopKStrategy(OneShotBatchOptimizerStrategy):
    """The `batch_size` best evaluations that are at least `min_distance` apart (greedy, best first; distance in units
    of the box, i.e. after scaling every coordinate by `scale`).  The natural strategy when the base optimiser is a
    device sweep that logged a million evaluations: the selection runs on the device-resident log
    (`select_on_device`), only the batch comes back."""

    def __init__(self, min_distance: float = 0.0, scale=None):
        super().__init__()
        self.min_distance = float(min_distance)
        self.scale = scale

    def select(self, x, a_x, batch_size):
        order = np.argsort(a_x, kind="stable")
        scale = np.ones(x.shape[1]) if self.scale is None else np.asarray(self.scale, dtype=np.float64)
        chosen = []
        for i in order:
            if np.isnan(a_x[i]):
                continue
            if all(np.linalg.norm((x[i] - x[j]) / scale) >= self.min_distance for j in chosen):
                chosen.append(int(i))
                if len(chosen) == batch_size:
                    break
        chosen = np.array(chosen, dtype=np.int64)
        return x[chosen], a_x[chosen]

    def select_on_device(self, x, a_x, batch_size):
        """Same rule on the device-resident log (x (N, d), a_x (N,) device tensors): `bopy_topk_min_distance` -- one pass
        over the log per pick (candidates too close to the newest pick die, the best survivor is next), no host round
        trip until the batch itself is copied back."""
        import torch
        x = x.contiguous()
        a_x = a_x.contiguous()
        idx, _ = _native.topk_min_distance(x, a_x, batch_size, self.min_distance, self.scale)
        idx = idx[idx >= 0]
        return x[idx].cpu().numpy(), a_x[idx].cpu().numpy()


class OneShotBatchOptimizerKDPPSamplingStrategy(OneShotBatchOptimizerStrategy):
    """k-DPP sample with likelihood kernel(x) + alpha I (bopy/optimizer.py:200-232); needs `dppy`.  The likelihood matrix is
    N x N: with a device sweep's log of a million evaluations the `max_points` best ones are kept first."""

    def __init__(self, kernel: Callable[[np.ndarray], np.ndarray], alpha: float = 1e-5, max_points: int = 4096):
        super().__init__()
        self.kernel = kernel
        self.alpha = alpha
        self.max_points = int(max_points)

    def select(self, x, a_x, batch_size):
        if len(x) > self.max_points:
            keep = np.argsort(np.where(np.isnan(a_x), np.inf, a_x), kind="stable")[: self.max_points]
            x, a_x = x[keep], a_x[keep]
        try:
            from dppy.finite_dpps import FiniteDPP
        except ImportError as exc:  # the dependency is optional and absent from this image
            raise ImportError("the k-DPP strategy needs the `dppy` package") from exc
        dpp = FiniteDPP("likelihood", L=self.kernel(x) + self.alpha * np.eye(len(x)))
        dpp.sample_exact_k_dpp(size=batch_size)
        chosen = dpp.list_of_samples[0]
        return x[chosen], a_x[chosen]


class OneShotBatchOptimizer(Optimizer):
    """One global pass of `base_optimizer`, then `strategy` picks the batch from the logged evaluations
    (bopy/optimizer.py:235-276)."""

    def __init__(self, acquisition_function: OneShotBatchAcquisitionFunction, bounds: Bounds,
                 base_optimizer: Optimizer, batch_size: int, strategy: OneShotBatchOptimizerStrategy):
        super().__init__(acquisition_function, bounds)
        self.base_optimizer = base_optimizer
        self.batch_size = batch_size
        self.strategy = strategy

    def _optimize(self) -> Tuple[np.ndarray, np.ndarray]:
        if getattr(self.base_optimizer, "distributed", False):
            # a sharded sweep logs only the rank's own slice: every rank would pick a different batch and the replicated
            # surrogates would drift apart silently
            raise ValueError("OneShotBatchOptimizer needs the whole evaluation log on every rank: use a base optimizer "
                             "without `distributed=True` / `process_group`")
        self.acquisition_function.start_optimization()
        self.base_optimizer.optimize()
        if hasattr(self.strategy, "select_on_device") and hasattr(self.acquisition_function, "get_evaluations_on_device") \
                and any(not isinstance(a, np.ndarray) for a in self.acquisition_function.xs):
            xs, a_xs = self.acquisition_function.get_evaluations_on_device()   # a device sweep logged them
            return self.strategy.select_on_device(xs, a_xs, self.batch_size)
        xs, a_xs = self.acquisition_function.get_evaluations()
        return self.strategy.select(xs, a_xs, self.batch_size)
