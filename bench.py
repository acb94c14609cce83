#!/usr/bin/env python
"""Benchmark of the hot path: fused GP posterior + EI + argmin over a candidate batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Headline workload = BASELINE.json configs[3] ("Hartmann-6D synthetic, n=2048 observations, 16M candidates sharded over
8xB200"): n=2048 training points in d=6, kernel 1.0*RBF(0.3), alpha=1e-6, normalize_y; 2^21 candidates per GPU per step
(weak scaling: 16M at 8 GPUs), generated on the device from (seed, global index).  A step is one fused sweep of the
rank's candidates (K* tile -> blocked triangular solve -> variance -> EI -> argmin) plus the single min-loc exchange
between ranks (bopy_minloc_allreduce: ncclAllGather of 16-byte records + a one-warp kernel, on the sweep's stream).
One JSON line is printed by rank 0.

value   : candidates/s with the candidates already resident in HBM.
e2e     : the same through the public API with HOST candidates in pinned memory: H2D copy, sweep, D2H of the
          (index, value) result inside the timed region.
roofline: the sweep kernel is bound by the FP64 tensor sub-pipe (DMMA) (F(n,d) = n^2 + n(3d+5)
          flops per candidate against 48 B of HBM input); peak = max(DFMA, DMMA) rate measured live by
          bopy_measure_peak (MEASURED_PEAKS.json has no FP64 figure); `traffic` = ncu dram bytes of the same launch
          shape, read from profiles/r02/ncu_traffic.json (captured on the build that is timed).
configs : the other BASELINE.json configs on SURVEY.md section 8(d)'s inputs, each with throughput, roofline fraction and
          a parity spot-check of the stream's first candidates against tests/golden/*.npz (= outputs of the unmodified
          reference): C1 (Forrester, n=10, d=1), C3 (Branin box, n=256, l = 0.2 range, 2^20 candidates per GPU), C4 in
          fp32 mode (tcgen05 engine), C5 (n=8192, d=20, l=1.0, 2^23 candidates per GPU = 64M at 8 GPUs, plus the 1024-start
          multi-start argmin sharded over the ranks).
strong_scaling : fixed totals (2^24 C4 candidates, 2^20 C3 candidates) split over the ranks, with the collective's share.
cpu_baseline / --impl reference: bopy's own call sequence (sklearn predict(return_cov=True) -> np.diag ->
          scipy.stats.norm EI, 64 candidates per call = the reference's best chunk) on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, DIM = 2048, 6
LENGTH_SCALE, ALPHA_REG = 0.3, 1e-6
CAND_PER_GPU = 1 << 21
SEED_TRAIN, SEED_CAND = 1234, 1235
METRIC = "fused posterior+EI candidate evals/sec (n=2048, d=6)"
UNIT = "evals/s"
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02", "ncu_traffic.json")


def flops_per_candidate(n, d):
    return n * n + n * (3 * d + 5)


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this shape from `ncu --set full` on the timed build
    (tools/ncu_summary.py writes the file); None when that shape was not captured."""
    try:
        with open(TRAFFIC_FILE) as f:
            return json.load(f).get(key, {}).get("dram_bytes")
    except (OSError, ValueError):
        return None


# ---- the BASELINE.json configs on SURVEY.md section 8(d)'s inputs (= what tools/make_golden.py froze) ---------------------
def config_problem(key):
    """(X, y, sklearn GP, lowers, uppers, golden fixture, device_fit) of one BASELINE config."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern

    from bopy_b200.benchmark_functions import branin, forrester, hartmann6
    rng = np.random.default_rng(SEED_TRAIN)
    if key == "C1":      # the reference's own fixture (tests/test_surrogate.py:11-27), fixed hyper-parameters
        X = np.linspace(0, 1, 10).reshape(-1, 1)
        gp = GaussianProcessRegressor(kernel=Matern(nu=1.5), alpha=1e-5, normalize_y=True, optimizer=None)
        return X, forrester(X), gp, np.zeros(1), np.ones(1), "ref_forrester_matern15_fixed", False
    if key == "C3":
        lo, hi = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
        X = lo + rng.random((256, 2)) * (hi - lo)
        gp = GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF([3.0, 3.0]), alpha=1e-6, normalize_y=True,
                                      optimizer=None)
        return X, branin(X), gp, lo, hi, "c3_branin_n256", False
    if key == "C4":
        X = rng.random((2048, 6))
        gp = GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(0.3 * np.ones(6)), alpha=1e-6, normalize_y=True,
                                      optimizer=None)
        return X, hartmann6(X), gp, np.zeros(6), np.ones(6), "c4_hartmann6_n2048", False
    if key == "C5":
        X = rng.random((8192, 20))
        y = np.sin(3.0 * X[:, :5].sum(1)) + 0.5 * np.cos(2.0 * X[:, 5:].sum(1) / 3.0)
        gp = GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(1.0 * np.ones(20)), alpha=1e-6, normalize_y=True,
                                      optimizer=None)
        return X, y, gp, np.zeros(20), np.ones(20), "c5_rbf_d20_n8192", True     # host fit: 7 s x ranks; device: 25 ms
    raise KeyError(key)


def make_problem(n=N_TRAIN, d=DIM):
    """The headline problem (C4); other (n, d) only for development runs."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel

    from bopy_b200.benchmark_functions import hartmann6
    rng = np.random.default_rng(SEED_TRAIN)
    X = rng.random((n, d))
    y = hartmann6(X) if d == 6 else np.sin(3.0 * X.sum(1))
    gp = GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(LENGTH_SCALE * np.ones(d)), alpha=ALPHA_REG,
                                  normalize_y=True, optimizer=None)
    return X, y, gp


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def host_info():
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return cores, model


def all_host_threads():
    """Context manager: let BLAS/OpenMP use every host core (torchrun exports OMP_NUM_THREADS=1)."""
    from threadpoolctl import threadpool_limits
    return threadpool_limits(limits=host_info()[0])


def time_reference_path(gp, eta, lowers, uppers, budget_s, chunk=64, base_index=0, max_candidates=1 << 17):
    """bopy's call sequence on the host: chunks of `chunk` candidates until `budget_s` seconds are used."""
    from oracle import gp_oracle as O
    from oracle import reference_path as R
    xs = O.candidates_uniform(SEED_CAND, base_index, max_candidates, lowers, uppers)
    best = (np.inf, -1)
    done = 0
    t0 = time.perf_counter()
    with np.errstate(invalid="ignore", divide="ignore"), all_host_threads():
        while done < max_candidates and time.perf_counter() - t0 < budget_s:
            a = R.ei(gp, xs[done:done + chunk], eta)
            i = int(np.argmin(a))
            if a[i] < best[0]:
                best = (float(a[i]), done + i)
            done += len(a)
    return done, time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the reference's CPU path, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores, model = host_info()
    X, y, gp = make_problem()
    gp.fit(X, y)
    eta = float(np.min(y))
    lo, hi = np.zeros(DIM), np.ones(DIM)
    per_step_budget = args.ref_budget or max(1.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    for w in range(args.warmup):
        time_reference_path(gp, eta, lo, hi, per_step_budget / 4, base_index=w << 17)
    total, elapsed = 0, 0.0
    for s in range(args.steps):
        done, dt = time_reference_path(gp, eta, lo, hi, per_step_budget, base_index=(args.warmup + s) << 17)
        total += done
        elapsed += dt
    value = total / elapsed
    sample = (f"{total // max(1, args.steps)} candidates/step of the step's batch (time-bounded {per_step_budget:.1f} s/step), "
              f"64 candidates per predict(return_cov=True) call")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"Hartmann-6D synthetic, n={N_TRAIN} observations, d={DIM}, 1.0*RBF({LENGTH_SCALE}), "
                               f"alpha={ALPHA_REG}, normalize_y, EI + argmin, candidates U[0,1]^d from (seed, global "
                               f"index); bounded CPU sample of the step's batch",
                   "n": N_TRAIN, "d": DIM, "acquisition": "EI", "host_cpu": model},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- helpers of the B200 arm ---------------------------------------------------------------------------------------------
class Ctx:
    """Process-group facts and the two timing primitives (device events, max over ranks)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # whatever NCCL_DEBUG level the environment asks for goes to a file: stdout carries the one JSON line only
            os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/bopy_b200_nccl.%h.%p.log")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local_rank])
        self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, steps):
        """Device time of `steps` calls of fn(i) between barriers: milliseconds, max over ranks."""
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        last = None
        for i in range(steps):
            last = fn(i)
        ev1.record()
        self.barrier()
        ms = ev0.elapsed_time(ev1)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def fetch_winner(out):
    """(value, index) of a sweep's device-resident arg-min with ONE device-to-host copy."""
    import torch
    both = torch.stack([out["min_idx"][0], out["min_val"].view(torch.int64)[0]]).cpu().numpy()
    return float(both[1:].view(np.float64)[0]), int(both[0])


def sweep_step(ctx, native, xs, eta, index_base, exchange=True):
    """One step: the fused sweep, then the min-loc exchange on the same stream, then one D2H of the winner."""
    from bopy_b200.distributed import all_reduce_minloc_device
    out = native.sweep(xs, acq="ei", eta=eta, want_min=True, index_base=index_base)
    if ctx.world > 1 and exchange:
        all_reduce_minloc_device(out["min_val"], out["min_idx"])
    return fetch_winner(out)


def golden_spot_check(sur, eta_golden, name, dtype):
    """The stream's first candidates ARE the golden fixture's candidate set (same seed, index range [0, m)): compare the
    device path with the frozen outputs of the unmodified reference (tests/golden/<name>.npz, tools/make_golden.py)
    under the tolerance convention of tests/parity_util.py."""
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False))
    xs = sur.native.candidates(g["Xs"])
    out = sur.native.sweep(xs, acq="ei", eta=float(eta_golden), want_mean=True, want_var=True, want_acq=True, want_min=True)
    mean, var, a = (out[k].cpu().numpy() for k in ("mean", "var", "acq"))
    y_std = float(g["y_std"])
    prior = (float(g["amplitude"]) + float(g["noise_level"])) * y_std ** 2
    floor = 0.0 if dtype == "f64" else ({"c3_branin_n256": 1e-1}.get(name, 1e-2))
    mean_bound = 1e-9 * np.abs(g["mean"]) + 1e-9 * y_std
    var_bound = (1e-9 * np.abs(g["var"]) + 1e-11 * prior) if dtype == "f64" else 1e-4 * np.maximum(np.abs(g["var"]), floor * prior)
    me, ve = np.abs(mean - g["mean"]) / mean_bound, np.abs(var - g["var"]) / var_bound
    idx, ref_idx = int(out["min_idx"].item()), int(g["argmin_ei"])
    ref = g["ei"]
    finite = ref[np.isfinite(ref)]
    spread = float(np.ptp(finite)) if finite.size else 1.0
    tie = 1e-9 if dtype == "f64" else 1e-4
    same = idx == ref_idx or (np.isfinite(ref[idx]) and abs(ref[idx] - ref[ref_idx]) <= tie * max(spread, abs(ref[ref_idx])))
    res = {"golden": name, "candidates_checked": int(len(g["Xs"])), "mean_err_over_bound": float(np.nanmax(me)),
           "var_err_over_bound": float(np.nanmax(ve)), "var_err_over_prior": float(np.nanmax(np.abs(var - g["var"])) / prior),
           "argmin": idx, "argmin_reference": ref_idx, "argmin_equal_or_stated_tie": bool(same),
           "nan_values": int(np.isnan(a).sum()), "nan_values_reference": int(np.isnan(ref).sum())}
    res["ok"] = bool(res["mean_err_over_bound"] <= 1.0 and res["var_err_over_bound"] <= 1.0 and same)
    return res


def run_config(ctx, key, dtype, m, steps, peaks, multistart=False):
    """Throughput + roofline fraction + parity spot-check of one BASELINE config on its section 8(d) inputs."""
    torch = ctx.torch
    from bopy_b200 import _native
    from bopy_b200.acquisition import EI
    from bopy_b200.surrogate import B200GPSurrogate
    X, y, gp, lo, hi, golden, device_fit = config_problem(key)
    n, d = X.shape
    sur = B200GPSurrogate(gp, dtype=dtype, device=ctx.dev, device_fit=device_fit)
    t0 = time.perf_counter()
    with all_host_threads():
        sur.fit(X, y)
    torch.cuda.synchronize(ctx.dev)
    fit_s = time.perf_counter() - t0
    ei = EI(sur)
    ei.fit(X, y)
    eta = float(ei._eta)
    native = sur.native
    native.set_latency_path(0)          # every call below is the throughput kernel, whatever its size
    res = {"workload": f"{key}: n={n}, d={d}, {gp.kernel}, alpha={gp.alpha}, normalize_y, EI + argmin, {m} candidates per GPU "
                       f"per step from (seed {SEED_CAND}, global index) in the config's box",
           "n": n, "d": d, "dtype": dtype, "candidates_per_gpu": m, "fit": "device" if sur.fitted_on_device else "host (sklearn)",
           "fit_s": fit_s}
    if ctx.rank == 0:
        try:
            res["parity"] = golden_spot_check(sur, eta, golden, dtype)
        except Exception as exc:          # a failed spot-check is reported, it does not take the bench line down
            res["parity"] = {"ok": False, "error": repr(exc)}
    bases = [(b * ctx.world + ctx.rank) * m for b in range(2)]
    bufs = [_native.candidates_uniform(SEED_CAND, bases[b], m, lo, hi, device=ctx.dev) for b in range(2)]
    sweep_step(ctx, native, bufs[0][: min(m, 1 << 15)], eta, bases[0])                      # warm-up: kernels loaded
    # ... and full-size warm-up steps where a step is milliseconds (C1, C3: three; C4 fp32: one; a C5 step is 17 s): a five-step
    # region of 0.5 ms steps once caught a 12 ms one-off of the first sharded exchange
    warm = 3 if key in ("C1", "C3") else (1 if key == "C4" else 0)
    for w in range(warm):
        sweep_step(ctx, native, bufs[w % 2], eta, bases[w % 2])
    ms, winner = ctx.timed(lambda i: sweep_step(ctx, native, bufs[i % 2], eta, bases[i % 2]), steps)
    kms, _ = ctx.timed(lambda i: native.sweep(bufs[i % 2], acq="ei", eta=eta, want_min=True, index_base=bases[i % 2]), steps)
    kernel_ms = kms / steps
    F = flops_per_candidate(n, d)
    peak = max(peaks["fp64_fma"], peaks["fp64_mma"]) if dtype == "f64" else peaks["tf32_tcgen05"] / 3.0
    achieved = F * m / (kernel_ms * 1e-3) / 1e12
    res.update(value=ctx.world * m * steps / (ms * 1e-3), unit=UNIT, steps=steps, warmup=warm, ms_per_step=ms / steps, kernel_ms=kernel_ms,
               argmin={"index": winner[1], "value": winner[0]},
               roofline={"bound": "tensor", "pipe": "FP64 DMMA" if dtype == "f64" else "tcgen05.mma kind::tf32, 3 MMAs per product",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "flops_per_candidate": F, "traffic": ncu_traffic(f"{key}_{dtype}")})
    del bufs
    if multistart:
        from bopy_b200.bounds import Bound, Bounds
        from bopy_b200.optimizer import MultiStartOptimizer
        native.set_latency_path(4096)
        opt = MultiStartOptimizer(ei, Bounds([Bound(float(a), float(b)) for a, b in zip(lo, hi)]), n_starts=1024,
                                  n_candidates=1 << 20, seed=7, method="gradient", iterations=40,
                                  distributed=ctx.world > 1)
        ctx.barrier()
        t0 = time.perf_counter()
        r = opt.optimize()
        ctx.barrier()
        res["multistart_1024"] = {"seconds": time.perf_counter() - t0, "f_min": float(r.f_min[0]),
                                  "x_min_head": [float(v) for v in r.x_min[0][:4]],
                                  "what": "1024 stratified starts from a 2^20-candidate segmented sweep, 40 projected-gradient "
                                          "iterations in one native call, starts sharded over the ranks, one min-loc exchange; "
                                          "f_min / x_min must be identical for every GPU count"}
    native.close()
    return res


def run_strong(ctx, key, total, steps):
    """Strong scaling: a fixed total of candidates split over the ranks; the collective's share of a step."""
    from bopy_b200 import _native
    from bopy_b200.acquisition import EI
    from bopy_b200.distributed import shard_range
    from bopy_b200.surrogate import B200GPSurrogate
    X, y, gp, lo, hi, _, _ = config_problem(key)
    sur = B200GPSurrogate(gp, dtype="f64", device=ctx.dev, device_fit=True)
    sur.fit(X, y)
    ei = EI(sur)
    ei.fit(X, y)
    eta = float(ei._eta)
    native = sur.native
    native.set_latency_path(0)
    start, stop = shard_range(total, ctx.rank, ctx.world)
    xs = _native.candidates_uniform(SEED_CAND, start, stop - start, lo, hi, device=ctx.dev)
    for _ in range(3 if key == "C3" else 1):                 # warm-up (a C3 step is a millisecond, a C4 step two seconds)
        sweep_step(ctx, native, xs, eta, start)
    ms, winner = ctx.timed(lambda i: sweep_step(ctx, native, xs, eta, start), steps)
    ms_nox, _ = ctx.timed(lambda i: sweep_step(ctx, native, xs, eta, start, exchange=False), steps)
    native.close()
    return {"config": key, "candidates_total": total, "value": total * steps / (ms * 1e-3), "unit": UNIT, "steps": steps,
            "ms_per_step": ms / steps, "ms_per_step_without_exchange": ms_nox / steps,
            "collective_share": max(0.0, 1.0 - ms_nox / ms) if ctx.world > 1 else 0.0,
            "argmin": {"index": winner[1], "value": winner[0]}}


def run_b200(args):
    import torch

    from bopy_b200 import _native
    from bopy_b200.acquisition import EI
    from bopy_b200.surrogate import B200GPSurrogate

    ctx = Ctx()
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    n, d, m = args.n, args.d, args.candidates
    X, y, gp = make_problem(n, d)
    sur = B200GPSurrogate(gp, dtype=args.dtype, device=dev)
    sur.fit(X, y)                                   # fixed hyper-parameters: Gram + Cholesky + alpha on the device
    ei = EI(sur)
    ei.fit(X, y)
    eta = float(ei._eta)
    lo, hi = np.zeros(d), np.ones(d)
    native = sur.native

    # two resident candidate buffers with disjoint global index ranges, rotated between steps
    nbuf = 2
    bases = [(b * world + rank) * m for b in range(nbuf)]
    bufs = [_native.candidates_uniform(SEED_CAND, bases[b], m, lo, hi, device=dev) for b in range(nbuf)]

    for i in range(args.warmup):
        sweep_step(ctx, native, bufs[i % nbuf], eta, bases[i % nbuf])
    with ClockSampler(ctx.local_rank) as clocks:
        ms, result = ctx.timed(lambda i: sweep_step(ctx, native, bufs[i % nbuf], eta, bases[i % nbuf]), args.steps)
    value = world * m * args.steps / (ms * 1e-3)

    # kernel-only timing for the roofline: back-to-back sweeps, no host round trip in between
    kms, _ = ctx.timed(lambda i: native.sweep(bufs[i % nbuf], acq="ei", eta=eta, want_min=True, index_base=bases[i % nbuf]),
                       args.steps)
    kernel_ms = kms / args.steps

    # end to end through the public API: pinned host candidates -> H2D -> sweep -> exchange -> D2H result
    host = [bufs[b].cpu().pin_memory() for b in range(nbuf)]

    def e2e_step(i):
        from bopy_b200.distributed import all_reduce_minloc_device
        idx_d, val_d = ei.argmin(host[i % nbuf], index_base=bases[i % nbuf], on_device=True)
        if world > 1:
            all_reduce_minloc_device(val_d, idx_d)
        return fetch_winner({"min_idx": idx_d, "min_val": val_d})

    e2e_step(0)
    t0 = time.perf_counter()
    e2e_ms, e2e_result = ctx.timed(e2e_step, args.steps)
    e2e_wall = time.perf_counter() - t0
    e2e_value = world * m * args.steps / (e2e_ms * 1e-3)

    peaks_all = {k: _native.measure_peak(k) for k in ("fp64_fma", "fp32_fma", "fp64_mma", "tf32_mma_sync", "tf32_tcgen05")}

    side, strong = {}, {}
    if not args.headline_only:
        plan = [("C1", "f64", 1 << 22, 50, False), ("C3", "f64", 1 << 20, 30, False), ("C4_f32", "f32", CAND_PER_GPU, 3, False),
                ("C5", "f64", args.c5_candidates, 1, True)]
        for name, dtype, mm, steps, ms_flag in plan:
            try:
                side[name] = run_config(ctx, name.split("_")[0], dtype, mm, steps, peaks_all, multistart=ms_flag)
            except Exception as exc:          # reported, never fatal for the headline line
                side[name] = {"error": repr(exc)}
            torch.cuda.empty_cache()
        for name, total, steps in (("C4", 1 << 24, 2), ("C3", 1 << 20, 20)):
            try:
                strong[name] = run_strong(ctx, name, total, steps)
            except Exception as exc:
                strong[name] = {"error": repr(exc)}

    if rank != 0:
        ctx.close()
        return

    # roofline of the dominant kernel: the FP64 pipe (DFMA and DMMA share it on B200; the larger of the two measured
    # rates is the denominator) or, for the fp32 mode, the tcgen05 TF32 rate / 3 (three MMAs per fp32-grade product)
    f32_engine = os.environ.get("BOPY_B200_F32_ENGINE", "tcgen05")
    if args.dtype == "f64":
        peak = max(peaks_all["fp64_fma"], peaks_all["fp64_mma"])
    elif f32_engine == "fma":
        peak = peaks_all["fp32_fma"]
    elif f32_engine == "mma_sync":
        peak = peaks_all["tf32_mma_sync"] / 3.0
    else:
        peak = peaks_all["tf32_tcgen05"] / 3.0
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    group = native.set_group_mode(-2)      # (thread blocks per candidate tile, slots, lead) in force; 0 = one tile per block
    nominal = sm_count * (64 if args.dtype == "f64" else 128) * 2 * 1.965e9 / 1e12
    F = flops_per_candidate(n, d)
    achieved = F * m / (kernel_ms * 1e-3) / 1e12
    measured = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            measured = json.load(f)
    except OSError:
        pass
    hbm_peak = measured.get("hbm_gbs", 6650.0)
    algo_bytes = m * d * 8 + 16
    default_shape = (n, d, m) == (N_TRAIN, DIM, CAND_PER_GPU)
    roofline = {
        "bound": "fp32_fma" if (args.dtype == "f32" and f32_engine == "fma") else "tensor",
        "pipe": ("FP64 tensor sub-pipe (mma.sync.m8n8k4.f64 = SASS DMMA; tcgen05 has no f64 kind)" if args.dtype == "f64"
                 else "tcgen05.mma kind::tf32 (SASS UTCHMMA), accumulators in TMEM, 3xTF32 split: peak = measured rate / 3; "
                      "K*, diagonal solve and sum v^2 on the FP64 pipe"),
        "kernel": ("sweep_group_kernel" if group[0] >= 2 else "sweep_kernel") if args.dtype == "f64" else "sweep_tc_kernel",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "peak_source": "bopy_measure_peak: register-resident DFMA / DMMA loops and a shared-memory-operand tcgen05 loop measured "
                       "live on this GPU (MEASURED_PEAKS.json holds only HBM and bf16 peaks; the path is FP-pipe bound, SURVEY.md 8d)",
        "peak_nominal": nominal, "frac_of_nominal": achieved / nominal,
        "flops_per_candidate": F, "candidates_per_launch": m, "kernel_ms": kernel_ms,
        "traffic": ncu_traffic(f"C4_{args.dtype}") if default_shape else None,
        "traffic_source": "profiles/r02/ncu_traffic.json (ncu --set full on this build, same launch shape)",
        "hbm": {"algorithmic_bytes_per_launch": algo_bytes, "achieved_gbs": algo_bytes / (kernel_ms * 1e-3) / 1e9,
                "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if measured else "fallback"},
        "measured_peaks_tflops": peaks_all,
    }

    # CPU baseline on this host: the reference's call sequence, bounded sample
    cores, model = host_info()
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # the surrogate above was fitted on the device (its sklearn object holds no L_): the CPU arm gets its own
        # host-fitted scikit-learn model, exactly what the reference's ScipyGPSurrogate.fit would build
        from sklearn.base import clone
        gp_host = clone(gp)
        with all_host_threads():
            gp_host.fit(X, y)
        gp_host.predict(X[:8], return_cov=True)
        done, dt = time_reference_path(gp_host, eta, lo, hi, budget_s=args.cpu_budget)
        cpu = {"value": done / dt, "unit": UNIT, "cores": cores, "kind": "port", "host_cpu": model,
               "sample": f"first {done} candidates of the step's batch ({dt:.1f} s), bopy's call sequence on "
                         f"sklearn/scipy: predict(return_cov=True) on 64 candidates per call -> np.diag -> norm EI"}

    # the reference's own calling pattern: one point per acquisition call (DIRECT, bopy/optimizer.py:96-97)
    probe = None
    if world == 1:
        x1 = np.ascontiguousarray(np.random.default_rng(7).random((1, d)))

        def per_call_ms(warm):
            for _ in range(warm):
                ei(x1)
            t0 = time.perf_counter()
            for _ in range(200):
                ei(x1)
            return 1e3 * (time.perf_counter() - t0) / 200
        if args.dtype == "f64":
            native.set_inverse_path(0)
            chained = per_call_ms(20)
            served = native.set_inverse_path(-1)          # the library default: W = L^-1 from the 16th probe of a state on
            probe = {"ms_per_call": per_call_ms(40), "ms_per_call_chained": chained,
                     "what": "EI(x) through the public API, x a (1, d) numpy array, numpy out, library defaults: after 16 "
                             "probes of one state a call is ONE matrix-vector product with W = L^-1 (probe_inv_kernel"
                             + (")" if served >= 1 else " -- not available on this handle)")
                             + "; ms_per_call_chained: the same call on probe_kernel (n/128 dependent hops)"}
        else:
            probe = {"ms_per_call": per_call_ms(20),
                     "what": "EI(x) through the public API, x a (1, d) numpy array, numpy out (fp32 handles have no "
                             "latency path: sweep kernel)"}

    # opt-in branch and bound for the same arg-min (not part of `value`: most candidates skip the full posterior)
    pruned = None
    if world == 1:
        pev0, pev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        native.argmin_pruned(bufs[0], "ei", eta=eta, index_base=bases[0])
        torch.cuda.synchronize(dev)
        pev0.record()
        for i in range(args.steps):
            pv, pi, pstats = native.argmin_pruned(bufs[i % nbuf], "ei", eta=eta, index_base=bases[i % nbuf])
        pev1.record()
        torch.cuda.synchronize(dev)
        plain = native.sweep(bufs[(args.steps - 1) % nbuf], acq="ei", eta=eta, want_min=True,
                             index_base=bases[(args.steps - 1) % nbuf])
        pruned = {"ms_per_step": pev0.elapsed_time(pev1) / args.steps, "candidates": pstats["candidates"],
                  "fully_evaluated": pstats["swept"],
                  "same_argmin_as_plain_sweep": bool(int(pi.item()) == int(plain["min_idx"].item())
                                                     and float(pv.item()) == float(plain["min_val"].item())),
                  "what": "bopy_acq_argmin_pruned: mean-only lower bounds + incumbent from a strided sample, full sweep "
                          "over the survivors only; NOT counted in value / e2e"}

    info = native.launch_info(m)
    launches = info["launches"] + (2 if world > 1 else 0)     # + record pack and gathered-min-loc kernels of the exchange
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"{'Hartmann-6D' if d == 6 else 'sin-sum'} synthetic, n={n} observations, d={d}, 1.0*RBF({LENGTH_SCALE}), "
                               f"alpha={ALPHA_REG}, normalize_y, EI + argmin, {m} candidates per GPU per step "
                               f"({world * m} per step in total), candidates U[0,1]^d from (seed, global index)",
                   "n": n, "d": d, "candidates_per_gpu": m, "acquisition": "EI",
                   "l2": f"{nbuf} candidate buffers rotated between steps ({nbuf * m * d * 8 / 1e6:.0f} MB: larger than the 126 MB "
                         f"L2) and {info['workspace_bytes'] / 1e6:.0f} MB of solve workspace in flight"
                         + (f" (group mode: {group[0]} thread blocks per candidate tile, {group[1]} slots per group, so that "
                            f"the workspace stays L2-resident)" if group[0] >= 2 else ""),
                   "group_mode": {"group_size": group[0], "slots": group[1], "lead": group[2]},
                   "parallelism": f"candidates sharded over {world} GPU(s), state replicated, one min-loc exchange "
                                  f"(ncclAllGather of 16-byte records + one-warp kernel, on the sweep's stream)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": m * d * 8, "d2h_bytes_per_step": 16,
                "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": 1e3 * e2e_wall / args.steps},
        "gpu_launches": args.steps * launches,
        "grid": info["grid"],
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks.summary(),
        "argmin": {"index": result[1], "value": result[0], "e2e_index": e2e_result[1]},
        "single_point_probe": probe,
        "argmin_branch_and_bound": pruned,
        "configs": side,
        "strong_scaling": strong,
    }
    print(json.dumps(line), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--n", type=int, default=N_TRAIN)
    ap.add_argument("--d", type=int, default=DIM)
    ap.add_argument("--candidates", type=int, default=CAND_PER_GPU, help="candidates per GPU per step")
    ap.add_argument("--c5-candidates", type=int, default=1 << 23, help="C5 candidates per GPU (2^23 = 64M on 8 GPUs)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-budget", type=float, default=None, help="--impl reference: seconds of CPU work per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip the `configs` / `strong_scaling` side runs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
