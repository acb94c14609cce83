#!/usr/bin/env python
"""Benchmark of the hot path: fused GP posterior + EI + argmin over a candidate batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[3] ("Hartmann-6D synthetic, n=2048 observations, 16M candidates sharded over
8xB200"): n=2048 training points in d=6, kernel 1.0*RBF(0.3), alpha=1e-6, normalize_y; 2^21 candidates per GPU
per step (weak scaling: 16M at 8 GPUs), generated on the device from (seed, global index).  A step is one fused
sweep of the rank's candidates (K* tile -> blocked triangular solve -> variance -> EI -> argmin) plus the single
min-loc exchange between ranks.  One JSON line is printed by rank 0.

value   : candidates/s with the candidates already resident in HBM.
e2e     : the same through the public API with HOST candidates in pinned memory: H2D copy, sweep, D2H of the
          (index, value) result inside the timed region.
roofline: the sweep kernel is bound by the FP64 tensor sub-pipe (DMMA) (F(n,d) = n^2 + n(3d+5)
          flops per candidate against 48 B of HBM input); peak = max(DFMA, DMMA) rate measured live by
          bopy_measure_peak (MEASURED_PEAKS.json has no FP64 figure).
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


# dram__bytes_read.sum + dram__bytes_write.sum of ONE sweep_kernel launch from `ncu --set full` at the default workload
# (profiles/r01/ncu_sweep_v3_f64_dmma_fullsize_summary.json): 191.7 GB read + 32.2 GB written, all of it re-reads /
# writes of the per-CTA V workspace (L2 hit rate 53 %).  Other workloads: not captured -> null.
NCU_TRAFFIC_BYTES = {(2048, 6, "f64", 1 << 21): 223.9e9}


def flops_per_candidate(n, d):
    return n * n + n * (3 * d + 5)


def make_problem(n=N_TRAIN, d=DIM):
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


def run_b200(args):
    import torch
    import torch.distributed as dist

    from bopy_b200 import _native
    from bopy_b200.acquisition import EI
    from bopy_b200.distributed import all_reduce_minloc
    from bopy_b200.surrogate import B200GPSurrogate

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # whatever NCCL_DEBUG level the environment asks for goes to a file: stdout carries the one JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/bopy_b200_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)

    n, d, m = args.n, args.d, args.candidates
    X, y, gp = make_problem(n, d)
    sur = B200GPSurrogate(gp, dtype=args.dtype, device=dev)
    sur.fit(X, y)                                   # host sklearn fit (not on the hot path) + state upload
    ei = EI(sur)
    ei.fit(X, y)
    eta = float(ei._eta)
    lo, hi = np.zeros(d), np.ones(d)
    native = sur.native

    # two resident candidate buffers with disjoint global index ranges, rotated between steps
    nbuf = 2
    bases = [(b * world + rank) * m for b in range(nbuf)]
    bufs = [_native.candidates_uniform(SEED_CAND, bases[b], m, lo, hi, device=dev) for b in range(nbuf)]

    def step(i):
        out = native.sweep(bufs[i % nbuf], acq="ei", eta=eta, want_min=True, index_base=bases[i % nbuf])
        val, idx = float(out["min_val"].item()), int(out["min_idx"].item())
        if world > 1:
            val, idx = all_reduce_minloc(val, idx, device=dev)
        return val, idx

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for i in range(args.steps):
            result = step(i)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * m * args.steps / (ms * 1e-3)

    # kernel-only timing for the roofline: back-to-back sweeps, no host round trip in between
    kev0, kev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    kev0.record()
    for i in range(args.steps):
        native.sweep(bufs[i % nbuf], acq="ei", eta=eta, want_min=True, index_base=bases[i % nbuf])
    kev1.record()
    torch.cuda.synchronize(dev)
    kernel_ms = kev0.elapsed_time(kev1) / args.steps

    # end to end through the public API: pinned host candidates -> H2D -> sweep -> D2H result
    host = [bufs[b].cpu().pin_memory() for b in range(nbuf)]
    eev0, eev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def e2e_step(i):
        idx, val = ei.argmin(host[i % nbuf], index_base=bases[i % nbuf])
        if world > 1:
            val, idx = all_reduce_minloc(val, idx, device=dev)
        return val, idx

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    eev0.record()
    for i in range(args.steps):
        e2e_result = e2e_step(i)
    eev1.record()
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = max(eev0.elapsed_time(eev1), 0.0)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * m * args.steps / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (sweep_kernel): the FP64 pipe (DFMA and DMMA share it on B200; the larger
    # of the two measured rates is the denominator) or, for the fp32 mode, the FP32 FMA pipe
    peaks_all = {k: _native.measure_peak(k) for k in ("fp64_fma", "fp32_fma", "fp64_mma", "tf32_mma_sync")}
    f32_fma_engine = os.environ.get("BOPY_B200_F32_ENGINE") == "fma"
    if args.dtype == "f64":
        peak = max(peaks_all["fp64_fma"], peaks_all["fp64_mma"])
    elif f32_fma_engine:
        peak = peaks_all["fp32_fma"]
    else:
        peak = peaks_all["tf32_mma_sync"] / 3.0      # 3xTF32: three tensor instructions per fp32-grade product
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
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
    roofline = {
        "bound": "fp32_fma" if (args.dtype == "f32" and f32_fma_engine) else "tensor",
        "pipe": ("FP64 tensor sub-pipe (mma.sync.m8n8k4.f64 = SASS DMMA; tcgen05 has no f64 kind)" if args.dtype == "f64"
                 else ("FP32 FMA pipe for the off-diagonal GEMM, FP64 DMMA for the diagonal solve" if f32_fma_engine else
                       "TF32 tensor pipe, 3xTF32 split (mma.sync.m16n8k8 = SASS HMMA.1688.F32.TF32; peak = measured "
                       "rate / 3) for the off-diagonal GEMM, FP64 DMMA for the kernel tile's diagonal solve")),
        "kernel": "sweep_kernel",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "peak_source": "bopy_measure_peak: register-resident DFMA / DMMA / FFMA loops measured live on this GPU "
                       "(MEASURED_PEAKS.json holds only HBM and bf16 peaks; the path is FP-pipe bound, SURVEY.md 8d)",
        "peak_nominal": nominal, "frac_of_nominal": achieved / nominal,
        "flops_per_candidate": F, "candidates_per_launch": m, "kernel_ms": kernel_ms,
        "traffic": NCU_TRAFFIC_BYTES.get((n, d, args.dtype, m)),
        "traffic_note": "ncu dram bytes per launch; V workspace re-reads (0.78 TB/s, 12 % of HBM peak), kernel is FP64-pipe bound",
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
        for _ in range(20):
            ei(x1)
        t0 = time.perf_counter()
        for _ in range(200):
            ei(x1)
        probe = {"ms_per_call": 1e3 * (time.perf_counter() - t0) / 200,
                 "what": "EI(x) through the public API, x a (1, d) numpy array, numpy out ("
                         + ("latency path: probe_kernel)" if args.dtype == "f64" else "fp32 handles have no latency path: sweep_kernel)")}

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
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"{'Hartmann-6D' if d == 6 else 'sin-sum'} synthetic, n={n} observations, d={d}, 1.0*RBF({LENGTH_SCALE}), "
                               f"alpha={ALPHA_REG}, normalize_y, EI + argmin, {m} candidates per GPU per step "
                               f"({world * m} per step in total), candidates U[0,1]^d from (seed, global index)",
                   "n": n, "d": d, "candidates_per_gpu": m, "acquisition": "EI",
                   "l2": f"{nbuf} candidate buffers rotated between steps ({nbuf * m * d * 8 / 1e6:.0f} MB) and a "
                         f"{info['workspace_bytes'] / 1e6:.0f} MB solve workspace: larger than the 126 MB L2",
                   "parallelism": f"candidates sharded over {world} GPU(s), state replicated, one min-loc all-gather"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": m * d * 8, "d2h_bytes_per_step": 16,
                "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": 1e3 * e2e_wall / args.steps},
        "gpu_launches": args.steps * info["launches"],
        "grid": info["grid"],
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks.summary(),
        "argmin": {"index": result[1], "value": result[0], "e2e_index": e2e_result[1]},
        "single_point_probe": probe,
        "argmin_branch_and_bound": pruned,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-budget", type=float, default=None, help="--impl reference: seconds of CPU work per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
