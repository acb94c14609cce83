"""Per-call latency for the reference's own calling pattern (few candidates per call: DIRECT probes with m = 1,
bopy/optimizer.py:96-97; plotting grids with m = 100..2500), beside the reference path on the host.

Columns: EI(x) through the public API (numpy in, numpy out) with the latency path (probe_kernel) and with the
throughput path (sweep_kernel), the device time of the fused launch alone (CUDA events, candidates resident), and
bopy's call sequence on scikit-learn / scipy (64 candidates per predict(return_cov=True) call)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bopy_b200.acquisition import EI  # noqa: E402
from bopy_b200.surrogate import B200GPSurrogate  # noqa: E402
from oracle import reference_path as R  # noqa: E402


def device_ms(native, xs, eta, reps):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    native.sweep(xs, acq="ei", eta=eta, want_acq=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        native.sweep(xs, acq="ei", eta=eta, want_acq=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def api_ms(ei, xs, reps):
    import torch
    ei(xs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        ei(xs)
    return 1e3 * (time.perf_counter() - t0) / reps


def main():
    shapes = [(10, 1), (256, 2), (2048, 6)] + ([(8192, 20)] if "--large" in sys.argv else [])
    ms = (1, 8, 64, 128, 512, 1024, 4096, 8192)
    rows = []
    for n, d in shapes:
        X, y, gp = bench.make_problem(n, d)
        sur = B200GPSurrogate(gp)
        sur.fit(X, y)
        ei = EI(sur)
        ei.fit(X, y)
        host = bench.make_problem(n, d)[2].fit(X, y)
        eta = float(y.min())
        rng = np.random.default_rng(0)
        import torch
        sur.native.set_latency_path(1 << 30)
        if sur.native.set_inverse_path(1) >= 1:
            x1 = rng.random((1, d))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ei(x1)
            print(f"n={n:5d} d={d:2d}: first probe in mode 1 (builds W = L^-1 by blocked TRTRI, then probes): "
                  f"{1e3 * (time.perf_counter() - t0):.3f} ms", flush=True)
        x1 = rng.random((1, d))
        for _ in range(20):
            sur.predict(x1)
        t0 = time.perf_counter()
        for _ in range(100):
            sur.predict(x1)
        t_pred = 1e3 * (time.perf_counter() - t0) / 100
        with bench.all_host_threads():
            host.predict(x1, return_cov=True)
            t0 = time.perf_counter()
            for _ in range(20):
                host.predict(x1, return_cov=True)
            t_pred_ref = 1e3 * (time.perf_counter() - t0) / 20
        print(f"n={n:5d} d={d:2d}: Surrogate.predict(x) of ONE point (mean, 1 x 1 covariance): {t_pred:.3f} ms/call   "
              f"scikit-learn predict(return_cov=True): {t_pred_ref:.3f}", flush=True)
        x100 = rng.random((100, d))
        for _ in range(5):
            sur.predict(x100)
        t0 = time.perf_counter()
        for _ in range(30):
            sur.predict(x100)
        t_pred = 1e3 * (time.perf_counter() - t0) / 30
        with bench.all_host_threads():
            host.predict(x100, return_cov=True)
            t0 = time.perf_counter()
            for _ in range(10):
                host.predict(x100, return_cov=True)
            t_pred_ref = 1e3 * (time.perf_counter() - t0) / 10
        print(f"n={n:5d} d={d:2d}: Surrogate.predict(x) of 100 points (mean, 100 x 100 covariance): {t_pred:.3f} ms/call   "
              f"scikit-learn predict(return_cov=True): {t_pred_ref:.3f}", flush=True)
        for m in ms:
            xs = rng.random((m, d))
            xd = sur.native.candidates(xs)
            reps = 50 if m <= 1024 else 10
            sur.native.set_latency_path(1 << 30)
            t_inv_api = t_inv_dev = float("nan")
            if m <= sur.native.set_inverse_path(1):     # W = L^-1 (probe_inv_kernel), built at the first call
                t_inv_api, t_inv_dev = api_ms(ei, xs, reps), device_ms(sur.native, xd, eta, reps)
            sur.native.set_inverse_path(0)
            t_lat_api, t_lat_dev = api_ms(ei, xs, reps), device_ms(sur.native, xd, eta, reps)
            sur.native.set_latency_path(0)
            t_swp_api, t_swp_dev = api_ms(ei, xs, reps), device_ms(sur.native, xd, eta, reps)
            sur.native.set_latency_path(4096)
            t_ref = float("nan")
            if m <= 1024:
                with bench.all_host_threads():
                    R.ei(host, xs[:min(m, 256)], eta)
                    t0 = time.perf_counter()
                    for _ in range(3):
                        for s in range(0, m, 64):
                            R.ei(host, xs[s:s + 64], eta)
                    t_ref = 1e3 * (time.perf_counter() - t0) / 3
            rows.append(dict(n=n, d=d, m=m, inverse_api_ms=t_inv_api, inverse_dev_ms=t_inv_dev, latency_api_ms=t_lat_api, latency_dev_ms=t_lat_dev, sweep_api_ms=t_swp_api,
                             sweep_dev_ms=t_swp_dev, reference_ms=t_ref))
            print(f"n={n:5d} d={d:2d} m={m:5d}: inverse path {t_inv_api:8.3f} ms/call (device {t_inv_dev:7.3f})   "
                  f"latency path {t_lat_api:8.3f} (device {t_lat_dev:7.3f})   "
                  f"throughput path {t_swp_api:8.3f} (device {t_swp_dev:7.3f})   reference {t_ref:8.3f}", flush=True)
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
