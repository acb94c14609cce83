"""Where a hop of the latency path's chain goes: %globaltimer stamps of one single-candidate probe_kernel launch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bopy_b200.surrogate import B200GPSurrogate  # noqa: E402


def main():
    n, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 6)
    X, y, gp = bench.make_problem(n, d)
    sur = B200GPSurrogate(gp)
    sur.fit(X, y)
    native = sur.native
    xs = native.candidates(np.random.default_rng(0).random((1, d)))
    for _ in range(5):
        native.sweep(xs, acq="ei", eta=float(y.min()), want_acq=True)
    native.probe_trace(arm=True)
    native.sweep(xs, acq="ei", eta=float(y.min()), want_acq=True)
    t = native.probe_trace()
    t0 = t[:, 0].min()
    # stamps: 0 start, 1 K* done, 3 early blocks done, 6 diagonal solve of P_I done, 2 flag of V_{I-1} seen,
    #         4 hop product done, 5 V_I released
    order = [0, 1, 3, 6, 2, 4, 5]
    names = ["start", "K* done", "early GEMM done", "diag done", "last flag seen", "hop GEMM done", "V released"]
    print("block row | " + " | ".join(f"{s:>15s}" for s in names) + "   (us since the first CTA started)")
    for I in range(len(t)):
        print(f"{I:9d} | " + " | ".join(f"{(t[I, k] - t0) / 1e3:15.2f}" if t[I, k] else f"{'-':>15s}" for k in order))
    pub = t[:-1, 5]
    hops = np.diff(pub) / 1e3
    print("hop (released[I] - released[I-1]) us: mean %.2f min %.2f max %.2f" % (hops.mean(), hops.min(), hops.max()))
    I = np.arange(1, len(t) - 1)
    print("  flag seen after the previous release: %.2f us" % np.mean((t[I, 2] - t[I - 1, 5]) / 1e3))
    print("  V load + barrier + M_I product: %.2f us" % np.mean((t[I, 4] - t[I, 2]) / 1e3))
    print("  release (stores, barrier, fence, flag): %.2f us" % np.mean((t[I, 5] - t[I, 4]) / 1e3))
    print("kernel: %.2f us from the first CTA's start to the last stamp" % ((t.max() - t0) / 1e3))


if __name__ == "__main__":
    main()
