"""C1 / C3 through bench.run_config (development aid, also the command the round-2 ncu captures of the small-n kernels profile)."""
import sys, time, json
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
ctx = bench.Ctx()
from bopy_b200 import _native
peaks = {k: _native.measure_peak(k) for k in ("fp64_fma", "fp64_mma", "tf32_tcgen05")}
for key, m in (("C1", 1 << 22), ("C3", 1 << 20)):
    r = bench.run_config(ctx, key, "f64", m, 5, peaks)
    print(key, "value %.4g kernel_ms %.3f frac %.4f parity %s" % (r["value"], r["kernel_ms"], r["roofline"]["frac"], r["parity"]["ok"]), r["parity"])
