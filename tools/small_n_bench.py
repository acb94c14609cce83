import sys, time, json
sys.path.insert(0, "/root/repo")
import bench, torch
ctx = bench.Ctx()
from bopy_b200 import _native
peaks = {k: _native.measure_peak(k) for k in ("fp64_fma", "fp64_mma", "tf32_tcgen05")}
for key, m in (("C1", 1 << 22), ("C3", 1 << 20)):
    r = bench.run_config(ctx, key, "f64", m, 5, peaks)
    print(key, "value %.4g kernel_ms %.3f frac %.4f parity %s" % (r["value"], r["kernel_ms"], r["roofline"]["frac"], r["parity"]["ok"]), r["parity"])
