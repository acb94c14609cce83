"""Diagnostic sweep on a GPU box: error statistics of the CUDA path vs oracle/golden for every fixture,
both dtypes, plus the measured FMA/MMA peaks.  Prints, never asserts (pytest -m gpu does the asserting)."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import golden_names, golden_state  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402
from parity_util import check_mean, check_var, prior_var  # noqa: E402
from test_gpu_parity import native_for  # noqa: E402


def main():
    import torch
    from bopy_b200 import _native
    print("device:", torch.cuda.get_device_name(0), "SMs", torch.cuda.get_device_properties(0).multi_processor_count,
          "L2", torch.cuda.get_device_properties(0).L2_cache_size)
    for what in ("fp64_fma", "fp32_fma", "fp64_mma"):
        try:
            print(f"peak {what}: {_native.measure_peak(what):.2f} TFLOP/s")
        except Exception as e:
            print("peak", what, "failed:", e)
    names = [n for n in golden_names() if not (len(sys.argv) > 1 and sys.argv[1] not in n)]
    for name in names:
        g, st = golden_state(name)
        for dtype in ("f64", "f32"):
            try:
                t0 = time.time()
                gp = native_for(st, dtype)
                t_set = time.time() - t0
                xs = gp.candidates(g["Xs"])
                out = gp.sweep(xs, acq="ei", eta=float(g["eta"]), want_mean=True, want_var=True, want_acq=True, want_min=True)
                torch.cuda.synchronize()
                mean, var, a = (out[k].cpu().numpy() for k in ("mean", "var", "acq"))
                em, bm = check_mean(mean, g["mean"], st, dtype)
                ev, bv = check_var(var, g["var"], st, dtype, name)
                pv = prior_var(st)
                rel = np.abs(var - g["var"]) / np.maximum(np.abs(g["var"]), 1e-300)
                print(f"{name:32s} {dtype} set_state {t_set*1e3:7.1f} ms | mean err/bound {np.max(em/bm):9.3g} (max abs {em.max():.2e}) "
                      f"| var err/bound {np.max(ev/bv):9.3g} (max abs/prior {ev.max()/pv:.2e}, max rel {rel.max():.2e}) "
                      f"| argmin {int(out['min_idx'].item())} ref {int(g['argmin_ei'])} np {int(np.argmin(a))} "
                      f"| nan {int(np.isnan(a).sum())}/{int(np.isnan(g['ei']).sum())}")
                if dtype == "f32":
                    # the pure-relative error of the variance (north star: 1e-4 in fp32) as a histogram, and where it sits
                    # against the posterior variance itself: relative error is only large where var << prior
                    ok = np.isfinite(var) & np.isfinite(g["var"]) & (np.abs(g["var"]) > 0)
                    edges = [0, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, np.inf]
                    hist = np.histogram(rel[ok], bins=edges)[0]
                    labels = ["<1e-7", "<1e-6", "<1e-5", "<1e-4", "<1e-3", "<1e-2", "<1e-1", ">=1e-1"]
                    print("    relative |dvar|/|var| histogram: " + "  ".join(f"{lb}:{int(c)}" for lb, c in zip(labels, hist)))
                    over = ok & (rel > 1e-4)
                    q = np.quantile(np.abs(var - g["var"])[ok] / pv, [0.5, 0.9, 0.99, 1.0])
                    print(f"    |dvar|/prior quantiles 50/90/99/100 %: {q[0]:.2e} {q[1]:.2e} {q[2]:.2e} {q[3]:.2e}; "
                          f"candidates with relative error > 1e-4: {int(over.sum())} of {int(ok.sum())}"
                          + (f", all with var/prior <= {float(np.max(np.abs(g['var'])[over]) / pv):.2e}" if over.any() else ""))
                gp.close()
            except Exception:
                print(name, dtype, "FAILED")
                traceback.print_exc()


if __name__ == "__main__":
    main()
