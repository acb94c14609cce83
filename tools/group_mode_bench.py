"""Group mode of the fp64 sweep (csrc/sweep_group_kernel.cuh) against the one-tile-per-CTA kernel: bit-identity of every
output and the throughput per (group size, slots, lead) setting.

    python tools/group_mode_bench.py [C4|C5|C3] [m_log2] [settings ...]      settings: G:S:lead, e.g. 8:2:2 4:2:4
Development aid; also the command the round-2 ncu capture of sweep_group_kernel profiles (one setting, --once).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from bopy_b200 import _native
from bopy_b200.acquisition import EI
from bopy_b200.surrogate import B200GPSurrogate


def make(key, setting):
    for k in ("BOPY_B200_SWEEP_GROUP", "BOPY_B200_SWEEP_SLOTS", "BOPY_B200_SWEEP_LEAD"):
        os.environ.pop(k, None)
    if setting is not None:
        g, s, lead = setting
        os.environ["BOPY_B200_SWEEP_GROUP"] = str(g)
        os.environ["BOPY_B200_SWEEP_SLOTS"] = str(s)
        os.environ["BOPY_B200_SWEEP_LEAD"] = str(lead)
    if key.startswith("n"):                      # n<rows>d<dims>: the headline problem's kernel on another shape
        n, d = (int(v) for v in key[1:].split("d"))
        X, y, gp = bench.make_problem(n, d)
        lo, hi = np.zeros(d), np.ones(d)
    else:
        X, y, gp, lo, hi, _, _ = bench.config_problem(key)
    sur = B200GPSurrogate(gp, dtype="f64", device=torch.device("cuda", 0), device_fit=True)
    sur.fit(X, y)
    ei = EI(sur)
    ei.fit(X, y)
    sur.native.set_latency_path(0)
    return sur, float(ei._eta), lo, hi


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    once = "--once" in sys.argv
    key = args[0] if args else "C4"
    mlog = int(args[1]) if len(args) > 1 else 21
    settings = [tuple(int(v) for v in a.split(":")) if a != "default" else None for a in args[2:]] or [(8, 2, 2), (4, 2, 4)]
    m = 1 << mlog
    dev = torch.device("cuda", 0)
    mc = min(m, 148 * 128 * 3 + 77)      # ragged check size: three waves of tiles and a partial tile
    ref = None
    rows = []
    for setting in ([] if once else [(0, 2, 0), (1, 2, 0)]) + settings:
        sur, eta, lo, hi = make(key, setting)
        native = sur.native
        xs = _native.candidates_uniform(bench.SEED_CAND, 0, m, lo, hi, device=dev)
        if not once:
            out = native.sweep(xs[:mc], acq="ei", eta=eta, want_mean=True, want_var=True, want_acq=True, want_min=True)
            got = {k: out[k].cpu().numpy().copy() for k in ("mean", "var", "acq")}
            got["min_idx"] = int(out["min_idx"].item())
            got["min_val"] = float(out["min_val"].item())
            if ref is None:
                ref = got
                same = "reference"
            else:
                same = all(np.array_equal(ref[k], got[k], equal_nan=True) for k in ("mean", "var", "acq")) and \
                    ref["min_idx"] == got["min_idx"] and ref["min_val"] == got["min_val"]
        else:
            same = "-"
        steps = 1 if once else 3
        native.sweep(xs[: 1 << 15], acq="ei", eta=eta, want_min=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            o = native.sweep(xs, acq="ei", eta=eta, want_min=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        rows.append((setting, same, ms, m / ms * 1e3, int(o["min_idx"].item())))
        print(f"{key} m=2^{mlog} setting(G,S,lead)={setting}: identical={same}  {ms:.3f} ms  {m / ms * 1e3:.4g} evals/s  "
              f"argmin {int(o['min_idx'].item())}", flush=True)
        native.close()
        del xs
    if not once:
        base = rows[0][2]
        for r in rows[1:]:
            print(f"  {r[0]}: {base / r[2]:.4f} x the one-tile-per-CTA kernel")


if __name__ == "__main__":
    main()
