"""The driver's entry points keep working: bench.py's CPU legs (no GPU needed) and __graft_entry__.smoke()
(GPU).  Both once broke silently when the default fit moved to the device and the scikit-learn object stopped
holding L_."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_cpu_baseline_leg_runs_on_a_fresh_model():
    sys.path.insert(0, ROOT)
    import bench
    from sklearn.base import clone
    X, y, gp = bench.make_problem(96, 6)
    gp_host = clone(gp).fit(X, y)
    done, dt = bench.time_reference_path(gp_host, float(np.min(y)), np.zeros(6), np.ones(6), budget_s=0.3,
                                         max_candidates=256)
    assert done >= 64 and dt > 0.0


def test_bench_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-budget", "2"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "evals/s"
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_smoke_entry_point():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.smoke()


@pytest.mark.gpu
def test_bench_default_line_small():
    """bench.py end to end on a reduced candidate count: every key of the contract is present."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3",
                          "--candidates", "65536", "--c5-candidates", "32768", "--cpu-budget", "1"], capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in line, key
    assert line["cpu_baseline"] is not None and line["cpu_baseline"]["value"] > 0
    assert line["roofline"]["frac"] > 0 and line["gpu_launches"] > 0
    assert line["argmin"]["index"] == line["argmin"]["e2e_index"]
    # every BASELINE config is in the line, on its own inputs, with a green parity spot-check against the golden vectors
    for key in ("C1", "C3", "C4_f32", "C5"):
        cfg = line["configs"][key]
        assert "error" not in cfg, cfg
        assert cfg["value"] > 0 and cfg["roofline"]["frac"] > 0 and cfg["parity"]["ok"], (key, cfg.get("parity"))
    assert line["configs"]["C5"]["multistart_1024"]["f_min"] < 0
    assert set(line["strong_scaling"]) == {"C4", "C3"} and all("error" not in v for v in line["strong_scaling"].values())
