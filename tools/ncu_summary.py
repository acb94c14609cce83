"""Summarise `.ncu-rep` captures (read here with `ncu -i ... --page raw --csv`) into the JSON files under profiles/r02/.

    python tools/ncu_summary.py gpurun_out/ncu_c4_f64_raw.csv  profiles/r02/ncu_sweep_f64_c4_summary.json  [traffic-key]

Writes the metrics the roofline discussion uses (duration, DRAM bytes, L2 hit rate, pipe utilisation, registers, shared
memory, stall reasons) for every kernel in the report and, with a traffic key ("C4_f64", "C4_f32", ...), records
dram__bytes_read.sum + dram__bytes_write.sum of the FIRST kernel in profiles/r02/ncu_traffic.json -- the file bench.py
reads `roofline.traffic` from, so that the number comes from the build that is timed.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
]


def rows_of(rep):
    """`rep`: a .ncu-rep (read through ncu) or the `ncu -i ... --page raw --csv` dump made on the GPU box."""
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    hdr, units, data = rows_of(rep)
    col = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in data:
        k = {"kernel": r[col["Kernel Name"]], "metrics": {}}
        for name in KEEP:
            if name in col and r[col[name]] != "":
                k["metrics"][name] = {"value": float(r[col[name]].replace(",", "")), "unit": units[col[name]]}
        stalls = {h.split("issue_stalled_")[1].split("_per_")[0]: float(r[i]) for h, i in col.items()
                  if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and r[i] != ""}
        k["warp_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        kernels.append(k)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as f:
        json.dump({"source": os.path.basename(rep), "command": "ncu --set full --clock-control none --import-source on",
                   "kernels": kernels}, f, indent=1)
    if key and kernels:
        m = kernels[0]["metrics"]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        rd = m["dram__bytes_read.sum"]["value"] * scale[m["dram__bytes_read.sum"]["unit"]]
        wr = m["dram__bytes_write.sum"]["value"] * scale[m["dram__bytes_write.sum"]["unit"]]
        tf = os.path.join(ROOT, "profiles", "r02", "ncu_traffic.json")
        table = json.load(open(tf)) if os.path.exists(tf) else {}
        table[key] = {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "kernel": kernels[0]["kernel"],
                      "l2_hit_pct": m.get("lts__t_sector_hit_rate.pct", {}).get("value"), "from": os.path.basename(dst)}
        with open(tf, "w") as f:
            json.dump(table, f, indent=1)
    for k in kernels:
        m = k["metrics"]
        print(k["kernel"][:70], {n.split(".")[0]: round(v["value"], 3) for n, v in list(m.items())[:6]})


if __name__ == "__main__":
    main()
