#!/usr/bin/env python
"""Compact per-launch table from `ncu -i X.ncu-rep --page raw --csv` (or the .ncu-rep itself): time, DRAM bytes, pipe utilisation, issue rate.

  python tools/ncu_table.py gpurun_out/r2p/step_full_raw.csv [--json profiles/r02_dram_traffic_per_launch.json] > profiles/r02_step_ncu_full_table.md"""
import csv
import json
import re
import subprocess
import sys

COLS = [("us", "gpu__time_duration.sum", 1.0), ("dram rd MB", "dram__bytes_read.sum", None), ("dram wr MB", "dram__bytes_write.sum", None),
        ("dram %", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1.0), ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0), ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0),
        ("fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0), ("lsu %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1.0),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0), ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
        ("regs", "launch__registers_per_thread", 1.0)]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def short(name):
    name = re.sub(r"\((?:int|bool)\)", "", name)
    name = re.sub(r"\((?:xfb::|const |ConvTc2Args|ConvArgs|MatchTcArgs).*$", "", name)
    return name.replace("void ", "").replace("xfb::", "")[:70]


def main():
    src = sys.argv[1]
    text = open(src).read() if src.endswith(".csv") else subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out, traffic = [], {}
    for vals in rows[2:]:
        name = short(vals[idx["Kernel Name"]])
        rec = [name, vals[idx["Grid Size"]].replace(" ", ""), vals[idx["Block Size"]].replace(" ", "")]
        for label, key, scale in COLS:
            i = idx.get(key)
            if i is None or vals[i] == "":
                rec.append("")
                continue
            v = float(vals[i].replace(",", ""))
            u = units[i]
            if label.endswith("MB") or label == "us":
                v *= UNIT.get(u, 1.0)
            rec.append("%.1f" % v if label != "regs" else "%d" % v)
        out.append(rec)
        t = traffic.setdefault(name, {"launches": 0, "bytes": 0.0})
        t["launches"] += 1
        t["bytes"] += (float(rec[4] or 0) + float(rec[5] or 0)) * 1e6
    print("| kernel | grid | block | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---|---|" + "---|" * len(COLS))
    for r in out:
        print("| `" + r[0] + "` | " + " | ".join(r[1:]) + " |")
    if "--json" in sys.argv:
        path = sys.argv[sys.argv.index("--json") + 1]
        json.dump({k: {"launches_captured": v["launches"], "dram_bytes_per_launch": v["bytes"] / v["launches"]} for k, v in traffic.items()}, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
