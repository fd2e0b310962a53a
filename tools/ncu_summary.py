#!/usr/bin/env python
"""Print the metrics that matter from an .ncu-rep (reads it with `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum ",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ",
        "smsp__inst_executed.sum ", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread ", "launch__occupancy_limit",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu", "sm__pipe_fma_cycles_active.avg.pct",
        "sm__pipe_fp64", "smsp__average_warps_issue_stalled", "sm__cycles_elapsed.max ", "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_lsu", "l1tex__t_bytes.sum ", "smsp__inst_executed_op_shared", "lts__t_sector_hit_rate.pct", "sm__sass_inst_executed_op_shared"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("==", d.get("Kernel Name", "")[:110], d.get("Grid Size", ""), d.get("Block Size", ""))
        for h, u, v in zip(hdr, units, vals):
            if any(h.startswith(k.strip()) if k.endswith(" ") else k in h for k in KEYS) and v not in ("", "0"):
                if "stalled" in h and "per_issue_active" not in h:
                    continue
                print("   %-92s %-10s %s" % (h, u, v))


if __name__ == "__main__":
    main()
