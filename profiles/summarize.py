#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report -> CSV (one `metric,value` line each).  Usage: summarize.py report.ncu-rep > out.csv"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_local", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
head, units, vals = rows[0], rows[1], rows[-1]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit", "value"])
for h, u, v in zip(head, units, vals):
    if h in ("Kernel Name", "Grid Size", "Block Size") or any(h.startswith(k) for k in KEYS):
        if "pcsamp" in h or "not_issued" in h:
            continue
        w.writerow([h, u, v])
