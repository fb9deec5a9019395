#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into a small CSV of the metrics DESIGN.md / VERDICT cite.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN/name.csv"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keep = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__inst_executed.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.avg",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
w = csv.writer(sys.stdout)
for k, vals in enumerate(rows[2:]):
    w.writerow([f"# launch {k}"])
    w.writerow(["metric", "unit", "value"])
    for i, h in enumerate(hdr):
        if h in keep or ("warp_issue_stalled" in h and h.endswith("per_warp_active.pct")):
            w.writerow([h, units[i], vals[i]])
