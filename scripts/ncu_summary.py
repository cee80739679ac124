#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into the few lines committed under profiles/:
    python scripts/ncu_summary.py gpurun_out/X.ncu-rep > profiles/X.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
print(f"# {rep}")
for i, h in enumerate(hdr):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")):
        v = vals[i]
        try:
            if h.startswith("smsp__average_warps") and float(v.replace(",", "")) < 0.05:
                continue
        except ValueError:
            pass
        print(f"{h:90s} {units[i]:12s} {v}")
