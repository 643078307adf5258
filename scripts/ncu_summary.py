#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "smsp__inst_executed_op_shared_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_global_loads"]
units = dict(zip(hdr, rows[1]))
for vals in rows[2:]:                                                     # one block per captured launch
    d = dict(zip(hdr, vals))
    if len(rows) > 3: print(f"## {d.get('Kernel Name', '?')}")
    for k in want:
        if k in d: print(f"{k} = {d[k]} {units.get(k,'')}")
    for k in hdr:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                if float(d[k]) > 0.2: print(f"{k} = {d[k]}")
            except ValueError: pass
