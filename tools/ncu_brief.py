#!/usr/bin/env python
"""headline metrics + stall reasons of the first kernel in an .ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
keep = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__thread_inst_executed_per_inst_executed.ratio")
for i, n in enumerate(h):
    if n in keep or ("issue_stalled" in n and "per_issue_active" in n and float(v[i] or 0) > 0.04):
        print(f"{n:90s} {u[i]:14s} {v[i]}")
