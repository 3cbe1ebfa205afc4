#!/usr/bin/env python
"""Key metrics of one kernel from `ncu -i X.ncu-rep --page raw --csv`."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_active.avg",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor"]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("==", d.get("Kernel Name", "")[:80])
    for k in KEYS:
        if k in d:
            print(f"  {k:70s} {d[k]:>16s} {units[hdr.index(k)]}")
    for h, u, v in zip(hdr, units, vals):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.05:
            print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {float(v):6.2f}")
