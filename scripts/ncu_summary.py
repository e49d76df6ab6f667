#!/usr/bin/env python
"""Text summary of an .ncu-rep (raw page) for profiles/: the metrics B200_PROFILING.md names plus
the instruction mix. usage: ncu_summary.py <report.ncu-rep> [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
ik = hdr.index("Kernel Name")
for d in data:
    if pat and not pat.search(d[ik]):
        continue
    print(f"== {d[ik]}  (launch id {d[0]})")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:70s} {d[i]:>18s} {units[i]}")
    for k in hdr:
        if k.startswith("smsp__warp_issue_stalled") and k.endswith("_per_warp_active.pct"):
            i = hdr.index(k)
            try:
                if float(d[i]) >= 3.0:
                    print(f"  {k:70s} {d[i]:>18s} %")
            except ValueError:
                pass
    for k in hdr:
        if k.startswith("smsp__sass_thread_inst_executed_op_d") and k.endswith("_pred_on.sum"):
            i = hdr.index(k)
            print(f"  {k:70s} {d[i]:>18s}")
