#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report, one block per profiled launch (the files under profiles/*_ncu_full_*.txt).
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv; python tools/ncu_full_summary.py raw.csv "header line" > profiles/rNN_ncu_full_x.txt"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ci = {n: i for i, n in enumerate(hdr)}
if len(sys.argv) > 2:
    print(sys.argv[2])
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    print("---")
    print("Kernel Name =", r[ci["Kernel Name"]][:200])
    for w in WANT:
        if w in ci:
            print(f"{w} = {r[ci[w]]} {units[ci[w]]}")
