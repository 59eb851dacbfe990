#!/usr/bin/env python
"""Key metrics of one ncu --set full capture as text (what profiles/*.txt hold).  usage: ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, d = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Grid Size", "Block Size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# ncu --set full --clock-control none, report {rep.split('/')[-1]}")
for k in want:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:75s} {d[i]:>20s} {units[i]}")
print("# warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active.ratio)")
for i, k in enumerate(hdr):
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        print(f"  {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {float(d[i]):8.3f}")
