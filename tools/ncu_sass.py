#!/usr/bin/env python
"""Dump the SASS stream (address order) of an ncu report with source-line tags and per-instruction counters.
usage: tools/ncu_sass.py report.ncu-rep > sass.txt"""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
sass = []; cur = None; line = None; ix = None
for r in csv.reader(io.StringIO(txt)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if ix and len(r) == len(hdr):
        if r[0]:
            line = r[0]; continue
        try:
            addr = int(r[2], 16); n = int(r[ix["Instructions Executed"]])
        except ValueError:
            continue
        sass.append((addr, cur, line, r[3].strip(), n, r[ix["# Samples"]], r[ix["Avg. Threads Executed"]],
                     r[ix["stall_wait"]], r[ix["stall_long_sb"]], r[ix["stall_barrier"]]))
sass.sort()
base = sass[0][0]
print("# offset file:line executed samples threads wait long_sb barrier | sass")
for s in sass:
    print(f"{s[0]-base:06x} {s[1].replace('gm_','').replace('.cuh',''):>12s}:{s[2]:>4s} {s[4]:>10d} {s[5]:>6s} {s[6]:>3s} "
          f"{s[7]:>5s} {s[8]:>5s} {s[9]:>5s} | {s[3][:90]}")
