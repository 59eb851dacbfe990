#!/usr/bin/env python
"""Summarise an ncu report's source page per CUDA source line / file / SASS opcode.
usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None; ix = None
fsamp = collections.Counter(); finst = collections.Counter(); lines = []
ops = collections.Counter(); osamp = collections.Counter(); othr = collections.Counter()
stall = collections.Counter()
for r in csv.reader(io.StringIO(txt)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]; continue
    if r and r[0] == "Line No":
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if not ix or len(r) != len(hdr):
        continue
    if r[0]:  # a CUDA source line with aggregated metrics
        try:
            s = int(r[ix["# Samples"]]); n = int(r[ix["Instructions Executed"]]); t = int(r[ix["Thread Instructions Executed"]])
        except ValueError:
            continue
        fsamp[cur] += s; finst[cur] += n
        lines.append((s, n, t, cur.split("/")[-1], r[0], r[1].strip()[:100]))
        for h, i in ix.items():
            if h.startswith("stall_") and "Not Issued" not in h:
                stall[h] += int(r[i] or 0)
    else:
        sass = r[3].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", sass)
        op = m.group(2) if m else sass
        try:
            n = int(r[ix["Instructions Executed"]])
        except ValueError:
            continue
        ops[op] += n
        osamp[op] += int(r[ix["# Samples"]]); othr[op] += int(r[ix["Thread Instructions Executed"]])
T = sum(fsamp.values()) or 1; N = sum(finst.values()) or 1
print(f"total samples {T}, warp instructions {N/1e9:.2f} G")
print("-- stall reasons (share of samples)")
for k, v in stall.most_common(12):
    print(f"  {v/T*100:6.2f}%  {k}")
print("-- per file")
for f, s in fsamp.most_common():
    print(f"  {s/T*100:6.2f}% samples {finst[f]/N*100:6.2f}% inst  {f}")
print("-- per opcode")
for k, v in ops.most_common(24):
    print(f"  {k:9s} {v/N*100:6.2f}% inst {osamp[k]/T*100:6.2f}% samples  avg threads {othr[k]/max(v,1):5.1f}")
print("-- top source lines")
lines.sort(reverse=True)
for s, n, t, f, l, src in lines[:top]:
    print(f"  {s/T*100:5.2f}% {n/N*100:5.2f}%inst thr {t/max(n,1):4.1f}  {f}:{l}  {src}")
