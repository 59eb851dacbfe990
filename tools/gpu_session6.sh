#!/bin/bash
set -u
out=gpurun_out
mkdir -p $out
GRMONTY_B200_TRACE=1 timeout 300 python - > $out/s6_trace.log 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
hm = gm.HarmModel(1000000, 4e19); hm.read_file(p); hm.init()
m = hm.model_dict()
c = gm.Context(m, threads_per_block=256, slots_per_thread=2)
c.run(); r = c.result(); print("transport_ms", r["stats"]["transport_ms"], flush=True); c.close()
PY
grep -A1 "batch first" $out/s6_trace.log | grep -v "^--" | paste - - | awk '{print $3,$4,$5,$8,$9,"|",$0}' | cut -c1-60 | head -0
grep -A1 "batch first" $out/s6_trace.log | grep -v "^--" | paste - - | sed 's/\[grmonty_b200\] batch //; s/records=.*carried_out=[0-9]*//' | cut -c1-420 | tail -45
