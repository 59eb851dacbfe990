#!/bin/bash
set -u
out=gpurun_out
mkdir -p $out
for spec in 256,2,0 384,2,0 512,1,0 128,2,2; do
IFS=, read t r mb <<< "$spec"
GRMONTY_B200_TRACE=1 timeout 300 python - $t $r $mb > $out/s7_trace_$t.$r.$mb.log 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
hm = gm.HarmModel(1000000, 4e19); hm.read_file(p); hm.init()
m = hm.model_dict()
t, r, mb = (int(v) for v in sys.argv[1:4])
c = gm.Context(m, threads_per_block=t, slots_per_thread=r, blocks_per_sm=mb)
c.run(); c.reset(); c.run(); r = c.result(); print("transport_ms", r["stats"]["transport_ms"], "recorded", r["recorded"], flush=True); c.close()
PY
echo "== $spec"; grep transport_ms $out/s7_trace_$t.$r.$mb.log
grep -A1 "batch first" $out/s7_trace_$t.$r.$mb.log | grep -v "^--" | paste - - | sed 's/\[grmonty_b200\] batch //; s/records=.*carried_out=[0-9]*//' | cut -c1-420 | awk 'NR%6==0' | tail -10
done
