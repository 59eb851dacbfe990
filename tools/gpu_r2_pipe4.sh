#!/bin/bash
# ncu --set full of one pipelined launch (photon_n = 1e5: one window, 1.6 M primaries, 35 generations)
set -u
out=gpurun_out
mkdir -p $out
cat > /tmp/pipe_run.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
hm = gm.HarmModel(int(float(sys.argv[1])), 4e19); hm.read_file(p); hm.init()
c = gm.Context(hm.model_dict())
c.run(0, 2000); c.reset()
c.run(); r = c.result(); print(r["stats"]["transport_ms"], r["recorded"]); c.close()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pipeline_kernel --launch-skip 1 -c 1 -f -o $out/p4_pipeline python /tmp/pipe_run.py 1e5 > $out/p4_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $out/p4_ncu.log
