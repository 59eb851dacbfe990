#!/bin/bash
set -u
out=gpurun_out
mkdir -p $out
cat > /tmp/p5.py <<'PY'
import os, sys, time, json
sys.path.insert(0, os.getcwd())
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
photon_n, ov = int(float(sys.argv[1])), int(sys.argv[2])
hm = gm.HarmModel(photon_n, 4e19); hm.read_file(p); hm.init()
c = gm.Context(hm.model_dict(), gen_overlap=ov)
c.run(0, 20000); c.reset()
best = None
for rep in range(2):
    c.reset(); t0 = time.time(); c.run(); r = c.result(); dt = time.time() - t0
    s = r["stats"]
    row = dict(photon_n=photon_n, overlap=ov, wall_ms=round(1e3 * dt, 1), transport_ms=round(s["transport_ms"], 1), recorded=r["recorded"],
               scattered=r["scattered"], gens=s["n_generations"], occ=round(s["n_live_iterations"] / max(1, s["n_slot_iterations"]), 4))
    if best is None or row["wall_ms"] < best["wall_ms"]:
        best = row
print(json.dumps(best))
PY
for ov in 1 2; do
GRMONTY_B200_TRACE=1 timeout 300 python /tmp/p5.py 1e5 $ov > $out/p6_trace_$ov.log 2>&1
done
grep -E "generation [0-9]+:" $out/p6_trace_1.log | tail -35
grep -E "batch first" $out/p6_trace_2.log | tail -36 | cut -c1-120
