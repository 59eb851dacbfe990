#!/bin/bash
# hybrid-lag pipeline: correctness subset, then timing A/B (variants in gpurun_ab) at 1e6 / 1e5 for gen_overlap 1, 3 and the round-1 scheduler
set -u
out=gpurun_out
mkdir -p $out
export GRMONTY_B200_WATCHDOG_S=20
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_invariance.py tests/test_gpu_edge.py -q -k "full_run or invariance or edge or sharding" > $out/p5_tests.log 2>&1
echo "tests rc=$?"; tail -12 $out/p5_tests.log | cut -c1-300
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
for lib in "$@"; do for pn in 1e6 1e5; do for ov in 1 3 2; do
  GRMONTY_B200_LIB=$PWD/gpurun_ab/lib_$lib.so timeout 300 python /tmp/p5.py $pn $ov 2>&1 | tail -1 | sed "s/^/$lib /" | tee -a $out/p5_ab.txt
done; done; done
