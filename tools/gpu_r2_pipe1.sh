#!/bin/bash
# first GPU visit of the pipelined scheduler: correctness (small), then timing against the round-1 scheduler
set -u
out=gpurun_out
mkdir -p $out
export GRMONTY_B200_WATCHDOG_S=15
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_invariance.py tests/test_gpu_edge.py -q -x -k "full_run or invariance or edge or sharding" > $out/p1_tests.log 2>&1
echo "tests rc=$?"; tail -25 $out/p1_tests.log
for ov in 1 0; do
GRMONTY_B200_OVERLAP=$ov GRMONTY_B200_TRACE=1 timeout 300 python - > $out/p1_trace_$ov.log 2>&1 <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
for photon_n in (100000, 1000000):
    hm = gm.HarmModel(photon_n, 4e19); hm.read_file(p); hm.init()
    c = gm.Context(hm.model_dict())
    c.run(0, 20000); c.reset()
    for rep in range(2):
        c.reset(); t0 = time.time(); c.run(); r = c.result(); dt = time.time() - t0
        s = r["stats"]
        print("RESULT", photon_n, "wall_ms", round(1e3 * dt, 2), "transport_ms", round(s["transport_ms"], 2), "recorded", r["recorded"],
              "scattered", r["scattered"], "created", r["created"], "attempts", s["n_push_attempts"], "gens", s["n_generations"],
              "occ", round(s["n_live_iterations"] / max(1, s["n_slot_iterations"]), 4), "max_tau", r["max_tau_scatt"], flush=True)
    c.close()
PY
echo "== overlap $ov rc=$?"; grep -E "RESULT|window|Error|error" $out/p1_trace_$ov.log | tail -14
done
