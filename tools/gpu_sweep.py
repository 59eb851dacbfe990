#!/usr/bin/env python
"""Throughput of kernel variants on the 192x192 bench dump: one subprocess per (library, threads x blocks/SM).
usage: tools/gpu_sweep.py photon_n lib1.so,lib2.so 256x1,128x2,...   (run on the GPU box)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, time
sys.path.insert(0, %r)
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
photon_n, t, b = int(float(sys.argv[1])), int(sys.argv[2]), int(sys.argv[3])
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
hm = gm.HarmModel(photon_n, 4e19); hm.read_file(p); hm.init()
c = gm.Context(hm.model_dict(), threads_per_block=t, blocks_per_sm=b)
c.run(0, 20000); c.reset()
best = None
for rep in range(2):
    c.reset(); t0 = time.time(); c.run(); dt = time.time() - t0
    r = c.result(); st = r["stats"]
    rec = dict(wall_ms=dt * 1e3, transport_ms=st["transport_ms"], rate=r["created"] / dt, recorded=r["recorded"],
               scattered=r["scattered"], attempts=st["n_push_attempts"], gens=st["n_generations"],
               occ=st["n_live_iterations"] / max(1, st["n_slot_iterations"]))
    if best is None or rec["wall_ms"] < best["wall_ms"]:
        best = rec
print(json.dumps(best))
''' % ROOT
photon_n = sys.argv[1]
libs = sys.argv[2].split(",")
variants = [v.split("x") for v in sys.argv[3].split(",")]
for lib in libs:
    for t, b in variants:
        env = dict(os.environ)
        if lib != "default":
            env["GRMONTY_B200_LIB"] = os.path.join(ROOT, lib)
        try:
            out = subprocess.run([sys.executable, "-c", CHILD, photon_n, t, b], env=env, capture_output=True, text=True,
                                 timeout=300)
            line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr.strip()[-300:]
        except subprocess.TimeoutExpired:
            line = "timeout"
        print(f"{lib} {t}x{b} {line}", flush=True)
