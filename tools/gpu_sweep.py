#!/usr/bin/env python
"""Throughput of kernel variants on the 192x192 bench dump: one subprocess per (library, threads x blocks/SM).
usage: tools/gpu_sweep.py photon_n lib1.so,lib2.so w384x2,w256x3,f256x1,...   (run on the GPU box)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, time
sys.path.insert(0, %r)
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
photon_n, kern, t, b, mb = int(float(sys.argv[1])), sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
hm = gm.HarmModel(photon_n, 4e19); hm.read_file(p); hm.init()
kw = dict(kernel=2, slots_per_thread=b, blocks_per_sm=mb) if kern == "w" else dict(kernel=1, blocks_per_sm=b)
c = gm.Context(hm.model_dict(), threads_per_block=t, **kw)
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
# w384x2 (wavefront: threads x slots), w128x2x3 (... x blocks per SM of the compiled variant), f256x1 (fused)
variants = [(v[0], *(v[1:].split("x") + ["0"])[:3]) for v in sys.argv[3].split(",")]
for lib in libs:
    for kern, t, b, mb in variants:
        env = dict(os.environ)
        if lib != "default":
            env["GRMONTY_B200_LIB"] = os.path.join(ROOT, lib)
        try:
            out = subprocess.run([sys.executable, "-c", CHILD, photon_n, kern, t, b, mb], env=env, capture_output=True, text=True,
                                 timeout=300)
            line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr.strip()[-300:]
            for ln in out.stderr.splitlines():
                if "wavefront block-phases" in ln:
                    last_phase_line = ln
            if os.environ.get("GRMONTY_B200_TRACE") and "last_phase_line" in dir():
                line += "\n    " + last_phase_line
        except subprocess.TimeoutExpired:
            line = "timeout"
        print(f"{lib} {kern}{t}x{b}x{mb} {line}", flush=True)
