#!/usr/bin/env python
"""Per-range transport time of one rank's share (finds pathological photons).
usage: tools/gpu_find_slow.py photon_n_total rank world [range_positions]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
photon_n, rank, world = int(float(sys.argv[1])), int(sys.argv[2]), int(sys.argv[3])
step = int(float(sys.argv[4])) if len(sys.argv) > 4 else 8 << 20
start = int(float(sys.argv[5])) if len(sys.argv) > 5 else 0
end = int(float(sys.argv[6])) if len(sys.argv) > 6 else -1
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
hm = gm.HarmModel(photon_n, 4e19); hm.read_file(p); hm.init()
c = gm.Context(hm.model_dict(), rank=rank, world=world)
tot = c.total_primaries()
print("total", tot, flush=True)
prev = None
gs = start
if end > 0:
    tot = min(tot, end)
while gs < tot:
    ge = min(tot, gs + step)
    t0 = time.time(); c.run(gs, ge); dt = (time.time() - t0) * 1e3
    st = c.result()["stats"]
    d = {k: st[k] - (prev[k] if prev else 0) for k in ("transport_ms", "n_push_attempts", "n_scatter_events", "n_generations", "n_tracked", "n_steps")}
    print(f"[{gs:>10d},{ge:>10d}) wall {dt:8.1f} ms transport {d['transport_ms']:8.1f} gens {d['n_generations']:3d} attempts {d['n_push_attempts']:>11d} "
          f"scatters {d['n_scatter_events']:>8d} tracked {d['n_tracked']:>9d} ns/attempt {1e6*d['transport_ms']/max(1,d['n_push_attempts']):7.3f}", flush=True)
    prev = st; gs = ge
