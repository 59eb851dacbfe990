#!/usr/bin/env python
"""Per-generation timing of the transport path on the 48x48 golden model (scaled)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_grmonty_b200 as gm
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
budget = int(sys.argv[2]) if len(sys.argv) > 2 else 0
grid = int(sys.argv[3]) if len(sys.argv) > 3 else 48
spec = sys.argv[4] if len(sys.argv) > 4 else "w0x0"   # w384x2: wavefront 384 threads x 2 slots; f256x1: fused loop
tb = [int(v) for v in spec[1:].split("x")]
kw = dict(kernel=2, slots_per_thread=tb[1]) if spec[0] == "w" else dict(kernel=1, blocks_per_sm=tb[1])
if grid == 48:
    d = dict(np.load(os.path.join(ROOT, "tests/golden/functions_48.npz")))
    m = {k[6:]: (v.item() if v.ndim == 0 else v) for k, v in d.items() if k.startswith("model_")}
    m["photon_n"] *= scale; m["weight"] = m["weight"] - np.log(scale)
    m["nint"] = m["nint"] + np.log(scale); m["dndlnu_max"] = m["dndlnu_max"] + np.log(scale)
else:
    from tools import make_harm_dump
    p = f"/tmp/gp_dump_{grid}.txt"
    if not os.path.exists(p):
        make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=grid, n1=grid))
    gm.build_host()
    hm = gm.HarmModel(int(2000 * scale), 4e19); hm.read_file(p); hm.init(); m = hm.model_dict()
c = gm.Context(m, gen_budget=budget, threads_per_block=tb[0], **kw)
tot = c.total_primaries()
c.run(0, 2000); c.reset()
gs, g = 0, 0
prev = dict(transport_ms=0.0, n_push_attempts=0, n_tracked=0, n_kernel_launches=0, n_live_iterations=0, n_slot_iterations=0)
while gs < tot:
    size = min(32 << g, 1 << 22); ge = min(gs + size, tot)
    t0 = time.time(); c.run(gs, ge); dt = (time.time() - t0) * 1e3
    st = c.result()["stats"]
    att = st["n_push_attempts"] - prev["n_push_attempts"]; ms = st["transport_ms"] - prev["transport_ms"]
    ideal_iters = att / 37888.0
    print(f"gen {g:2d} n={ge-gs:8d} wall {dt:8.2f} ms transport {ms:8.2f} ms launches {st['n_kernel_launches']-prev['n_kernel_launches']:3d} "
          f"occ {(st['n_live_iterations']-prev['n_live_iterations'])/max(1,st['n_slot_iterations']-prev['n_slot_iterations']):5.3f} slot_it/thr {(st['n_slot_iterations']-prev['n_slot_iterations'])/37888:8.1f} attempts {att:10d} tracked {st['n_tracked']-prev['n_tracked']:8d} ideal_iters {ideal_iters:9.1f} us/ideal_iter {ms*1e3/max(ideal_iters,1):8.2f}")
    prev = st; gs = ge; g += 1
