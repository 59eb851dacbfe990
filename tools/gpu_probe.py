#!/usr/bin/env python
"""Quick GPU probe on the 48x48 golden model: FP64 peak, variant sweep, work counters."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_grmonty_b200 as gm

d = dict(np.load(os.path.join(ROOT, "tests/golden/functions_48.npz")))
m = {k[6:]: (v.item() if v.ndim == 0 else v) for k, v in d.items() if k.startswith("model_")}
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
variants = [(128, 2), (128, 3), (128, 4), (256, 1), (64, 4)]
if len(sys.argv) > 2:
    variants = [tuple(int(v) for v in a.split("x")) for a in sys.argv[2:]]
if scale != 1.0:  # more photons on the same model: weights scale as 1/photon_n (reference harm_model.cpp:302)
    m["photon_n"] *= scale
    m["weight"] = m["weight"] - np.log(scale)
    m["nint"] = m["nint"] + np.log(scale)
    m["dndlnu_max"] = m["dndlnu_max"] + np.log(scale)
out = []
for (t, b) in variants:
    try:
        c = gm.Context(m, threads_per_block=t, blocks_per_sm=b)
    except Exception as e:
        print("variant", t, b, "failed:", e); continue
    if not out:
        print("fp64 peak TFLOP/s:", c.fp64_peak(), "total primaries:", c.total_primaries())
    c.run(0, 2000)  # warm-up
    c.reset()
    t0 = time.time(); c.run(); dt = time.time() - t0
    r = c.result(); st = r["stats"]
    rec = dict(threads=t, bps=b, wall_s=dt, transport_ms=st["transport_ms"], kernel_ms=st["kernel_ms"],
               primaries=r["created"], rate=r["created"] / dt, steps=st["n_steps"], attempts=st["n_push_attempts"],
               interactions=st["n_interactions"], scatters=st["n_scatter_events"], tracked=st["n_tracked"],
               recorded=r["recorded"], scattered=r["scattered"], gens=st["n_generations"],
               steps_per_s=st["n_steps"] / (st["transport_ms"] * 1e-3))
    print(json.dumps(rec)); out.append(rec)
    c.close()
