#!/usr/bin/env python
"""BASELINE.json configs[3] (M_unit = 4e20, Compton-dominated) and configs[4] (1024x1024 grid) on the CUDA path:
work counters, rates and size-independent checks.  usage (GPU box): tools/gpu_configs.py [c4] [c5] [photon_n]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
what = [a for a in sys.argv[1:] if a in ("c4", "c5")] or ["c4", "c5"]
pn = [float(a) for a in sys.argv[1:] if a not in ("c4", "c5")]
photon_n = int(pn[0]) if pn else 100000


def run(tag, n, mass_unit, photon_n):
    t0 = time.time()
    p = f"/tmp/gp_dump_{n}.txt"
    if not os.path.exists(p):
        make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=n, n1=n))
    t1 = time.time()
    hm = gm.HarmModel(photon_n, mass_unit); hm.read_file(p)
    t2 = time.time()
    hm.init()
    t3 = time.time()
    c = gm.Context(hm.model_dict(), seed=123)
    t4 = time.time()
    c.run()
    t5 = time.time()
    r = c.result(); st = r["stats"]; c.close()
    spec = r["spectrum"]
    out = dict(config=tag, grid=n, mass_unit=mass_unit, photon_n=photon_n, write_dump_s=t1 - t0, read_file_s=t2 - t1, init_s=t3 - t2,
               create_s=t4 - t3, run_s=t5 - t4, created=r["created"], recorded=r["recorded"], scattered=r["scattered"],
               rate=r["created"] / (t5 - t4), tracked_per_primary=st["n_tracked"] / max(1, r["created"]),
               scatter_events_per_primary=st["n_scatter_events"] / max(1, r["created"]),
               steps_per_tracked=st["n_steps"] / max(1, st["n_tracked"]), gens=st["n_generations"],
               transport_ms=st["transport_ms"], queue_high_water=st["queue_high_water"],
               # size-independent checks: counts are consistent, the spectrum holds every recorded photon once
               nph_sum_equals_recorded=bool(spec[:, :, 2].sum() == r["recorded"]),
               nscatt_sum_equals_scattered=bool(spec[:, :, 3].sum() == r["scattered"]),
               finite=bool(np.isfinite(spec).all()), lum=float(spec[:, :, 1].sum()))
    print(json.dumps(out), flush=True)


if "c4" in what:
    run("configs[3] Compton-dominated", 192, 4e20, photon_n)
if "c5" in what:
    run("configs[4] large grid", 1024, 4e19, photon_n)
