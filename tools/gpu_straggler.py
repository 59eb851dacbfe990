#!/usr/bin/env python
"""Time one known pathological lineage (primary 76576157 of the photon_n = 8e6 run: ~7000 scatterings of one parent)
with frozen statistics, through the t_track test export."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
hm = gm.HarmModel(8000000, 4e19); hm.read_file(p); hm.init()
c = gm.Context(hm.model_dict())
idx = np.array([int(a) for a in sys.argv[1:]] or [76576157], dtype=np.int64)
births, rng = c.t_make_primaries(idx)
for rep in range(2):
    c.reset()
    t0 = time.time()
    got, _, status = c.t_track(births, rng, 0.0025, 0.86e6, 1.0e6)
    dt = time.time() - t0
    st = c.result()["stats"]
    print(f"track {len(idx)} primaries: {dt*1e3:.1f} ms  transport {st['transport_ms']:.1f} ms  tracked {st['n_tracked']} scatters {st['n_scatter_events']} "
          f"attempts {st['n_push_attempts']} steps {st['n_steps']}", flush=True)
