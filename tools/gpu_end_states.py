#!/usr/bin/env python
"""End-of-run bias statistics (max_tau_scatt, scattered / created, recorded / created) of complete CUDA runs on the
bench dump at photon_n = 1e6 .. 8e6: what bench.py seeds the reference arm's bounded sample with (the reference's own
complete runs exist only at photon_n = 1e6, tests/golden/spectrum_192_4e19_1e6*.npz: 31 - 40 min each on one core).
usage (GPU box): tools/gpu_end_states.py [photon_n ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
for photon_n in [int(float(a)) for a in sys.argv[1:]] or [1000000, 2000000, 4000000, 8000000]:
    hm = gm.HarmModel(photon_n, 4e19); hm.read_file(p); hm.init()
    model = hm.model_dict()
    for seed in (123, 124):
        c = gm.Context(model, seed=seed)
        t0 = time.time(); c.run(); dt = time.time() - t0
        r = c.result(); c.close()
        print(json.dumps(dict(photon_n=photon_n, seed=seed, created=r["created"], max_tau_scatt=r["max_tau_scatt"],
                              scattered_per_created=r["scattered"] / r["created"],
                              recorded_per_created=r["recorded"] / r["created"], run_s=dt,
                              rate=r["created"] / dt)), flush=True)
