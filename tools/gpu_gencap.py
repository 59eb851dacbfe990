#!/usr/bin/env python
"""Throughput at configs[1] (photon_n = 1e6) for several generation-schedule settings."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
hm = gm.HarmModel(1000000, 4e19); hm.read_file(p); hm.init()
model = hm.model_dict()
for cfg in (dict(), dict(gen_budget=256), dict(gen_budget=512)):
    c = gm.Context(model, **cfg)
    c.run(0, 20000); c.reset()
    t0 = time.time(); c.run(); dt = time.time() - t0
    r = c.result(); st = r["stats"]; c.close()
    print(json.dumps(dict(cfg=cfg, ms=dt * 1e3, rate=r["created"] / dt, gens=st["n_generations"],
                          recorded=r["recorded"], scattered=r["scattered"],
                          occ=st["n_live_iterations"] / max(1, st["n_slot_iterations"]))), flush=True)
