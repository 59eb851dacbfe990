#!/usr/bin/env python
"""Throughput at configs[1] for several (queue_capacity, gen_cap) pairs."""
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
for cap, gcap, budget in ((1 << 22, 1 << 20, 256), (1 << 23, 1 << 21, 256), (1 << 24, 1 << 22, 256), (1 << 24, 1 << 22, 512), (1 << 22, 1 << 20, 128)):
    c = gm.Context(model, queue_capacity=cap, gen_cap=gcap, gen_budget=budget)
    c.run(0, 20000); c.reset()
    t0 = time.time(); c.run(); dt = time.time() - t0
    r = c.result(); st = r["stats"]; c.close()
    print(json.dumps(dict(cap=cap, gen_cap=gcap, budget=budget, ms=dt * 1e3, rate=r["created"] / dt, gens=st["n_generations"],
                          launches=st["n_kernel_launches"], recorded=r["recorded"], scattered=r["scattered"],
                          occ=st["n_live_iterations"] / max(1, st["n_slot_iterations"]), high_water=st["queue_high_water"])), flush=True)
