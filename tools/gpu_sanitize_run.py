#!/usr/bin/env python
"""A small complete run for compute-sanitizer (memcheck / racecheck): 48x48 golden model, M_unit raised so that
scattering, carry-over (small budget) and the issue-order sort are all active.  usage: compute-sanitizer --tool
memcheck python tools/gpu_sanitize_run.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_grmonty_b200 as gm
gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "functions_48.npz")))
model = {k[6:]: (v.item() if v.ndim == 0 else v) for k, v in gold.items() if k.startswith("model_")}
last = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
ctx = gm.Context(model, seed=123, device=0, gen0=64, gen_cap=1 << 11, gen_budget=96, queue_capacity=1 << 15)
ctx.run(0, last)
r = ctx.result()
st = r["stats"]
print("sanitize run:", dict(created=r["created"], recorded=r["recorded"], scattered=r["scattered"],
                            scatter_events=st["n_scatter_events"], generations=st["n_generations"],
                            launches=st["n_kernel_launches"], attempts=st["n_push_attempts"]))
assert r["created"] == last and r["recorded"] > 0
ctx.close()
