#!/usr/bin/env python
"""Time the init tables: threaded host builders vs grmonty_b200_init_tables.
usage (GPU box): tools/gpu_init_tables.py [n ...]   default 192 1024"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cuda_grmonty_b200 as gm
from tools import make_harm_dump

for n in [int(a) for a in sys.argv[1:]] or [192, 1024]:
    p = f"/tmp/it_dump_{n}.txt"
    if not os.path.exists(p):
        make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=n, n1=n))
    hm = gm.HarmModel(100000, 4e19)
    hm.read_file(p)
    hm.init_stage(2)
    t = {}
    for threads in (1, 0):
        for name, stage in (("geometry", 0), ("weight", 3), ("nint", 4)):
            t0 = time.perf_counter(); hm.init_stage(stage, threads); t[(name, threads)] = time.perf_counter() - t0
    host = hm.model_dict()
    gm.init_tables(host)  # warm-up: CUDA context + module load
    w = []
    for _ in range(3):
        t0 = time.perf_counter(); dev = gm.init_tables(host); w.append(time.perf_counter() - t0)
    err = max(float(np.max(np.abs(dev["weight"] - host["weight"]))),
              float(np.max(np.abs((dev["nint"] - host["nint"])[np.isfinite(host["nint"])]))),
              float(np.max(np.abs(dev["geom_det"] / host["geom_det"] - 1))))
    h1 = sum(t[(k, 1)] for k in ("geometry", "weight", "nint")); hn = sum(t[(k, 0)] for k in ("geometry", "weight", "nint"))
    print(f"{n}x{n}: host 1 thread {h1 * 1e3:.0f} ms (geometry {t[('geometry', 1)] * 1e3:.0f} weight {t[('weight', 1)] * 1e3:.0f} "
          f"nint {t[('nint', 1)] * 1e3:.0f}) | host {os.cpu_count()} threads {hn * 1e3:.0f} ms | device kernels {dev['device_ms']:.3f} ms, "
          f"call incl. H2D/D2H + cudaMalloc {min(w) * 1e3:.1f} ms | max |diff| vs host {err:.2e}", flush=True)
