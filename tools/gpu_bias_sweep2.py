#!/usr/bin/env python
"""Round 2: recorded / scattered counts of the CUDA path for several generation schedules against the reference
ensembles at photon_n = 1e5 in BOTH regimes -- M_unit = 4e19 (configs[0], 60 reference runs) and 4e20 (configs[3],
21 runs) -- and their run time.  usage (GPU box): tools/gpu_bias_sweep2.py [n_seeds]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 32
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
refs = {4e19: dict(np.load(os.path.join(ROOT, "tests/golden/spectrum_192_4e19.npz"))),
        4e20: dict(np.load(os.path.join(ROOT, "tests/golden/spectrum_192_4e20_1e5.npz")))}
models = {}
for mu, ref in refs.items():
    hm = gm.HarmModel(int(ref["photon_n"]), mu); hm.read_file(p); hm.init()
    models[mu] = hm.model_dict()
    print("reference %g: recorded %.0f (sd %.2f%%) scattered %.0f (sd %.2f%%) n=%d" % (
        mu, ref["recorded"].mean(), 100 * ref["recorded"].std(ddof=1) / ref["recorded"].mean(), ref["scattered"].mean(),
        100 * ref["scattered"].std(ddof=1) / ref["scattered"].mean(), len(ref["recorded"])), flush=True)
configs = [dict(gen_fine_div=6, gen_budget=384), dict(gen_fine_div=8, gen_budget=384), dict(gen_fine_div=8, gen_budget=512),
           dict(gen_fine_div=6, gen_budget=512), dict(gen_fine_div=12, gen_budget=384), dict(gen_fine_div=12, gen_budget=512),
           dict(gen_fine_div=16, gen_budget=768), dict(gen_fine_div=8, gen_budget=384, gen_ramp=4),
           dict(gen_fine_div=6, gen_budget=384, gen_fine_from=4096)]
for cfg in configs:
    row = dict(cfg=cfg)
    for mu, ref in refs.items():
        rec, scat, lum, ms = [], [], [], []
        for s in range(n_seeds):
            c = gm.Context(models[mu], seed=1000 + s, **cfg)
            t0 = time.time(); c.run(); dt = time.time() - t0
            r = c.result(); c.close()
            rec.append(r["recorded"]); scat.append(r["scattered"]); lum.append(r["spectrum"][:, :, 1].sum()); ms.append(dt * 1e3)
        rec, scat, lum = np.array(rec, float), np.array(scat, float), np.array(lum)
        r_rec, r_scat, r_lum = ref["recorded"].astype(float), ref["scattered"].astype(float), ref["spec"][..., 1].sum(axis=(1, 2))
        se = lambda g, r: 100 * np.hypot(g.std(ddof=1) / np.sqrt(len(g)) / g.mean(), r.std(ddof=1) / np.sqrt(len(r)) / r.mean())
        tag = "%g" % mu
        row[tag] = dict(d_rec=round(100 * (rec.mean() / r_rec.mean() - 1), 2), se_rec=round(se(rec, r_rec), 2),
                        d_scat=round(100 * (scat.mean() / r_scat.mean() - 1), 2), se_scat=round(se(scat, r_scat), 2),
                        d_lum=round(100 * (lum.mean() / r_lum.mean() - 1), 2), se_lum=round(se(lum, r_lum), 2), ms=round(float(np.median(ms)), 1))
    print(json.dumps(row), flush=True)
