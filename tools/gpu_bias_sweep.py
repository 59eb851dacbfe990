#!/usr/bin/env python
"""Recorded / scattered counts of the CUDA path at configs[0] for several generation schedules, against the
reference ensemble in tests/golden/spectrum_192_4e19.npz.  usage (GPU box): tools/gpu_bias_sweep.py [n_seeds]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ref = dict(np.load(os.path.join(ROOT, "tests/golden/spectrum_192_4e19.npz")))
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
hm = gm.HarmModel(int(ref["photon_n"]), 4e19); hm.read_file(p); hm.init()
model = hm.model_dict()
r_rec, r_scat, r_lum = ref["recorded"].mean(), ref["scattered"].mean(), ref["spec"][..., 1].sum(axis=(1, 2)).mean()
print("reference: recorded %.0f (sd %.2f%%) scattered %.0f (sd %.2f%%) n=%d" % (
    r_rec, 100 * ref["recorded"].std(ddof=1) / r_rec, r_scat, 100 * ref["scattered"].std(ddof=1) / r_scat, len(ref["recorded"])))
configs = [dict(gen_fine_div=6, gen_budget=384),
           dict(gen_fine_div=6, gen_budget=128, gen_budget_spread=384), dict(gen_fine_div=6, gen_budget=192, gen_budget_spread=384),
           dict(gen_fine_div=6, gen_budget=128, gen_budget_spread=256), dict(gen_fine_div=8, gen_budget=128, gen_budget_spread=384),
           dict(gen_fine_div=6, gen_budget=256, gen_budget_spread=384)]
for cfg in configs:
    rec, scat, lum, ms, gens = [], [], [], [], []
    for s in range(n_seeds):
        c = gm.Context(model, seed=1000 + s, **cfg)
        t0 = time.time(); c.run(); dt = time.time() - t0
        r = c.result(); c.close()
        rec.append(r["recorded"]); scat.append(r["scattered"]); lum.append(r["spectrum"][:, :, 1].sum()); ms.append(dt * 1e3)
        gens.append(r["stats"]["n_generations"])
    rec, scat, lum = np.array(rec, float), np.array(scat, float), np.array(lum)
    print(json.dumps(dict(cfg=cfg, d_recorded_pct=100 * (rec.mean() / r_rec - 1), d_scattered_pct=100 * (scat.mean() / r_scat - 1),
                          d_lum_pct=100 * (lum.mean() / r_lum - 1), se_scat_pct=100 * scat.std(ddof=1) / scat.mean() / np.sqrt(n_seeds),
                          sd_scat_pct=100 * scat.std(ddof=1) / scat.mean(), ms=float(np.median(ms)), gens=int(np.median(gens)))), flush=True)
