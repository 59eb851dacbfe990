#!/usr/bin/env python
"""CPU study for the generation-overlap design: what happens to the scattered / recorded
counts when a generation uses the scattering-bias statistics frozen ONE GENERATION EARLIER (oracle knob stats_lag = 1,
what a pipeline that overlaps the tail of generation g with the bulk of g+1 would see)?  Complete configs[0] runs of the
oracle (photon_n = 1e5, 1.6 M primaries, ~5 min per run and core), same seeds with and without the lag, compared with
the reference ensemble of tests/golden/spectrum_192_4e19.npz.
usage: tools/oracle_lag_study.py [n_seeds] [fine_div ...]"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MODEL = "/tmp/orc_model_192.npz"


def one(args):
    seed, lag, fine_div = args
    from oracle import orc
    d = dict(np.load(MODEL))
    model = {k: (v.item() if v.ndim == 0 else v) for k, v in d.items()}
    M = orc.Model(model, seed=seed)
    t0 = time.time()
    # BUDGET: attempts per lineage and generation before suspension (0 = never suspend, what an overlapping pipeline
    # could afford); default = the CUDA path's 384
    M.run(0, -1, 0, 1, 32, 1 << 20, fine_div=fine_div, stats_lag=lag, budget=int(os.environ.get("BUDGET", 384)))
    return dict(seed=seed, lag=lag, fine_div=fine_div, created=int(M.m.n_created), recorded=int(M.m.acc_n_recorded),
                scattered=int(M.m.acc_n_scatt), lum=float(M.spectrum()[:, :, 1].sum()), s=time.time() - t0)


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    divs = [int(a) for a in sys.argv[2:]] or [6]
    if not os.path.exists(MODEL):
        import cuda_grmonty_b200 as gm
        from tools import make_harm_dump
        p = "/tmp/orc_dump_192.txt"
        make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
        hm = gm.HarmModel(100000, 4e19)
        hm.read_file(p)
        hm.init()
        np.savez(MODEL, **{k: np.asarray(v) for k, v in hm.model_dict().items()})
    ref = np.load(os.path.join(ROOT, "tests", "golden", "spectrum_192_4e19.npz"))
    lags = [int(a) for a in os.environ.get("LAGS", "0,1").split(",")]
    jobs = [(1000 + s, lag, dv) for dv in divs for s in range(n_seeds) for lag in lags]
    with mp.get_context("spawn").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        rows = pool.map(one, jobs, chunksize=1)
    print(f"reference ensemble ({int(ref['seeds'])} runs): recorded {ref['recorded'].mean():.0f} +- "
          f"{ref['recorded'].std(ddof=1) / np.sqrt(len(ref['recorded'])):.0f}, scattered {ref['scattered'].mean():.0f} +- "
          f"{ref['scattered'].std(ddof=1) / np.sqrt(len(ref['scattered'])):.0f}")
    for dv in divs:
        for lag in lags:
            r = [x for x in rows if x["lag"] == lag and x["fine_div"] == dv]
            rec = np.array([x["recorded"] for x in r], float)
            sc = np.array([x["scattered"] for x in r], float)
            n = len(r)
            print(f"fine_div {dv} stats_lag {lag}: {n} runs, recorded {rec.mean():.0f} +- {rec.std(ddof=1) / np.sqrt(n):.0f} "
                  f"({100 * (rec.mean() / ref['recorded'].mean() - 1):+.2f} % vs reference), scattered {sc.mean():.0f} +- "
                  f"{sc.std(ddof=1) / np.sqrt(n):.0f} ({100 * (sc.mean() / ref['scattered'].mean() - 1):+.2f} %), "
                  f"{np.mean([x['s'] for x in r]):.0f} s per run")
        for x in rows:
            print("   ", x)
        if len(lags) < 2:
            continue
        a = {x["seed"]: x for x in rows if x["lag"] == 0 and x["fine_div"] == dv}
        b = {x["seed"]: x for x in rows if x["lag"] == 1 and x["fine_div"] == dv}
        for key in ("recorded", "scattered", "lum"):
            dlt = np.array([b[s][key] / a[s][key] - 1 for s in a])
            print(f"  paired lag1/lag0 - 1, {key}: {100 * dlt.mean():+.3f} % +- {100 * dlt.std(ddof=1) / np.sqrt(len(dlt)):.3f} %")


if __name__ == "__main__":
    main()
