"""North-star parity on BASELINE.json configs[0]: the CUDA path against the UNMODIFIED reference CPU build on the
synthetic dump019-shaped dump (192x192, a = 0.9375), photon_n = 1e5, M_unit = 4e19.

Reference side: tests/golden/spectrum_192_4e19.npz -- complete runs of the reference CLI (oracle/_ref/grmonty_ref,
mt19937 seeds 123, 124, ...) written by oracle/make_golden.py.  CUDA side: complete runs through the C ABI with
Philox seeds 1000, 1001, ...  The two use different random streams, so the comparison is statistical:

  * per-bin nu L_nu (the de_dle accumulator, reference harm_model.cpp:1324) chi-square consistent over the bins
    holding >= 1e3 superphotons per run, with the bin variances measured from the seed-to-seed spread;
  * L1 distance of the ensemble-mean spectra over those bins < 2 %;
  * integrated luminosity, recorded and scattered counts within 1 %, as HARD bars: the ensembles are large enough
    (64 CUDA seeds against 60 reference runs) that the standard error of every difference is below 0.5 %, which the
    test asserts too.  (The reference's own seed-to-seed spread of the counts is 2.5 % at this photon_n, because its
    scattering bias divides by a running maximum, harm_model.cpp:1296,1391-1404.)

configs[1] (photon_n = 1e6, the bench workload) is compared the same way against all 14 complete reference runs
(two fixture files); there the counts are compared CONDITIONAL on the running maximum they are driven by, see
test_bench_workload_photon_n_1e6_vs_reference.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_GPU_SEEDS = 64


def report(section, values):
    """measured parity numbers -> the JSON file named by GRMONTY_B200_PARITY_REPORT, if set (tools/gpu_r2_check.sh sets
    it; the file is copied to profiles/ for the record).  Without the variable the test writes nothing."""
    import json
    path = os.environ.get("GRMONTY_B200_PARITY_REPORT")
    if not path:
        return
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[section] = values
    json.dump(data, open(path, "w"), indent=1)


FIELDS = [0, 1, 2, 3, 7, 8]  # dn_dle de_dle nph nscatt tau_abs tau_scatt: the fields stored in the fixture


@pytest.fixture(scope="module")
def ref():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "spectrum_192_4e19.npz")))


@pytest.fixture(scope="module")
def gpu_runs(ref, tmp_path_factory):
    import cuda_grmonty_b200 as gm
    from tools import make_harm_dump
    dump = str(tmp_path_factory.mktemp("dump") / "dump192.txt")
    make_harm_dump.write_dump(dump, *make_harm_dump.make_dump(n0=192, n1=192))
    hm = gm.HarmModel(int(ref["photon_n"]), float(ref["mass_unit"]))
    hm.read_file(dump)
    hm.init()
    model = hm.model_dict()
    runs = []
    for s in range(N_GPU_SEEDS):
        ctx = gm.Context(model, seed=1000 + s)
        ctx.run()
        r = ctx.result()
        ctx.close()
        runs.append(r)
    return runs


def test_counts_and_luminosity_within_1pct(ref, gpu_runs):
    g_created = np.array([r["created"] for r in gpu_runs], float)
    g_rec = np.array([r["recorded"] for r in gpu_runs], float)
    g_scat = np.array([r["scattered"] for r in gpu_runs], float)
    g_lum = np.array([r["spectrum"][:, :, 1].sum() for r in gpu_runs])
    r_created, r_rec, r_scat = (ref[k].astype(float) for k in ("created", "recorded", "scattered"))
    r_lum = ref["spec"][..., 1].sum(axis=(1, 2))
    # primaries: the per-zone expectation is the same number in both codes (stochastic rounding only)
    assert abs(g_created.mean() / r_created.mean() - 1) < 1e-3
    rep = {}
    for name, g, r in (("luminosity", g_lum, r_lum), ("recorded", g_rec, r_rec), ("scattered", g_scat, r_scat)):
        d = g.mean() / r.mean() - 1
        se = np.hypot(g.std(ddof=1) / np.sqrt(len(g)) / g.mean(), r.std(ddof=1) / np.sqrt(len(r)) / r.mean())
        rep[name] = (d, se)
    print("relative difference of ensemble means (difference, standard error):", rep)
    report("counts_configs0", {k: {"rel_diff": float(v[0]), "std_err": float(v[1])} for k, v in rep.items()} |
           {"n_gpu_seeds": len(gpu_runs), "n_ref_seeds": int(len(r_rec))})
    for name, (d, se) in rep.items():
        assert se < 0.005, (name, se)       # the comparison has the power to see a 1 % difference at two sigma
        assert abs(d) < 0.01, (name, d, se)  # north-star bar, no statistical allowance


def test_spectrum_chi_square_and_l1(ref, gpu_runs):
    g = np.array([r["spectrum"][:, :, FIELDS] for r in gpu_runs])  # [Ng][6][200][6]
    r = ref["spec"]                                                 # [Nr][6][200][6]
    ng, nr = len(g), len(r)
    mask = r[..., 2].mean(0) >= 1e3
    assert mask.sum() > 300
    gm_, rm = g[..., 1].mean(0), r[..., 1].mean(0)
    var = g[..., 1].var(0, ddof=1) / ng + r[..., 1].var(0, ddof=1) / nr
    z = (gm_ - rm)[mask] / np.sqrt(var[mask])
    chi2 = float((z ** 2).mean())
    l1 = float(np.abs(gm_ - rm)[mask].sum() / rm[mask].sum())
    # the same statistic for reference-vs-reference (half the seeds against the other half): the noise floor
    a, b = r[: nr // 2, ..., 1], r[nr // 2:, ..., 1]
    l1_floor = float(np.abs(a.mean(0) - b.mean(0))[mask].sum() / rm[mask].sum())
    print(f"bins {int(mask.sum())}  chi2/bin {chi2:.3f}  L1 {l1:.4f}  (reference half-vs-half L1 {l1_floor:.4f})  "
          f"max |z| {np.abs(z).max():.2f}")
    report("spectrum_configs0", {"bins": int(mask.sum()), "chi2_per_bin": chi2, "l1": l1, "l1_ref_half_vs_half": l1_floor,
                                 "max_abs_z": float(np.abs(z).max())})
    # variances estimated from ~8-24 samples make z Student-t like: E[z^2] ~ 1.1-1.4 for identical distributions
    assert chi2 < 1.6, chi2
    assert np.abs(z).max() < 6.0
    assert l1 < 0.02, l1
    # photon-number spectrum and the scattering-depth moments follow the same distribution as well
    for fld, tol in ((0, 0.02), (2, 0.02)):
        a, b = g[..., fld].mean(0), r[..., fld].mean(0)
        assert np.abs(a - b)[mask].sum() / b[mask].sum() < tol, fld


def test_run_is_deterministic(gpu_runs, ref):
    """same seed, same launch geometry => identical integer counters and (up to atomic summation order) spectrum"""
    import cuda_grmonty_b200 as gm
    from tools import make_harm_dump
    import tempfile
    dump = os.path.join(tempfile.mkdtemp(), "dump192.txt")
    make_harm_dump.write_dump(dump, *make_harm_dump.make_dump(n0=192, n1=192))
    hm = gm.HarmModel(int(ref["photon_n"]), float(ref["mass_unit"]))
    hm.read_file(dump)
    hm.init()
    ctx = gm.Context(hm.model_dict(), seed=1000)
    ctx.run()
    r = ctx.result()
    ctx.close()
    r0 = gpu_runs[0]
    assert (r["created"], r["recorded"], r["scattered"]) == (r0["created"], r0["recorded"], r0["scattered"])
    assert np.array_equal(r["spectrum"][:, :, 2], r0["spectrum"][:, :, 2])
    assert np.allclose(r["spectrum"][:, :, 1], r0["spectrum"][:, :, 1], rtol=1e-9, atol=0)


def conditional_residual(g_tau, g_val, r_tau, r_val, n_boot=2000):
    """Counts at a given value of the running maximum they are driven by.  ln(count) = a + b ln(max_tau_scatt) is fitted
    to the CUDA ensemble; returns the median residual of the reference runs about that line, its bootstrap standard
    error (resampling both ensembles, fixed generator) and the fitted exponent b."""
    lg_t, lg_v, lr_t, lr_v = np.log(g_tau), np.log(g_val), np.log(r_tau), np.log(r_val)

    def med(lgt, lgv, lrt, lrv):
        b, a = np.polyfit(lgt, lgv, 1)
        return float(np.median(lrv - (a + b * lrt))), float(b)
    m, b = med(lg_t, lg_v, lr_t, lr_v)
    rng = np.random.default_rng(7)
    boots = []
    for _ in range(n_boot):
        ig, ir = rng.integers(0, len(lg_t), len(lg_t)), rng.integers(0, len(lr_t), len(lr_t))
        boots.append(med(lg_t[ig], lg_v[ig], lr_t[ir], lr_v[ir])[0])
    return m, float(np.std(boots)), b


def test_bench_workload_photon_n_1e6_vs_reference():
    """configs[1] (the bench workload, photon_n = 1e6): 32 CUDA runs against all 26 complete runs of the reference CLI
    (31 - 40 minutes each on one core; tests/golden/spectrum_192_4e19_1e6.npz + ..._1e6_more.npz + ..._more2.npz + ..._more3.npz,
    written by `oracle/make_golden.py spectrum_1e6` / `spectrum_1e6_more` / `..._more2` / `..._more3`; a 27th run,
    seed 526, ended with SIGSEGV inside the reference binary after ~35 minutes and is not part of the ensemble).

    Luminosity and the spectrum are bias-independent observables and get hard bars.  The recorded / scattered COUNTS
    are not: the reference's scattering bias is ~ 1 / (running maximum of tau_scatt) (harm_model.cpp:1296,1391-1404),
    so a run whose maximum jumped early ends with fewer, heavier scattered superphotons -- one of the 26 reference runs
    ends 32 % below the others for exactly that reason, and the per-run spread of the scattered count is 10 %.  The
    CUDA path keeps the same statistic and has the same tail.  Comparing ensemble means would therefore be decided by
    whether such a run is in the sample; the counts are compared at equal max_tau_scatt instead (median residual of the
    reference runs about the CUDA ensemble's count-vs-maximum relation), which has a standard error below 0.6 %."""
    import tempfile
    import cuda_grmonty_b200 as gm
    from tools import make_harm_dump
    parts = [dict(np.load(os.path.join(ROOT, "tests", "golden", f))) for f in
             ("spectrum_192_4e19_1e6.npz", "spectrum_192_4e19_1e6_more.npz", "spectrum_192_4e19_1e6_more2.npz",
              "spectrum_192_4e19_1e6_more3.npz")]
    ref = {k: np.concatenate([p[k] for p in parts]) for k in ("created", "recorded", "scattered", "max_tau_scatt", "spec")}
    assert len(ref["recorded"]) == 26
    dump = os.path.join(tempfile.mkdtemp(), "dump192.txt")
    make_harm_dump.write_dump(dump, *make_harm_dump.make_dump(n0=192, n1=192))
    hm = gm.HarmModel(int(parts[0]["photon_n"]), float(parts[0]["mass_unit"]))
    hm.read_file(dump)
    hm.init()
    model = hm.model_dict()
    runs = []
    for s in range(32):
        ctx = gm.Context(model, seed=4000 + s)
        ctx.run()
        runs.append(ctx.result())
        ctx.close()
    rep = {}
    g_lum = np.array([r["spectrum"][:, :, 1].sum() for r in runs])
    r_lum = ref["spec"][..., 1].sum(axis=(1, 2))
    g_tau = np.array([r["max_tau_scatt"] for r in runs])
    for name, g, r in (("luminosity", g_lum, r_lum),
                       ("recorded", np.array([r["recorded"] for r in runs], float), ref["recorded"].astype(float)),
                       ("scattered", np.array([r["scattered"] for r in runs], float), ref["scattered"].astype(float))):
        d = g.mean() / r.mean() - 1
        se = np.hypot(g.std(ddof=1) / np.sqrt(len(g)) / g.mean(), r.std(ddof=1) / np.sqrt(len(r)) / r.mean())
        rep[name] = {"rel_diff_of_means": float(d), "std_err": float(se),
                     "rel_diff_of_medians": float(np.median(g) / np.median(r) - 1)}
        if name != "luminosity":
            m, mse, b = conditional_residual(g_tau, g, ref["max_tau_scatt"], r)
            rep[name] |= {"ref_minus_cuda_at_equal_max_tau": m, "std_err_conditional": mse, "exponent": b}
    gs = np.array([r["spectrum"][:, :, 1] for r in runs])
    rs = ref["spec"][..., 1]
    mask = ref["spec"][..., 2].mean(0) >= 1e3
    var = gs.var(0, ddof=1) / len(gs) + rs.var(0, ddof=1) / len(rs)
    z = (gs.mean(0) - rs.mean(0))[mask] / np.sqrt(var[mask])
    l1 = float(np.abs(gs.mean(0) - rs.mean(0))[mask].sum() / rs.mean(0)[mask].sum())
    rep["spectrum"] = {"bins": int(mask.sum()), "chi2_per_bin": float((z ** 2).mean()), "l1": l1,
                       "max_abs_z": float(np.abs(z).max())}
    print(rep)
    report("configs1_photon_n_1e6", rep | {"n_gpu_seeds": len(runs), "n_ref_seeds": int(len(r_lum))})
    assert abs(rep["luminosity"]["rel_diff_of_means"]) < 0.01 and rep["luminosity"]["std_err"] < 0.005
    for name in ("recorded", "scattered"):
        assert rep[name]["std_err_conditional"] < 0.007, (name, rep[name])   # bootstrap estimate, 26 reference runs
        assert abs(rep[name]["ref_minus_cuda_at_equal_max_tau"]) < 0.01, (name, rep[name])
        # the unconditional means are heavy-tailed (see above): reported, and held to the bar within their own error
        assert abs(rep[name]["rel_diff_of_means"]) < 0.01 + 2 * rep[name]["std_err"], (name, rep[name])
    assert rep["spectrum"]["chi2_per_bin"] < 1.6          # variances from 32 + 26 samples
    assert l1 < 0.02
