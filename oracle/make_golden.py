#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference CPU build (oracle/_ref, `make -C oracle ref`).

Runs only in the build container (needs /root/reference to have been compiled into oracle/_ref/).  The
fixtures it writes are committed; tests never read /root/reference.

  python oracle/make_golden.py functions     # function-level vectors on a 48x48 synthetic dump
  python oracle/make_golden.py samplers      # empirical quantiles of the reference's samplers
  python oracle/make_golden.py spectrum      # 8-seed reference spectra on the 192x192 dump (minutes)
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refharness as rh  # noqa: E402
from tools import make_harm_dump  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SMALL = dict(n0=48, n1=48, photon_n=2000, mass_unit=4e19)


def small_dump(path):
    header, table = make_harm_dump.make_dump(n0=SMALL["n0"], n1=SMALL["n1"])
    make_harm_dump.write_dump(path, header, table)


def null_k(R, rng, x):
    """random future-directed null wave-vector at x (k.k = 0 in the reference's own metric)"""
    g = R.gcov(x)
    ksp = rng.normal(size=3) * np.array([1.0, 0.3, 0.3])
    # solve g00 k0^2 + 2 g0i k0 ki + gij ki kj = 0 for k0
    a = g[0, 0]
    b = 2.0 * (g[0, 1:] @ ksp)
    c = ksp @ g[1:, 1:] @ ksp
    disc = b * b - 4 * a * c
    k0 = (-b - np.sqrt(disc)) / (2 * a) if a < 0 else (-c / b)
    k = np.array([k0, *ksp])
    if k0 < 0:
        k = -k
    return k * 10 ** rng.uniform(-9, -3)


def gen_functions():
    rng = np.random.default_rng(20261018)
    tmp = tempfile.mkdtemp()
    dump = os.path.join(tmp, "dump48.txt")
    small_dump(dump)
    R = rh.Ref(dump, SMALL["photon_n"], SMALL["mass_unit"], seed=123)
    d = R.model_dict()
    out = {("model_" + k): np.asarray(v) for k, v in d.items()}
    x1lo, x1hi = d["x_start1"], np.log(100.0)

    # ---- geometry ----
    n = 256
    X = np.zeros((n, 4))
    X[:, 1] = rng.uniform(x1lo - 0.05, x1hi, n)
    X[:, 2] = rng.uniform(1e-3, 1 - 1e-3, n)
    X[:8, 2] = [1e-6, 1e-4, 0.5, 0.999999, 0.25, 0.75, 0.01, 0.99]
    X[:, 3] = rng.uniform(0, 2 * np.pi, n)
    out["geom_x"] = X
    out["geom_gcov"] = np.array([R.gcov(x) for x in X])
    out["geom_gcon"] = np.array([R.gcon(x) for x in X])
    out["geom_conn"] = np.array([R.connection(x) for x in X])
    K = np.array([null_k(R, rng, x) for x in X])
    out["geom_k"] = K
    out["geom_dkdlam"] = np.array([R.init_dkdlam(x, k) for x, k in zip(X, K)])
    out["geom_step"] = np.array([R.step_size(x, k) for x, k in zip(X, K)])

    # ---- single push_photon calls ----
    P0 = np.zeros((n, 25))
    P1 = np.zeros((n, 25))
    DL = np.zeros(n)
    for t in range(n):
        x, k = X[t], K[t]
        g = R.gcov(x)
        f = np.zeros(25)
        f[0:4], f[4:8] = x, k
        f[8:12] = R.init_dkdlam(x, k)
        f[12] = 1e40
        f[23] = -(k @ g[0])  # e_0_s
        f[13] = f[23]
        scale = 1.0 if t % 4 else 6.0  # every 4th: oversized step to force the halving branch
        DL[t] = R.step_size(x, k) * scale
        P0[t] = f
        P1[t] = R.push_photon(f, DL[t])
    out["push_in"], out["push_dl"], out["push_out"] = P0, DL, P1

    # ---- vacuum trajectories: step_size + push_photon until stop or 400 steps ----
    ntraj, nsteps, stride = 24, 400, 20
    T0 = np.zeros((ntraj, 25))
    TR = np.full((ntraj, nsteps // stride, 9), np.nan)  # x[4] k[4] e_0_s every `stride` steps
    for t in range(ntraj):
        x = np.array([0.0, rng.uniform(np.log(2.0), np.log(30.0)), rng.uniform(0.1, 0.9), 0.0])
        k = null_k(R, rng, x)
        g = R.gcov(x)
        f = np.zeros(25)
        f[0:4], f[4:8] = x, k
        f[8:12] = R.init_dkdlam(x, k)
        f[12] = 1e40
        f[23] = f[13] = -(k @ g[0])
        T0[t] = f
        for s in range(nsteps):
            if f[1] < d["x1_min"] or f[1] > x1hi:
                break
            dl = R.step_size(f[0:4], f[4:8])
            f = R.push_photon(f, dl)
            if (s + 1) % stride == 0:
                TR[t, s // stride, 0:4] = f[0:4]
                TR[t, s // stride, 4:8] = f[4:8]
                TR[t, s // stride, 8] = f[23]
    out["traj_in"], out["traj_out"] = T0, TR
    out["traj_meta"] = np.array([nsteps, stride])

    # ---- fluid ----
    nf = 400
    XF = np.zeros((nf, 4))
    XF[:, 1] = rng.uniform(x1lo - 0.02, d["x_stop1"] + 0.05, nf)
    XF[:, 2] = rng.uniform(-0.01, 1.01, nf)
    XF[:6, 1] = [x1lo + 1e-9, x1lo + 0.4 * d["dx1"], d["x_stop1"] - 1e-9, d["x_stop1"] - 0.3 * d["dx1"],
                 0.5 * (x1lo + d["x_stop1"]), x1lo + 3.5 * d["dx1"]]
    XF[:6, 2] = [0.5, 0.2 * d["dx2"], 1 - 0.2 * d["dx2"], 0.5, 1e-9, 1 - 1e-9]
    out["fluid_x"] = XF
    out["fluid_params"] = np.array([R.fluid_params(x) for x in XF])
    zi, zj = np.meshgrid(np.arange(0, 48, 5), np.arange(0, 48, 3), indexing="ij")
    zij = np.stack([zi.ravel(), zj.ravel()], 1)
    out["zone_ij"] = zij
    out["zone_fluid"] = np.array([R.fluid_zone(int(i), int(j)) for i, j in zij])
    out["zone_init_all"] = np.array([[R.init_zone(i, j) for j in range(48)] for i in range(48)])

    # ---- radiation ----
    nr = 400
    nu = 10 ** rng.uniform(8, 22, nr)
    te = 10 ** rng.uniform(-1.0, 3.0, nr)
    te[:20] = 10 ** rng.uniform(-5, -0.6, 20)  # below theta_e_min / hotcross min_t branches
    ne = 10 ** rng.uniform(2, 9, nr)
    bb = 10 ** rng.uniform(-2, 4, nr)
    th = rng.uniform(0.01, np.pi - 0.01, nr)
    out["rad_args"] = np.stack([nu, te, ne, bb, th], 1)
    L = R.L
    out["rad_alpha_scatt"] = np.array([L.ref_alpha_inv_scatt(a, b, c) for a, b, c in zip(nu, te, ne)])
    out["rad_alpha_abs"] = np.array([L.ref_alpha_inv_abs(a, b, c, e, f) for a, b, c, e, f in zip(nu, te, ne, bb, th)])
    out["rad_synch"] = np.array([L.ref_synch(a, c, b, e, f) for a, b, c, e, f in zip(nu, te, ne, bb, th)])
    out["rad_k2"] = np.array([L.ref_k2_eval(b) for b in te])
    out["rad_f"] = np.array([L.ref_f_eval(b, e, a) for a, b, e in zip(nu, te, bb)])
    w = 10 ** rng.uniform(-11.5, 5.5, nr)
    tt = 10 ** rng.uniform(-4.5, 3.9, nr)
    out["hc_args"] = np.stack([w, tt], 1)
    out["hc_lkup"] = np.array([L.ref_hotcross_lkup(a, b) for a, b in zip(w, tt)])
    # angles / frequencies with real fluid states
    inside = out["fluid_params"][:, 0] > 0
    XA = XF[inside][:128]
    FP = out["fluid_params"][inside][:128]
    KA = np.array([null_k(R, rng, x) for x in XA])
    out["ang_x"], out["ang_k"], out["ang_fluid"] = XA, KA, FP
    out["ang_theta"] = np.array([R.bk_angle(x, k, fp[7:11], fp[15:19], fp[2]) for x, k, fp in zip(XA, KA, FP)])
    out["ang_nu"] = np.array([R.fluid_nu(x, k, fp[7:11]) for x, k, fp in zip(XA, KA, FP)])

    # ---- bias ----
    R.L.ref_set_bias_stats(3.3e-3, 777, 1000)
    tb = 10 ** rng.uniform(-1, 2.5, 64)
    wb = 10 ** rng.uniform(30, 40, 64)
    out["bias_stats"] = np.array([3.3e-3, 777, 1000])
    out["bias_args"] = np.stack([tb, wb], 1)
    out["bias_out"] = np.array([L.ref_bias_func(a, b) for a, b in zip(tb, wb)])

    # ---- tetrads ----
    ET = []
    for x, fp in zip(XA[:64], FP[:64]):
        g = R.gcov(x)
        bhat = fp[11:15] / (fp[2] / d["b_unit"])
        ec, ev = R.make_tetrad(fp[3:7], bhat, g)
        ET.append(np.concatenate([g.ravel(), fp[3:7], bhat, ec.ravel(), ev.ravel()]))
    out["tetrad"] = np.array(ET)

    # ---- whole-photon tracks that are RNG independent (no scattering, no roulette) ----
    # bias statistics pinned so that bias = 1 (max_tau_scatt huge); photons whose result differs between two
    # reference RNG seeds consumed a decision-relevant random number and are dropped.
    R.L.ref_set_bias_stats(1e30, 0, 0)
    zinit = out["zone_init_all"]
    zones = [(i, j) for i in range(48) for j in range(48) if zinit[i, j, 0] > 0]
    pick = rng.choice(len(zones), size=min(96, len(zones)), replace=False)
    starts = []
    R.L.ref_rng_init(1)
    for zidx in pick:
        i, j = zones[zidx]
        ips = R.sample_zone_photons(i, j, zinit[i, j, 1], 3)
        starts += [rh.init_to_flat(ip) for ip in ips]
    starts = np.array(starts)
    R.L.ref_rng_init(11)
    endA = np.array([R.track(s) for s in starts])
    R.L.ref_rng_init(12)
    endB = np.array([R.track(s) for s in starts])
    same = np.all((endA == endB) | (np.isnan(endA) & np.isnan(endB)), axis=1)
    out["track_bias_stats"] = np.array([1e30, 0, 0])
    out["track_in"], out["track_out"] = starts[same], endA[same]
    print(f"track: kept {same.sum()} of {len(same)} RNG-independent photons", file=sys.stderr)
    # record: spectrum produced by recording the kept end states that escaped
    R.L.ref_clear_spectrum()
    R.L.ref_set_bias_stats(1e30, 0, 0)
    esc = out["track_out"][:, 1] > np.log(100.0)
    for f in out["track_out"][esc]:
        ff = np.ascontiguousarray(f)
        R.L.ref_record_super_photon(ff.ctypes.data_as(rh.dp))
    out["record_spectrum"] = R.spectrum()
    out["record_counters"] = R.counters()

    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "functions_48.npz"), **out)
    print("wrote functions_48.npz", {k: v.shape for k, v in out.items() if not k.startswith("model_")},
          file=sys.stderr)


def quantiles(v, nq=501):
    return np.quantile(np.asarray(v), np.linspace(0, 1, nq))


def gen_samplers():
    tmp = tempfile.mkdtemp()
    dump = os.path.join(tmp, "dump48.txt")
    small_dump(dump)
    R = rh.Ref(dump, SMALL["photon_n"], SMALL["mass_unit"], seed=4242)
    L = R.L
    N = 200000
    out = {"n_samples": np.array(N)}
    for dof in (3, 4, 5, 6):
        out[f"chi_sq_{dof}"] = quantiles([L.ref_chi_sq(dof) for _ in range(N)])
    for te in (0.5, 3.0, 30.0):
        out[f"y_{te}"] = quantiles([L.ref_sample_y(te) for _ in range(N)])
    for beta in (0.3, 0.9, 0.9999):
        out[f"mu_{beta}"] = quantiles([L.ref_sample_mu(beta) for _ in range(N)])
    for k0 in (0.01, 1.0, 30.0):
        out[f"kn_{k0}"] = quantiles([L.ref_sample_kn(k0) for _ in range(N)])
    out["thomson"] = quantiles([L.ref_sample_thomson() for _ in range(N)])
    # electron sampling: gamma and cosine between p and k, for a soft and a hard photon
    for k0, te in ((1e-6, 5.0), (2.0, 5.0), (1e-3, 50.0)):
        k = np.array([k0, k0, 0.0, 0.0])
        P = np.array([R.sample_electron(k, te) for _ in range(N // 4)])
        out[f"el_gamma_{k0}_{te}"] = quantiles(P[:, 0])
        out[f"el_mu_{k0}_{te}"] = quantiles(P[:, 1] / np.sqrt((P[:, 1:] ** 2).sum(1)))
        # full scatter in the tetrad frame: energy ratio and deflection cosine
        KP = np.array([R.sample_scattered_photon(k, p) for p in P])
        out[f"sc_eratio_{k0}_{te}"] = quantiles(KP[:, 0] / k0)
        out[f"sc_cos_{k0}_{te}"] = quantiles(KP[:, 1] / KP[:, 0])
    np.savez_compressed(os.path.join(GOLD, "samplers.npz"), **out)
    print("wrote samplers.npz", file=sys.stderr)


def gen_spectrum(photon_n=100000, seeds=8, mass_units=(4e19,), first_seed=0, merge=False, suffix=""):
    """`seeds` runs of the reference CLI (seed 123 + first_seed + s), each a complete run at `photon_n`
    (BASELINE.json configs[0]); merge=True appends to the existing fixture (more seeds = less Monte Carlo
    noise in the reference ensemble)."""
    tmp = tempfile.mkdtemp()
    dump = os.path.join(tmp, "dump192.txt")
    header, table = make_harm_dump.make_dump()
    make_harm_dump.write_dump(dump, header, table)
    for mu in mass_units:
        procs = []
        for s in range(first_seed, first_seed + seeds):
            sb = os.path.join(tmp, f"spec_{s}.bin")
            cmd = [rh.CLI_PATH, "--harm_dump_path", dump, "--photon_n", str(photon_n), "--mass_unit", repr(mu),
                   "--seed", str(123 + s), "--hotcross_cache", rh.HOTCROSS_CACHE, "--spectrum_bin", sb]
            procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True), sb))
        specs, metas = [], []
        for p, sb in procs:
            o, _ = p.communicate()
            metas.append(json.loads(o.strip().splitlines()[-1]))
            specs.append(np.fromfile(sb).reshape(6, 200, 13))
        specs = np.array(specs)
        out = dict(photon_n=np.array(photon_n), mass_unit=np.array(mu), seeds=np.array(first_seed + seeds),
                   created=np.array([m["created"] for m in metas]),
                   scattered=np.array([m["scattered"] for m in metas]),
                   recorded=np.array([m["recorded"] for m in metas]),
                   run_s=np.array([m["run_s"] for m in metas]),
                   max_tau_scatt=np.array([m["max_tau_scatt"] for m in metas]),
                   # per-seed: dn_dle, de_dle, nph, nscatt, tau_abs, tau_scatt (fields 0,1,2,3,7,8)
                   spec=specs[:, :, :, [0, 1, 2, 3, 7, 8]].astype(np.float64))
        name = f"spectrum_192_{mu:.0e}{suffix}.npz".replace("+", "")
        if merge and os.path.exists(os.path.join(GOLD, name)):
            old = dict(np.load(os.path.join(GOLD, name)))
            assert int(old["photon_n"]) == photon_n and int(old["seeds"]) == first_seed
            for k in ("created", "scattered", "recorded", "run_s", "max_tau_scatt", "spec"):
                out[k] = np.concatenate([old[k], out[k]])
        np.savez_compressed(os.path.join(GOLD, name), **out)
        print("wrote", name, "rates", out["created"] / out["run_s"], file=sys.stderr)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "functions"
    if what == "functions":
        gen_functions()
    elif what == "samplers":
        gen_samplers()
    elif what == "spectrum":
        gen_spectrum()
    elif what in ("spectrum_1e6", "spectrum_1e6_more", "spectrum_1e6_more2", "spectrum_1e6_more3"):
        pass  # handled at the end of the file
    elif what == "spectrum_4e20":  # configs[3]: Compton-dominated regime, 8 seeds at photon_n = 2e4
        gen_spectrum(photon_n=20000, seeds=8, mass_units=(4e20,))
    elif what == "spectrum_more":  # spectrum_more <first_seed> <n> [mass_unit [photon_n]]
        gen_spectrum(first_seed=int(sys.argv[2]), seeds=int(sys.argv[3]), merge=True,
                     mass_units=(float(sys.argv[4]),) if len(sys.argv) > 4 else (4e19,),
                     photon_n=int(float(sys.argv[5])) if len(sys.argv) > 5 else 100000)
    elif what == "spectrum_4e20_1e5":  # configs[3] at the photon_n of configs[0]: spectrum_4e20_1e5 <first_seed> <n>
        gen_spectrum(photon_n=100000, first_seed=int(sys.argv[2]), seeds=int(sys.argv[3]), merge=True,
                     mass_units=(4e20,), suffix="_1e5")
    elif what in ("spectrum_file", "spectrum_grid", "functions_grid"):
        pass  # handled at the end of the file
    else:
        raise SystemExit(__doc__)


def gen_spectrum_big(photon_n=1000000, seeds=6, mu=4e19, first_seed=0, name="spectrum_192_4e19_1e6.npz"):
    """configs[1]: complete reference runs at the bench's photon_n (about 40 minutes per run on one core);
    only counters and theta-summed spectra are kept (small fixture).  `first_seed` / `name`: further seeds of the same
    ensemble into a second file (`spectrum_1e6_more`: seeds 506..513)."""
    tmp = tempfile.mkdtemp()
    dump = os.path.join(tmp, "dump192.txt")
    make_harm_dump.write_dump(dump, *make_harm_dump.make_dump())
    procs = []
    for s in range(first_seed, first_seed + seeds):
        sb = os.path.join(tmp, f"spec_{s}.bin")
        cmd = [rh.CLI_PATH, "--harm_dump_path", dump, "--photon_n", str(photon_n), "--mass_unit", repr(mu),
               "--seed", str(500 + s), "--hotcross_cache", rh.HOTCROSS_CACHE, "--spectrum_bin", sb]
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True), sb))
    metas, specs = [], []
    for i, (p, sb) in enumerate(procs):
        o, _ = p.communicate()
        line = next((ln for ln in reversed(o.splitlines()) if ln.startswith("{")), None)
        if p.returncode != 0 or line is None or not os.path.exists(sb):
            # the reference has undefined behaviour on some photon paths: a run that dies is
            # dropped and reported, the ensemble is made of the runs that completed
            print(f"seed {500 + first_seed + i}: reference process ended with rc={p.returncode}, no result: dropped",
                  file=sys.stderr)
            continue
        metas.append(json.loads(line))
        specs.append(np.fromfile(sb).reshape(6, 200, 13)[:, :, [0, 1, 2, 3]])
    np.savez_compressed(os.path.join(GOLD, name), photon_n=np.array(photon_n), first_seed=np.array(500 + first_seed),
                        mass_unit=np.array(mu), created=np.array([m["created"] for m in metas]),
                        scattered=np.array([m["scattered"] for m in metas]),
                        recorded=np.array([m["recorded"] for m in metas]),
                        max_tau_scatt=np.array([m["max_tau_scatt"] for m in metas]),
                        run_s=np.array([m["run_s"] for m in metas]), spec=np.array(specs))
    print("wrote", name, [m["run_s"] for m in metas], file=sys.stderr)


def gen_spectrum_file():
    """fixture for the spectrum-writer test: a Spectrum array and the bytes the reference writes for it"""
    tmp = tempfile.mkdtemp()
    dump = os.path.join(tmp, "dump48.txt")
    small_dump(dump)
    R = rh.Ref(dump, SMALL["photon_n"], SMALL["mass_unit"], seed=5)
    R.L.ref_run_simulation()
    spec = R.spectrum()
    out = os.path.join(tmp, "spectrum.txt")
    R.L.ref_report_spectrum(out.encode())
    text = np.frombuffer(open(out, "rb").read(), dtype=np.uint8)
    # luminosity as the reference logs it (harm_model.cpp:461)
    lines = open(out).read().splitlines()
    d = R.model_dict()
    dx2 = (d["x_stop2"] - d["x_start2"]) / 12.0
    lum = 0.0
    for ln in lines:
        v = [float(t) for t in ln.split()]
        for j in range(6):
            lum += v[1 + 6 * j] * 2.0 * R.L.ref_d_omega(j * dx2, (j + 1) * dx2) * 0.25
    np.savez_compressed(os.path.join(GOLD, "spectrum_file.npz"), spectrum=spec, text=text, luminosity=np.array(lum))
    print("wrote spectrum_file.npz", len(text), "bytes, L =", lum, file=sys.stderr)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "spectrum_file":
    gen_spectrum_file()
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "spectrum_1e6":
    gen_spectrum_big()  # configs[1] (the bench workload): 6 complete reference runs at photon_n = 1e6
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "spectrum_1e6_more":
    # eight further seeds (506..513) of the configs[1] ensemble, kept in their own file until a GPU run has confirmed
    # the parity test against the enlarged ensemble
    gen_spectrum_big(seeds=8, first_seed=6, name="spectrum_192_4e19_1e6_more.npz")


def _grid_dump(n, tmp=None):
    """the deterministic synthetic dump at n x n, cached under /tmp (0.9 GB of text at 1024^2)"""
    p = os.path.join(tmp or tempfile.gettempdir(), f"grmonty_golden_dump_{n}.txt")
    if not os.path.exists(p):
        make_harm_dump.write_dump(p + ".tmp", *make_harm_dump.make_dump(n0=n, n1=n))
        os.replace(p + ".tmp", p)
    return p


def gen_spectrum_grid(n=1024, photon_n=20000, seeds=8, first_seed=0, mu=4e19):
    """BASELINE.json configs[4] grid (1024 x 1024): complete runs of the reference CLI at a photon_n one core finishes
    in about a minute (the grid, not photon_n, is what this config is about: get_fluid_params / x_to_ij at another
    stride, harm_model.cpp:595-671,1406-1434, and a 67 MB footprint).  Same fixture layout as gen_spectrum."""
    dump = _grid_dump(n)
    tmp = tempfile.mkdtemp()
    procs = []
    for s in range(first_seed, first_seed + seeds):
        sb = os.path.join(tmp, f"spec_{s}.bin")
        cmd = [rh.CLI_PATH, "--harm_dump_path", dump, "--photon_n", str(photon_n), "--mass_unit", repr(mu),
               "--seed", str(700 + s), "--hotcross_cache", rh.HOTCROSS_CACHE, "--spectrum_bin", sb]
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True), sb))
    specs, metas = [], []
    for p, sb in procs:
        o, _ = p.communicate()
        metas.append(json.loads(o.strip().splitlines()[-1]))
        specs.append(np.fromfile(sb).reshape(6, 200, 13))
    specs = np.array(specs)
    name = f"spectrum_{n}_{mu:.0e}.npz".replace("+", "")
    out = dict(photon_n=np.array(photon_n), mass_unit=np.array(mu), grid=np.array([n, n]), first_seed=np.array(700),
               created=np.array([m["created"] for m in metas]), scattered=np.array([m["scattered"] for m in metas]),
               recorded=np.array([m["recorded"] for m in metas]), run_s=np.array([m["run_s"] for m in metas]),
               max_tau_scatt=np.array([m["max_tau_scatt"] for m in metas]),
               spec=specs[:, :, :, [0, 1, 2, 3, 7, 8]].astype(np.float64))
    if first_seed > 0:  # further seeds of the same ensemble: append to the fixture
        old = dict(np.load(os.path.join(GOLD, name)))
        assert int(old["photon_n"]) == photon_n and len(old["created"]) == first_seed
        for k in ("created", "scattered", "recorded", "run_s", "max_tau_scatt", "spec"):
            out[k] = np.concatenate([old[k], out[k]])
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print("wrote", name, "rates", [m["created"] / m["run_s"] for m in metas], file=sys.stderr)


def gen_functions_grid(n):
    """Function-level vectors on the n x n bench dumps (192, 1024): get_fluid_params at random and edge points,
    get_fluid_zone / init_zone on a sample of zones, plus the model scalars the lookups depend on.  The grids
    themselves are not stored: tests rebuild the same dump with tools/make_harm_dump.py and load it through the
    product's own loader."""
    rng = np.random.default_rng(1000 + n)
    R = rh.Ref(_grid_dump(n), 100000, 4e19, seed=123)
    d = R.model_dict()
    out = {k: np.asarray(d[k]) for k in ("x_start1", "x_start2", "dx1", "dx2", "x_stop1", "x_stop2", "bias_norm",
                                          "n0", "n1")}
    nf = 600
    X = np.zeros((nf, 4))
    X[:, 1] = rng.uniform(d["x_start1"] - 0.02, d["x_stop1"] + 0.05, nf)
    X[:, 2] = rng.uniform(-0.01, 1.01, nf)
    # cell edges, first / last half cells, the poles and the outer boundary (x_to_ij clamps, harm_model.cpp:1406-1434)
    e1, e2 = d["x_start1"], d["x_stop1"]
    X[:12, 1] = [e1 + 1e-9, e1 + 0.4 * d["dx1"], e2 - 1e-9, e2 - 0.3 * d["dx1"], 0.5 * (e1 + e2), e1 + 3.5 * d["dx1"],
                 e1 + 0.5 * d["dx1"], e1 + 1.5 * d["dx1"] - 1e-13, e2 - 0.5 * d["dx1"], e1 + 100.5 * d["dx1"],
                 e1 + (n - 1.5) * d["dx1"], e1 + (n // 2) * d["dx1"]]
    X[:12, 2] = [0.5, 0.2 * d["dx2"], 1 - 0.2 * d["dx2"], 0.5, 1e-9, 1 - 1e-9, 0.5 * d["dx2"], 1.5 * d["dx2"],
                 1 - 0.5 * d["dx2"], (n // 2) * d["dx2"], (n - 1.5) * d["dx2"], 0.5]
    out["fluid_x"] = X
    out["fluid_params"] = np.array([R.fluid_params(x) for x in X])
    zi = rng.integers(0, n, 400)
    zj = rng.integers(0, n, 400)
    zi[:4], zj[:4] = [0, 0, n - 1, n - 1], [0, n - 1, 0, n - 1]
    out["zone_ij"] = np.stack([zi, zj], 1)
    out["zone_fluid"] = np.array([R.fluid_zone(int(i), int(j)) for i, j in zip(zi, zj)])
    out["zone_init"] = np.array([R.init_zone(int(i), int(j)) for i, j in zip(zi, zj)])
    np.savez_compressed(os.path.join(GOLD, f"functions_grid_{n}.npz"), **out)
    print(f"wrote functions_grid_{n}.npz", file=sys.stderr)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "spectrum_grid":
    # spectrum_grid [n] [photon_n] [seeds] [first_seed]: configs[4] (1024 x 1024) reference ensemble
    gen_spectrum_grid(*(int(float(v)) for v in sys.argv[2:6]))
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "functions_grid":
    gen_functions_grid(int(sys.argv[2]))  # one model per process: functions_grid 192, then functions_grid 1024
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "spectrum_1e6_more2":
    # six further seeds (514..519): 20 complete reference runs at the bench's photon_n in total
    gen_spectrum_big(seeds=6, first_seed=14, name="spectrum_192_4e19_1e6_more2.npz")
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "spectrum_1e6_more3":
    # seven further seeds (520..526): 27 complete reference runs at the bench's photon_n in total
    gen_spectrum_big(seeds=7, first_seed=20, name="spectrum_192_4e19_1e6_more3.npz")
