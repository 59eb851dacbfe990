"""The plain-C oracle (oracle/grmonty_oracle.c) against golden vectors produced by the UNMODIFIED reference
CPU build (oracle/make_golden.py -> tests/golden/functions_48.npz).  This is what pins the oracle.

Tolerances: the restatement follows the reference expression by expression and is built without FMA
contraction, so results agree to a few ulps; 1e-13 relative (to the largest component) is used throughout.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import orc

TOL = 1e-13


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    scale = np.nanmax(np.abs(b)) + 1e-300
    return np.nanmax(np.abs(a - b)) / scale


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    L = orc.lib()
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        L.orc_philox4x32_10(c, k, o)
        assert tuple(o) == want


def test_uniform_open_interval_and_moments(orc_model):
    M = orc_model
    r = orc.OrcRng()
    M.L.orc_rng_primary(C.byref(r), 12345)
    u = np.array([M.L.orc_uniform(M.ptr, C.byref(r)) for _ in range(20000)])
    assert u.min() > 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.005
    assert r.ctr == 20000


def test_geometry(golden, orc_model):
    M = orc_model
    X, K = golden["geom_x"], golden["geom_k"]
    for t in range(len(X)):
        assert relerr(M.gcov(X[t]), golden["geom_gcov"][t]) < TOL
        assert relerr(M.gcon(X[t]), golden["geom_gcon"][t]) < TOL
        assert relerr(M.connection(X[t]), golden["geom_conn"][t]) < TOL
        assert relerr(M.init_dkdlam(X[t], K[t]), golden["geom_dkdlam"][t]) < TOL
        assert abs(M.step_size(X[t], K[t]) / golden["geom_step"][t] - 1) < TOL


def test_push_photon_single_calls(golden, orc_model):
    M = orc_model
    n_halved = 0
    for f0, dl, f1 in zip(golden["push_in"], golden["push_dl"], golden["push_out"]):
        M.m.n_push_attempts = 0
        got = M.push_photon(f0, dl)
        n_halved += M.m.n_push_attempts > 1
        for sl in (slice(0, 4), slice(4, 8), slice(8, 12), slice(23, 24)):
            assert relerr(got[sl], f1[sl]) < 1e-12
    assert n_halved > 10  # the halving branch is exercised


def test_vacuum_trajectories(golden, orc_model):
    """whole trajectories (step_size + push_photon), north-star bar: 1e-10 relative"""
    M = orc_model
    nsteps, stride = golden["traj_meta"]
    x1_max = np.log(100.0)
    for f0, tr in zip(golden["traj_in"], golden["traj_out"]):
        f = f0.copy()
        for s in range(nsteps):
            if f[1] < M.m.x1_min or f[1] > x1_max:
                break
            f = M.push_photon(f, M.step_size(f[0:4], f[4:8]))
            if (s + 1) % stride == 0:
                want = tr[s // stride]
                assert relerr(f[0:4], want[0:4]) < 1e-10
                assert relerr(f[4:8], want[4:8]) < 1e-10
                assert abs(f[23] / want[8] - 1) < 1e-10


def test_fluid(golden, orc_model):
    M = orc_model
    n_out = 0
    for x, want in zip(golden["fluid_x"], golden["fluid_params"]):
        got = M.fluid_params(x)
        if want[0] == 0.0:
            n_out += 1
            assert got[0] == 0.0
        else:
            assert relerr(got, want) < TOL
    assert n_out > 5
    for (i, j), want in zip(golden["zone_ij"], golden["zone_fluid"]):
        assert relerr(M.fluid_zone(int(i), int(j)), want) < TOL
    zi = golden["zone_init_all"]
    got = np.array([[M.init_zone(i, j) for j in range(48)] for i in range(48)])
    assert (zi[:, :, 0] > 0).sum() > 100
    assert relerr(got, zi) < TOL


def test_radiation(golden, orc_model):
    M, L = orc_model, orc_model.L
    a = golden["rad_args"]
    for t, (nu, te, ne, b, th) in enumerate(a):
        for got, want in ((L.orc_alpha_inv_scatt(M.ptr, nu, te, ne), golden["rad_alpha_scatt"][t]),
                          (L.orc_alpha_inv_abs(M.ptr, nu, te, ne, b, th), golden["rad_alpha_abs"][t]),
                          (L.orc_synch(M.ptr, nu, ne, te, b, th), golden["rad_synch"][t]),
                          (L.orc_k2_eval(M.ptr, te), golden["rad_k2"][t]),
                          (L.orc_f_eval(M.ptr, te, b, nu), golden["rad_f"][t])):
            assert got == pytest.approx(want, rel=1e-12, abs=0.0) or (got == 0.0 and want == 0.0)
    for (w, te), want in zip(golden["hc_args"], golden["hc_lkup"]):
        assert L.orc_hotcross_lkup(M.ptr, w, te) == pytest.approx(want, rel=1e-12)
    dp = orc.dp
    for k, fp, th, nu in zip(golden["ang_k"], golden["ang_fluid"], golden["ang_theta"], golden["ang_nu"]):
        kk = np.ascontiguousarray(k)
        uc, bc = np.ascontiguousarray(fp[7:11]), np.ascontiguousarray(fp[15:19])
        L.orc_bk_angle.argtypes = [C.POINTER(orc.OrcModel), dp, dp, dp, C.c_double]
        L.orc_fluid_nu.argtypes = [dp, dp]
        got_th = L.orc_bk_angle(M.ptr, kk.ctypes.data_as(dp), uc.ctypes.data_as(dp), bc.ctypes.data_as(dp), fp[2])
        got_nu = L.orc_fluid_nu(kk.ctypes.data_as(dp), uc.ctypes.data_as(dp))
        assert got_th == pytest.approx(th, rel=1e-12)
        assert got_nu == pytest.approx(nu, rel=1e-12)


def test_bias(golden, orc_model):
    M = orc_model
    mt, ns, nr = golden["bias_stats"]
    M.set_bias_stats(mt, ns, nr)
    for (te, w), want in zip(golden["bias_args"], golden["bias_out"]):
        assert M.L.orc_bias_func(M.ptr, te, w) == pytest.approx(want, rel=1e-14)


def test_permutation_is_a_bijection():
    L = orc.lib()
    for total in (1, 2, 3, 10, 97, 1000, 32240):
        mult = L.orc_perm_multiplier(total)
        idx = sorted(L.orc_permute(j, mult, total) for j in range(total))
        assert idx == list(range(total))
    # large totals: 128-bit product, still in range
    total = 16_120_000_123
    mult = L.orc_perm_multiplier(total)
    assert 0 <= L.orc_permute(total - 1, mult, total) < total


def test_tetrads(golden, orc_model):
    M = orc_model
    for row in golden["tetrad"]:
        g, u, bh = row[0:16].reshape(4, 4), row[16:20], row[20:24]
        ec, ev = M.make_tetrad(u, bh, g)
        assert relerr(ec, row[24:40].reshape(4, 4)) < 1e-12
        assert relerr(ev, row[40:56].reshape(4, 4)) < 1e-12


def test_track_rng_independent_photons(golden, orc_model):
    """Whole track_super_photon on photons whose reference result does not depend on the RNG
    (no scattering, no roulette): end state must match the reference to 1e-9."""
    M = orc_model
    mt, ns, nr = golden["track_bias_stats"]
    M.set_bias_stats(mt, ns, nr)
    M.clear()
    n_cmp = 0
    for t, (f0, f1) in enumerate(zip(golden["track_in"], golden["track_out"])):
        before = M.m.n_scatter_events
        got = M.track(f0, rng_id=(t, 0, 0))
        if M.m.n_scatter_events != before:
            continue  # the oracle's own stream produced a scattering; not comparable
        n_cmp += 1
        for sl in (slice(0, 4), slice(4, 8), slice(12, 13), slice(17, 19), slice(23, 24)):
            assert relerr(got[sl], f1[sl]) < 1e-9, (t, sl)
    assert n_cmp > 0.95 * len(golden["track_in"])


def test_record(golden, orc_model):
    M = orc_model
    M.clear()
    M.set_bias_stats(*golden["track_bias_stats"])
    M.m.acc_max_tau_scatt = golden["track_bias_stats"][0]
    esc = golden["track_out"][:, 1] > np.log(100.0)
    for f in golden["track_out"][esc]:
        ph = M.photon(f)
        M.L.orc_record_super_photon(M.ptr, C.byref(ph))
    assert M.m.acc_n_recorded == golden["record_counters"][2]
    assert relerr(M.spectrum(), golden["record_spectrum"]) < 1e-13


def test_stats_lag_knob(golden_model):
    """orc.Model.run(stats_lag=1): every generation uses the bias statistics frozen one generation earlier (the study
    knob for the generation-overlap design).  The first TWO generations then run on the initial statistics, so a run
    a run of one generation cannot depend on it; a run of several generations does."""
    from oracle import orc

    def run(last, lag, gen0, budget=384):
        M = orc.Model(golden_model, seed=7)
        M.m.acc_max_tau_scatt = float(golden_model["max_tau_scatt0"])
        M.run(0, last, 0, 1, gen0, 1 << 20, budget, stats_lag=lag)
        return int(M.m.n_created), int(M.m.acc_n_recorded), int(M.m.acc_n_scatt), M.spectrum()[:, :, 1].sum()

    # one generation without an attempt budget (nothing is carried into a drain generation): the lag cannot matter
    assert run(300, 0, 512, 0) == run(300, 1, 512, 0)
    # with the budget the suspended lineages finish in the drain generation, which is one more generation of the
    # pipeline: lag 1 runs it on the initial statistics, lag 0 on those after the first generation
    assert run(300, 0, 512)[0] == run(300, 1, 512)[0] == 300
    # several generations: same primaries, different bias history
    a, b = run(3000, 0, 64), run(3000, 1, 64)
    assert a[0] == b[0] == 3000
    assert a[1:] != b[1:]
    # (with generations this small the lag keeps the initial statistics -- a far too small max tau -- in use for twice
    # as many photons, so the counts differ by a large factor; tools/oracle_lag_study.py measures the real schedule)
