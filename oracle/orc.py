"""ctypes binding of the plain-C oracle (oracle/grmonty_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgrmonty_oracle.so")

N_TH, N_E, N_F = 6, 200, 13
SPEC_FIELDS = ["dn_dle", "de_dle", "nph", "nscatt", "x1i_av", "x2i_sq", "x3f_sq", "tau_abs", "tau_scatt",
               "ne_0", "theta_e_0", "b_0", "e_0"]
PHOTON_FLAT = 25
dp = C.POINTER(C.c_double)


class OrcModel(C.Structure):
    _fields_ = [
        ("n0", C.c_int), ("n1", C.c_int),
        ("x_start1", C.c_double), ("x_start2", C.c_double), ("dx1", C.c_double), ("dx2", C.c_double),
        ("dx3", C.c_double), ("x_stop1", C.c_double), ("x_stop2", C.c_double),
        ("a", C.c_double), ("h_slope", C.c_double), ("r_0", C.c_double),
        ("l_unit", C.c_double), ("rho_unit", C.c_double), ("b_unit", C.c_double),
        ("theta_e_unit", C.c_double), ("n_e_unit", C.c_double),
        ("k_rho", dp), ("u", dp), ("u_1", dp), ("u_2", dp), ("u_3", dp), ("b_1", dp), ("b_2", dp), ("b_3", dp),
        ("geom_det", dp),
        ("hotcross", dp), ("f", dp), ("k2", dp), ("weight", dp), ("nint", dp), ("dndlnu_max", dp),
        ("photon_n", C.c_double), ("bias_norm", C.c_double), ("d_tau_k", C.c_double), ("x1_min", C.c_double),
        ("seed", C.c_uint64),
        ("bias_max_tau_scatt", C.c_double), ("bias_n_scatt", C.c_double), ("bias_n_recorded", C.c_double),
        ("acc_max_tau_scatt", C.c_double), ("acc_n_scatt", C.c_uint64), ("acc_n_recorded", C.c_uint64),
        ("stats_mode", C.c_int), ("zone_order", C.c_int),
        ("budget", C.c_int), ("gen_fine_from", C.c_int64), ("gen_fine_div", C.c_int64), ("gen_ramp", C.c_int64), ("gen_budget_spread", C.c_int64), ("carry", C.c_void_p), ("n_carry", C.c_uint64), ("cap_carry", C.c_uint64),
        ("spectrum", C.c_double * (N_TH * N_E * N_F)),
        ("n_created", C.c_uint64),
        ("n_steps", C.c_uint64), ("n_push_attempts", C.c_uint64), ("n_interactions", C.c_uint64),
        ("n_scatter_events", C.c_uint64), ("n_tracked", C.c_uint64),
        ("stats_lag", C.c_int),
    ]


class OrcRng(C.Structure):
    _fields_ = [("id", C.c_uint32 * 3), ("ctr", C.c_uint32)]


class OrcPhoton(C.Structure):
    _fields_ = [("x", C.c_double * 4), ("k", C.c_double * 4), ("dkdlam", C.c_double * 4),
                ("w", C.c_double), ("e", C.c_double), ("l", C.c_double), ("x1i", C.c_double), ("x2i", C.c_double),
                ("tau_abs", C.c_double), ("tau_scatt", C.c_double), ("n_e_0", C.c_double),
                ("theta_e_0", C.c_double), ("b_0", C.c_double), ("e_0", C.c_double), ("e_0_s", C.c_double),
                ("n_scatt", C.c_int), ("rng", OrcRng)]


class OrcFluid(C.Structure):
    _fields_ = [("n_e", C.c_double), ("theta_e", C.c_double), ("b", C.c_double), ("u_con", C.c_double * 4),
                ("u_cov", C.c_double * 4), ("b_con", C.c_double * 4), ("b_cov", C.c_double * 4)]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "grmonty_oracle.c")
    hdr = os.path.join(HERE, "grmonty_oracle.h")
    if (force or not os.path.exists(LIB_PATH)
            or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["gcc", "-std=c99", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-o", LIB_PATH,
                               src, "-lm"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        assert L.orc_sizeof_model() == C.sizeof(OrcModel), "orc_model layout mismatch"
        L.orc_uniform.restype = C.c_double
        L.orc_chi_sq.restype = C.c_double
        for name in ["orc_step_size", "orc_bias_func", "orc_bk_angle", "orc_fluid_nu", "orc_alpha_inv_scatt",
                     "orc_alpha_inv_abs", "orc_synch", "orc_k2_eval", "orc_f_eval", "orc_hotcross_lkup",
                     "orc_hotcross_num", "orc_sample_y", "orc_sample_mu", "orc_sample_klein_nishina",
                     "orc_sample_thomson"]:
            getattr(L, name).restype = C.c_double
        L.orc_zone_counts.restype = C.c_uint64
        L.orc_generation_size.restype = C.c_int64
        L.orc_generation_size.argtypes = [C.c_int64] * 6
        L.orc_perm_multiplier.restype = C.c_int64
        L.orc_perm_multiplier.argtypes = [C.c_int64]
        L.orc_permute.restype = C.c_int64
        L.orc_permute.argtypes = [C.c_int64, C.c_int64, C.c_int64]
        L.orc_run.argtypes = [C.POINTER(OrcModel), C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64]
        L.orc_bias_func.argtypes = [C.POINTER(OrcModel), C.c_double, C.c_double]
        L.orc_alpha_inv_scatt.argtypes = [C.POINTER(OrcModel), C.c_double, C.c_double, C.c_double]
        L.orc_alpha_inv_abs.argtypes = [C.POINTER(OrcModel)] + [C.c_double] * 5
        L.orc_synch.argtypes = [C.POINTER(OrcModel)] + [C.c_double] * 5
        L.orc_k2_eval.argtypes = [C.POINTER(OrcModel), C.c_double]
        L.orc_f_eval.argtypes = [C.POINTER(OrcModel)] + [C.c_double] * 3
        L.orc_hotcross_lkup.argtypes = [C.POINTER(OrcModel), C.c_double, C.c_double]
        L.orc_hotcross_num.argtypes = [C.c_double] * 3
        L.orc_sample_y.argtypes = [C.POINTER(OrcModel), C.POINTER(OrcRng), C.c_double]
        L.orc_sample_mu.argtypes = [C.POINTER(OrcModel), C.POINTER(OrcRng), C.c_double]
        L.orc_sample_klein_nishina.argtypes = [C.POINTER(OrcModel), C.POINTER(OrcRng), C.c_double]
        L.orc_sample_thomson.argtypes = [C.POINTER(OrcModel), C.POINTER(OrcRng)]
        L.orc_chi_sq.argtypes = [C.POINTER(OrcModel), C.POINTER(OrcRng), C.c_int]
        L.orc_push_photon.argtypes = [C.POINTER(OrcModel), C.POINTER(OrcPhoton), C.c_double, C.c_int]
        L.orc_sample_zone_photon.argtypes = [C.POINTER(OrcModel), C.c_int, C.c_int, C.c_double, C.POINTER(OrcRng),
                                             C.POINTER(OrcPhoton)]
        L.orc_make_primary.argtypes = [C.POINTER(OrcModel), C.POINTER(C.c_int64), dp, C.c_int64,
                                       C.POINTER(OrcPhoton)]
        L.orc_run_primary.argtypes = [C.POINTER(OrcModel), C.POINTER(C.c_int64), dp, C.c_int64, C.c_int]
        L.orc_rng_primary.argtypes = [C.POINTER(OrcRng), C.c_uint64]
        L.orc_rng_zone.argtypes = [C.POINTER(OrcRng), C.c_uint64]
        _lib = L
    return _lib


def _vec(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(dp)


class Model:
    """Owns the numpy arrays an orc_model borrows.  `d` is a dict with the keys of tests/golden model files."""

    GRIDS = ["k_rho", "u", "u_1", "u_2", "u_3", "b_1", "b_2", "b_3", "geom_det"]
    TABLES = ["hotcross", "f", "k2", "weight", "nint", "dndlnu_max"]
    SCALARS = ["x_start1", "x_start2", "dx1", "dx2", "dx3", "x_stop1", "x_stop2", "a", "h_slope", "r_0", "l_unit",
               "rho_unit", "b_unit", "theta_e_unit", "n_e_unit", "photon_n", "bias_norm", "d_tau_k", "x1_min"]

    def __init__(self, d, seed=123, stats_mode=0, zone_order=0):
        self.L = lib()
        self.m = OrcModel()
        self._keep = {}
        self.m.n0, self.m.n1 = int(d["n0"]), int(d["n1"])
        for k in self.SCALARS:
            setattr(self.m, k, float(d[k]))
        for k in self.GRIDS + self.TABLES:
            arr, ptr = _vec(d[k])
            self._keep[k] = arr
            setattr(self.m, k, ptr)
        self.m.seed = seed
        self.m.stats_mode = stats_mode
        self.m.zone_order = zone_order
        self.set_bias_stats(float(d["max_tau_scatt0"]), 0, 0)
        self.m.acc_max_tau_scatt = float(d["max_tau_scatt0"])

    @property
    def ptr(self):
        return C.byref(self.m)

    def set_bias_stats(self, max_tau, n_scatt, n_rec):
        self.m.bias_max_tau_scatt = max_tau
        self.m.bias_n_scatt = float(n_scatt)
        self.m.bias_n_recorded = float(n_rec)

    def spectrum(self):
        return np.array(self.m.spectrum, dtype=np.float64).reshape(N_TH, N_E, N_F)

    def clear(self):
        self.L.orc_clear_outputs(self.ptr)

    # ---- function-level wrappers (numpy in / numpy out) ----
    def gcov(self, x):
        x, px = _vec(x)
        out = np.zeros((4, 4))
        self.L.orc_gcov(self.ptr, px, out.ctypes.data_as(dp))
        return out

    def gcon(self, x):
        x, px = _vec(x)
        out = np.zeros((4, 4))
        self.L.orc_gcon(self.ptr, px, out.ctypes.data_as(dp))
        return out

    def connection(self, x):
        x, px = _vec(x)
        out = np.zeros((4, 4, 4))
        self.L.orc_get_connection(self.ptr, px, out.ctypes.data_as(dp))
        return out

    def init_dkdlam(self, x, k):
        x, px = _vec(x)
        k, pk = _vec(k)
        out = np.zeros(4)
        self.L.orc_init_dkdlam(self.ptr, px, pk, out.ctypes.data_as(dp))
        return out

    def step_size(self, x, k):
        x, px = _vec(x)
        k, pk = _vec(k)
        return self.L.orc_step_size(self.ptr, px, pk)

    def photon(self, flat):
        ph = OrcPhoton()
        f, pf = _vec(flat)
        self.L.orc_photon_from_flat(pf, C.byref(ph))
        return ph

    def flat(self, ph):
        out = np.zeros(PHOTON_FLAT)
        self.L.orc_photon_to_flat(C.byref(ph), out.ctypes.data_as(dp))
        return out

    def push_photon(self, flat, dl):
        ph = self.photon(flat)
        self.L.orc_push_photon(self.ptr, C.byref(ph), dl, 0)
        return self.flat(ph)

    def fluid_params(self, x):
        x, px = _vec(x)
        g = self.gcov(x)
        f = OrcFluid()
        self.L.orc_get_fluid_params(self.ptr, px, g.ctypes.data_as(dp), C.byref(f))
        return np.array([f.n_e, f.theta_e, f.b, *f.u_con, *f.u_cov, *f.b_con, *f.b_cov])

    def fluid_zone(self, i, j):
        f = OrcFluid()
        self.L.orc_get_fluid_zone(self.ptr, i, j, C.byref(f))
        return np.array([f.n_e, f.theta_e, f.b, *f.u_con, *f.b_con])

    def init_zone(self, i, j):
        nz, dn = C.c_double(), C.c_double()
        self.L.orc_init_zone(self.ptr, i, j, C.byref(nz), C.byref(dn))
        return nz.value, dn.value

    def zone_counts(self):
        nz = self.m.n0 * self.m.n1
        num = np.zeros(nz, dtype=np.int64)
        dn = np.zeros(nz)
        tot = self.L.orc_zone_counts(self.ptr, num.ctypes.data_as(C.POINTER(C.c_int64)), dn.ctypes.data_as(dp))
        return int(tot), num, dn

    def make_tetrad(self, u_con, trial, gcov):
        u, pu = _vec(u_con)
        t, pt = _vec(np.array(trial, dtype=np.float64).copy())
        g, pg = _vec(gcov)
        ec, ev = np.zeros((4, 4)), np.zeros((4, 4))
        self.L.orc_make_tetrad(pu, pt, pg, ec.ctypes.data_as(dp), ev.ctypes.data_as(dp))
        return ec, ev

    def track(self, flat, rng_id=(0, 0, 0), ctr=0):
        ph = self.photon(flat)
        ph.rng.id[0], ph.rng.id[1], ph.rng.id[2] = rng_id
        ph.rng.ctr = ctr
        self.L.orc_track_super_photon(self.ptr, C.byref(ph))
        return self.flat(ph)

    def run(self, first=0, last=-1, rank=0, world=1, gen0=32, gen_cap=1 << 22, budget=384, fine_from=16384,
            fine_div=6, ramp=8, spread=0, stats_lag=0):
        self.m.stats_lag = stats_lag
        self.m.budget = budget if self.m.stats_mode == 0 else 0
        self.m.gen_fine_from, self.m.gen_fine_div, self.m.gen_ramp = fine_from, fine_div, ramp
        self.m.gen_budget_spread = spread
        self.L.orc_run(self.ptr, first, last, rank, world, gen0, gen_cap)
