"""ctypes binding of oracle/_ref/libgrmonty_ref.so = the UNMODIFIED reference CPU sources + oracle/ref_harness.cpp.

TEST INFRASTRUCTURE ONLY.  Used (in the build container, where /root/reference exists) to generate golden
vectors and to validate the C restatement; on the GPU box only the prebuilt files under oracle/_ref/ are used.
One model per process (the reference keeps function-static state).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
LIB_PATH = os.path.join(REF_DIR, "libgrmonty_ref.so")
CLI_PATH = os.path.join(REF_DIR, "grmonty_ref")
HOTCROSS_CACHE = os.path.join(REF_DIR, "hotcross_table.bin")
dp = C.POINTER(C.c_double)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _p(a):
    return a.ctypes.data_as(dp)


class Ref:
    def __init__(self, dump_path: str, photon_n: int, mass_unit: float, seed: int = 123, verbose: int = 0):
        self.L = L = C.CDLL(LIB_PATH)
        for n in ["ref_uniform", "ref_chi_sq", "ref_step_size", "ref_d_omega", "ref_bias_func", "ref_bk_angle",
                  "ref_fluid_nu", "ref_alpha_inv_scatt", "ref_alpha_inv_abs", "ref_synch", "ref_k2_eval",
                  "ref_f_eval", "ref_hotcross_lkup", "ref_sample_y", "ref_sample_mu", "ref_sample_kn",
                  "ref_sample_thomson", "ref_run_simulation"]:
            getattr(L, n).restype = C.c_double
        L.ref_create.argtypes = [C.c_int, C.c_double, C.c_int]
        L.ref_set_bias_stats.argtypes = [C.c_double, C.c_uint64, C.c_uint64]
        L.ref_bias_func.argtypes = [C.c_double, C.c_double]
        L.ref_alpha_inv_scatt.argtypes = [C.c_double] * 3
        L.ref_alpha_inv_abs.argtypes = [C.c_double] * 5
        L.ref_synch.argtypes = [C.c_double] * 5
        L.ref_k2_eval.argtypes = [C.c_double]
        L.ref_f_eval.argtypes = [C.c_double] * 3
        L.ref_hotcross_lkup.argtypes = [C.c_double] * 2
        L.ref_sample_y.argtypes = [C.c_double]
        L.ref_sample_mu.argtypes = [C.c_double]
        L.ref_sample_kn.argtypes = [C.c_double]
        L.ref_d_omega.argtypes = [C.c_double] * 2
        L.ref_push_photon.argtypes = [dp, C.c_double]
        L.ref_bk_angle.argtypes = [dp, dp, dp, dp, C.c_double]
        L.ref_sample_electron.argtypes = [dp, C.c_double, dp]
        L.ref_sample_zone_photons.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, dp]
        L.ref_create(photon_n, mass_unit, verbose)
        if L.ref_read_file(dump_path.encode()) != 0:
            raise RuntimeError("ref_read_file failed")
        L.ref_init(HOTCROSS_CACHE.encode())
        L.ref_rng_init(seed)
        self.photon_n = photon_n

    # ---- model data ----
    def model_dict(self) -> dict:
        L = self.L
        h = np.zeros(13)
        L.ref_get_header(_p(h))
        u = np.zeros(8)
        L.ref_get_units(_p(u))
        s = np.zeros(5)
        L.ref_get_scalars(_p(s))
        n0, n1 = int(h[0]), int(h[1])
        d = dict(n0=n0, n1=n1, x_start1=h[2], x_start2=h[3], dx1=h[4], dx2=h[5], dx3=h[6], x_stop1=h[7],
                 x_stop2=h[8], a=h[9], h_slope=h[10], r_0=h[11], gamma=h[12], mass_unit=u[0], l_unit=u[1],
                 t_unit=u[2], rho_unit=u[3], u_unit=u[4], b_unit=u[5], theta_e_unit=u[6], n_e_unit=u[7],
                 bias_norm=s[0], rh=s[1], max_tau_scatt0=s[2], d_tau_k=s[3], x1_min=s[4],
                 photon_n=float(self.photon_n))
        names = ["k_rho", "u", "u_1", "u_2", "u_3", "b_1", "b_2", "b_3", "geom_det"]
        for i, nm in enumerate(names):
            a = np.zeros((n0, n1))
            L.ref_get_grid(i, _p(a))
            d[nm] = a
        for i, (nm, n) in enumerate([("hotcross", 221 * 81), ("f", 201), ("k2", 201), ("weight", 201),
                                     ("nint", 20001), ("dndlnu_max", 20001)]):
            a = np.zeros(n)
            L.ref_get_table(i, _p(a))
            d[nm] = a
        return d

    def spectrum(self):
        a = np.zeros((6, 200, 13))
        self.L.ref_get_spectrum(_p(a))
        return a

    def counters(self):
        c = np.zeros(3, dtype=np.uint64)
        self.L.ref_get_counters(c.ctypes.data_as(C.POINTER(C.c_uint64)))
        return c

    # ---- functions ----
    def gcov(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        o = np.zeros((4, 4))
        self.L.ref_gcov(_p(x), _p(o))
        return o

    def gcon(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        o = np.zeros((4, 4))
        self.L.ref_gcon(_p(x), _p(o))
        return o

    def connection(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        o = np.zeros((4, 4, 4))
        self.L.ref_get_connection(_p(x), _p(o))
        return o

    def init_dkdlam(self, x, k):
        x = np.ascontiguousarray(x, dtype=np.float64)
        k = np.ascontiguousarray(k, dtype=np.float64)
        o = np.zeros(4)
        self.L.ref_init_dkdlam(_p(x), _p(k), _p(o))
        return o

    def step_size(self, x, k):
        x = np.ascontiguousarray(x, dtype=np.float64)
        k = np.ascontiguousarray(k, dtype=np.float64)
        return self.L.ref_step_size(_p(x), _p(k))

    def push_photon(self, flat, dl):
        f = np.array(flat, dtype=np.float64).copy()
        self.L.ref_push_photon(_p(f), dl)
        return f

    def fluid_params(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        o = np.zeros(19)
        self.L.ref_get_fluid_params(_p(x), _p(o))
        return o

    def fluid_zone(self, i, j):
        o = np.zeros(11)
        self.L.ref_get_fluid_zone(i, j, _p(o))
        return o

    def init_zone(self, i, j):
        o = np.zeros(2)
        self.L.ref_init_zone(i, j, _p(o))
        return o

    def make_tetrad(self, u_con, trial, gcov):
        u = np.ascontiguousarray(u_con, dtype=np.float64)
        t = np.ascontiguousarray(trial, dtype=np.float64)
        g = np.ascontiguousarray(gcov, dtype=np.float64)
        ec, ev = np.zeros((4, 4)), np.zeros((4, 4))
        self.L.ref_make_tetrad(_p(u), _p(t), _p(g), _p(ec), _p(ev))
        return ec, ev

    def bk_angle(self, x, k, ucov, bcov, b):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, k, ucov, bcov)]
        return self.L.ref_bk_angle(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), b)

    def fluid_nu(self, x, k, ucov):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, k, ucov)]
        return self.L.ref_fluid_nu(_p(a[0]), _p(a[1]), _p(a[2]))

    def track(self, flat):
        f = np.array(flat, dtype=np.float64).copy()
        self.L.ref_track_super_photon(_p(f))
        return f

    def make_super_photon(self):
        o = np.zeros(16)
        q = self.L.ref_make_super_photon(_p(o))
        return None if q else o

    def sample_zone_photons(self, i, j, dn_max, n):
        o = np.zeros((n, 16))
        self.L.ref_sample_zone_photons(i, j, dn_max, n, _p(o))
        return o

    def sample_electron(self, k, te):
        k = np.ascontiguousarray(k, dtype=np.float64)
        p = np.zeros(4)
        self.L.ref_sample_electron(_p(k), te, _p(p))
        return p

    def sample_scattered_photon(self, k, p):
        k = np.ascontiguousarray(k, dtype=np.float64)
        p = np.array(p, dtype=np.float64).copy()
        kp = np.zeros(4)
        self.L.ref_sample_scattered_photon(_p(k), _p(p), _p(kp))
        return kp


def init_to_flat(ip: np.ndarray) -> np.ndarray:
    """InitPhoton (16) -> Photon flat (25), the copy the reference does at harm_model.cpp:373-391."""
    f = np.zeros(25)
    f[0:4] = ip[0:4]
    f[4:8] = ip[4:8]
    f[12] = ip[8]   # w
    f[13] = ip[9]   # e
    f[14] = ip[10]  # l
    f[15] = ip[1]   # x1i
    f[16] = ip[2]   # x2i
    f[19] = ip[11]  # n_e_0
    f[20] = ip[12]  # theta_e_0
    f[21] = ip[13]  # b_0
    f[22] = ip[14]  # e_0
    f[23] = ip[9]   # e_0_s = e
    f[24] = 0
    return f
