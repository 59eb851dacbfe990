"""cuda-grmonty_b200 -- B200-native superphoton transport (grmonty) behind the reference's host surface.

This Python module is plumbing only: it builds the native artefacts in-tree and binds the C ABI
(include/grmonty_b200.h) with ctypes so that tests/, bench.py and __graft_entry__.py can drive the CUDA
path.  There is no Python or CPU implementation of the transport path here: if the CUDA library cannot be
loaded or no CUDA device is present, every call fails loudly.

Artefacts (all git-ignored, built by `build()`):
    cuda-grmonty_b200/libgrmonty_b200.so        CUDA kernels + C ABI (csrc/)
    cuda-grmonty_b200/libgrmonty_b200_host.so   C++20 host: HARM dump loader, table builders, spectrum writer
    cuda-grmonty_b200/grmonty_b200              CLI with the reference's flags (host/main.cpp)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
HOST = os.path.join(PKG_DIR, "host")
INCLUDE = os.path.join(ROOT, "include")
LIB_CUDA = os.environ.get("GRMONTY_B200_LIB", os.path.join(PKG_DIR, "libgrmonty_b200.so"))
# the same sources + the test-only batch exports of include/grmonty_b200_test.h (never loaded by the product path)
LIB_CUDA_TEST = os.environ.get("GRMONTY_B200_TEST_LIB", os.path.join(PKG_DIR, "libgrmonty_b200_test.so"))
LIB_HOST = os.path.join(PKG_DIR, "libgrmonty_b200_host.so")
CLI = os.path.join(PKG_DIR, "grmonty_b200")

# --cudart=shared: the CUDA runtime is taken from the process (torch's copy, or /usr/local/cuda/lib64 for the CLI)
# instead of being linked statically into every artefact
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--cudart=shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64"]

ABI_VERSION = 4
N_TH, N_E, N_F = 6, 200, 13
SPEC_FIELDS = ["dn_dle", "de_dle", "nph", "nscatt", "x1i_av", "x2i_sq", "x3f_sq", "tau_abs", "tau_scatt",
               "ne_0", "theta_e_0", "b_0", "e_0"]
dp = C.POINTER(C.c_double)


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _sources(d: str, exts) -> list[str]:
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(tuple(exts)))


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a: cross-compiles without a GPU.  Two libraries from the same
    sources: the product library and the test superset (-DGRMONTY_B200_TEST_EXPORTS), compiled side by side."""
    srcs = _sources(CSRC, (".cu", ".cuh", ".h", ".inc")) + [os.path.join(INCLUDE, "grmonty_b200.h"),
                                                             os.path.join(INCLUDE, "grmonty_b200_test.h")]
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    procs = []
    for target, extra in ((LIB_CUDA, []), (LIB_CUDA_TEST, ["-DGRMONTY_B200_TEST_EXPORTS"])):
        if force or _stale(target, srcs):
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-shared", "-o", target, os.path.join(CSRC, "gm_api.cu"), "-ldl"]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    return LIB_CUDA


def build_host(force: bool = False) -> str:
    srcs = _sources(HOST, (".cpp", ".hpp", ".h")) + [os.path.join(INCLUDE, "grmonty_b200.h")]
    cpps = [s for s in srcs if s.endswith(".cpp") and not s.endswith("main.cpp")]
    if not cpps:
        return ""
    if force or _stale(LIB_HOST, srcs):
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-I" + INCLUDE, "-o", LIB_HOST, *cpps,
                               "-lpthread", "-ldl"])
    main = os.path.join(HOST, "main.cpp")
    if os.path.exists(main) and (force or _stale(CLI, srcs)):
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-I" + INCLUDE, "-o", CLI, main, *cpps, "-lpthread",
                               "-ldl"])
    return LIB_HOST


def build(force: bool = False) -> None:
    build_cuda(force)
    build_host(force)


# ---- C ABI structs (must mirror include/grmonty_b200.h) ------------------------------------------------------
class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("struct_size", C.c_uint32),
        ("n0", C.c_int32), ("n1", C.c_int32),
        ("x_start1", C.c_double), ("x_start2", C.c_double),
        ("dx1", C.c_double), ("dx2", C.c_double), ("dx3", C.c_double),
        ("x_stop1", C.c_double), ("x_stop2", C.c_double),
        ("a", C.c_double), ("h_slope", C.c_double), ("r_0", C.c_double),
        ("mass_unit", C.c_double), ("l_unit", C.c_double), ("t_unit", C.c_double), ("rho_unit", C.c_double),
        ("u_unit", C.c_double), ("b_unit", C.c_double), ("theta_e_unit", C.c_double), ("n_e_unit", C.c_double),
        ("k_rho", dp), ("u", dp), ("u_1", dp), ("u_2", dp), ("u_3", dp), ("b_1", dp), ("b_2", dp), ("b_3", dp),
        ("geom_det", dp),
        ("hotcross", dp), ("f", dp), ("k2", dp), ("weight", dp), ("nint", dp), ("dndlnu_max", dp),
        ("photon_n", C.c_double), ("bias_norm", C.c_double), ("max_tau_scatt0", C.c_double),
        ("seed", C.c_uint64),
        ("rank", C.c_int32), ("world", C.c_int32), ("device", C.c_int32),
        ("threads_per_block", C.c_int32), ("blocks_per_sm", C.c_int32),
        ("queue_capacity", C.c_int64), ("gen0", C.c_int64), ("gen_cap", C.c_int64), ("gen_budget", C.c_int64),
        ("gen_budget_spread", C.c_int64), ("gen_fine_from", C.c_int64), ("gen_ramp", C.c_int64), ("gen_fine_div", C.c_int64),
        ("kernel", C.c_int32), ("slots_per_thread", C.c_int32), ("wf_thr_interact", C.c_int32), ("wf_thr_service", C.c_int32),
        ("gen_overlap", C.c_int32), ("reserved0", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [("n_tracked", C.c_uint64), ("n_steps", C.c_uint64), ("n_push_attempts", C.c_uint64),
                ("n_interactions", C.c_uint64), ("n_scatter_events", C.c_uint64), ("n_generations", C.c_uint64),
                ("n_kernel_launches", C.c_uint64), ("queue_high_water", C.c_uint64),
                ("n_live_iterations", C.c_uint64), ("n_slot_iterations", C.c_uint64),
                ("kernel_ms", C.c_double), ("transport_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


ABI_SYMBOLS = [
    "grmonty_b200_create", "grmonty_b200_total_primaries", "grmonty_b200_run_range", "grmonty_b200_run",
    "grmonty_b200_set_progress",
    "grmonty_b200_allreduce", "grmonty_b200_device_accumulators", "grmonty_b200_result", "grmonty_b200_reset",
    "grmonty_b200_destroy", "grmonty_b200_trim_cache", "grmonty_b200_last_error", "grmonty_b200_fp64_peak", "grmonty_b200_hotcross_table",
    "grmonty_b200_init_tables", "grmonty_b200_nccl_unique_id", "grmonty_b200_nccl_comm_init_rank",
    "grmonty_b200_nccl_comm_init_all", "grmonty_b200_nccl_comm_destroy",
]
# include/grmonty_b200_test.h: only in libgrmonty_b200_test.so
TEST_ABI_SYMBOLS = [
    "grmonty_b200_test_geometry", "grmonty_b200_test_dkdlam_step", "grmonty_b200_test_push_photon",
    "grmonty_b200_test_trajectory", "grmonty_b200_test_fluid_params", "grmonty_b200_test_radiation",
    "grmonty_b200_test_hotcross", "grmonty_b200_test_angles", "grmonty_b200_test_tetrad",
    "grmonty_b200_test_zones", "grmonty_b200_test_bias", "grmonty_b200_test_make_primaries",
    "grmonty_b200_test_track", "grmonty_b200_test_samplers", "grmonty_b200_test_philox",
    "grmonty_b200_test_bounds_violations",
]

# launch geometries compiled into the library (csrc/gm_api.cu): (kernel, threads per block, blocks per SM or slots per
# thread).  kernel 1 = the fused loop (the default, also selected by 0), 2 = wavefront (third number: photon slots
# per thread)
KERNEL_FUSED, KERNEL_WAVEFRONT = 1, 2
OVERLAP_ON, OVERLAP_OFF = 1, 2     # grmonty_b200_config.gen_overlap
KERNEL_VARIANTS = [(1, 32, 8), (1, 256, 1), (1, 64, 4), (1, 128, 2), (1, 384, 1),
                   (2, 384, 2), (2, 256, 2), (2, 256, 3), (2, 512, 1), (2, 384, 1), (2, 128, 2)]

_libs = {}


class GrmontyError(RuntimeError):
    pass


def lib(test_exports: bool = False):
    """Load the CUDA library (test_exports: the superset library with the grmonty_b200_test_* entry points).
    Raises if it is missing: there is no fallback implementation."""
    if test_exports not in _libs:
        path = LIB_CUDA_TEST if test_exports else LIB_CUDA
        if not os.path.exists(path):
            raise GrmontyError(f"{path} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        L.grmonty_b200_last_error.restype = C.c_char_p
        L.grmonty_b200_last_error.argtypes = [C.c_void_p]
        L.grmonty_b200_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config)]
        L.grmonty_b200_run_range.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.grmonty_b200_run.argtypes = [C.c_void_p]
        L.grmonty_b200_reset.argtypes = [C.c_void_p]
        L.grmonty_b200_destroy.argtypes = [C.c_void_p]
        L.grmonty_b200_destroy.restype = None
        L.grmonty_b200_total_primaries.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.grmonty_b200_result.argtypes = [C.c_void_p, dp, C.POINTER(C.c_uint64), dp, C.POINTER(Stats)]
        L.grmonty_b200_device_accumulators.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 3
        L.grmonty_b200_allreduce.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.grmonty_b200_fp64_peak.argtypes = [C.c_void_p, dp]
        L.grmonty_b200_hotcross_table.argtypes = [C.c_int, dp]
        L.grmonty_b200_init_tables.argtypes = [C.POINTER(Config), dp, dp, dp, dp, C.POINTER(C.c_double)]
        L.grmonty_b200_nccl_unique_id.argtypes = [C.c_void_p]
        L.grmonty_b200_nccl_comm_init_rank.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.grmonty_b200_nccl_comm_init_all.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int)]
        L.grmonty_b200_nccl_comm_destroy.argtypes = [C.c_void_p]
        _libs[test_exports] = L
    return _libs[test_exports]


NCCL_ID_BYTES = 128


def nccl_unique_id() -> bytes:
    """ncclGetUniqueId through the product library (rank 0; ship the bytes to the other ranks)"""
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    rc = lib().grmonty_b200_nccl_unique_id(buf)
    if rc != 0:
        raise GrmontyError(f"grmonty_b200_nccl_unique_id failed ({rc}): {lib().grmonty_b200_last_error(None).decode()}")
    return buf.raw


def nccl_comm_init_rank(uid: bytes, rank: int, world: int, device: int) -> int:
    """ncclCommInitRank through the product library; returns the ncclComm_t as an integer handle"""
    comm = C.c_void_p()
    buf = C.create_string_buffer(uid, NCCL_ID_BYTES)
    rc = lib().grmonty_b200_nccl_comm_init_rank(C.byref(comm), buf, rank, world, device)
    if rc != 0:
        raise GrmontyError(f"grmonty_b200_nccl_comm_init_rank failed ({rc}): "
                           f"{lib().grmonty_b200_last_error(None).decode()}")
    return comm.value


def nccl_comm_destroy(comm: int):
    lib().grmonty_b200_nccl_comm_destroy(C.c_void_p(comm))


def _arr(a, dtype=np.float64):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


class Context:
    """One grmonty_b200 context = one GPU's share of a run.  `model` is a dict with the grids, tables and scalars
    HARMModel holds after read_file() + init() (see host.HarmModel.model_dict / tests/golden)."""

    GRIDS = ["k_rho", "u", "u_1", "u_2", "u_3", "b_1", "b_2", "b_3", "geom_det"]
    TABLES = {"hotcross": 221 * 81, "f": 201, "k2": 201, "weight": 201, "nint": 20001, "dndlnu_max": 20001}
    SCALARS = ["x_start1", "x_start2", "dx1", "dx2", "dx3", "x_stop1", "x_stop2", "a", "h_slope", "r_0",
               "mass_unit", "l_unit", "t_unit", "rho_unit", "u_unit", "b_unit", "theta_e_unit", "n_e_unit",
               "photon_n", "bias_norm", "max_tau_scatt0"]

    def __init__(self, model: dict, seed: int = 123, rank: int = 0, world: int = 1, device: int = 0,
                 threads_per_block: int = 0, blocks_per_sm: int = 0, queue_capacity: int = 0, gen0: int = 0,
                 gen_cap: int = 0, gen_budget: int = 0, gen_fine_from: int = 0, gen_fine_div: int = 0,
                 gen_ramp: int = 0, gen_budget_spread: int = 0, kernel: int = 0, slots_per_thread: int = 0,
                 wf_thr_interact: int = 0, wf_thr_service: int = 0, gen_overlap: int = 0, test_exports: bool = False):
        # test_exports: create the context in libgrmonty_b200_test.so so that the t_* batch exports can be called on it
        self.L = lib(test_exports)
        cfg, self._keep = make_config(model, seed=seed, rank=rank, world=world, device=device,
                                      threads_per_block=threads_per_block, blocks_per_sm=blocks_per_sm,
                                      queue_capacity=queue_capacity, gen0=gen0, gen_cap=gen_cap, gen_budget=gen_budget,
                                      gen_fine_from=gen_fine_from, gen_fine_div=gen_fine_div, gen_ramp=gen_ramp,
                                      gen_budget_spread=gen_budget_spread, kernel=kernel,
                                      slots_per_thread=slots_per_thread, wf_thr_interact=wf_thr_interact,
                                      wf_thr_service=wf_thr_service, gen_overlap=gen_overlap)
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.L.grmonty_b200_create(C.byref(self.h), C.byref(cfg))
        if rc != 0:
            raise GrmontyError(f"grmonty_b200_create failed ({rc}): "
                               f"{self.L.grmonty_b200_last_error(None).decode()}")
        self.n0, self.n1 = cfg.n0, cfg.n1

    def _ck(self, rc):
        if rc != 0:
            raise GrmontyError(f"grmonty_b200 error {rc}: {self.L.grmonty_b200_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.L.grmonty_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- product API ----
    def total_primaries(self) -> int:
        t = C.c_int64()
        self._ck(self.L.grmonty_b200_total_primaries(self.h, C.byref(t)))
        return t.value

    def run(self, first: int = 0, last: int = -1):
        self._ck(self.L.grmonty_b200_run_range(self.h, first, last))

    def reset(self):
        self._ck(self.L.grmonty_b200_reset(self.h))

    def result(self):
        spec = np.zeros((N_TH, N_E, N_F))
        counts = (C.c_uint64 * 3)()
        mt = C.c_double()
        st = Stats()
        self._ck(self.L.grmonty_b200_result(self.h, _ptr(spec), counts, C.byref(mt), C.byref(st)))
        return dict(spectrum=spec, created=counts[0], scattered=counts[1], recorded=counts[2],
                    max_tau_scatt=mt.value, stats=st.as_dict())

    def allreduce(self, nccl_comm: int, stream: int = 0):
        """the path's only collective: sum / max of the device accumulators over the ranks of `nccl_comm`"""
        self._ck(self.L.grmonty_b200_allreduce(self.h, C.c_void_p(nccl_comm), C.c_void_p(stream)))

    def device_accumulators(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(self.L.grmonty_b200_device_accumulators(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def fp64_peak(self) -> float:
        t = C.c_double()
        self._ck(self.L.grmonty_b200_fp64_peak(self.h, C.byref(t)))
        return t.value

    # ---- test exports ----
    def t_geometry(self, x):
        x = _arr(x).reshape(-1, 4)
        n = len(x)
        gcov, gcon, conn = np.zeros((n, 4, 4)), np.zeros((n, 4, 4)), np.zeros((n, 4, 4, 4))
        self._ck(self.L.grmonty_b200_test_geometry(self.h, C.c_int64(n), _ptr(x), _ptr(gcov), _ptr(gcon),
                                                   _ptr(conn)))
        return gcov, gcon, conn

    def t_dkdlam_step(self, x, k):
        x, k = _arr(x).reshape(-1, 4), _arr(k).reshape(-1, 4)
        n = len(x)
        dk, st = np.zeros((n, 4)), np.zeros(n)
        self._ck(self.L.grmonty_b200_test_dkdlam_step(self.h, C.c_int64(n), _ptr(x), _ptr(k), _ptr(dk), _ptr(st)))
        return dk, st

    def t_push_photon(self, photons, dl):
        p = _arr(photons).reshape(-1, 25).copy()
        dl = _arr(dl).reshape(-1)
        att = np.zeros(len(p), dtype=np.int32)
        self._ck(self.L.grmonty_b200_test_push_photon(self.h, C.c_int64(len(p)), _ptr(p), _ptr(dl),
                                                      _ptr(att, C.c_int32)))
        return p, att

    def t_trajectory(self, photons, nsteps, stride):
        p = _arr(photons).reshape(-1, 25).copy()
        tr = np.zeros((len(p), nsteps // stride, 9))
        self._ck(self.L.grmonty_b200_test_trajectory(self.h, C.c_int64(len(p)), _ptr(p), nsteps, stride, _ptr(tr)))
        return p, tr

    def t_fluid_params(self, x):
        x = _arr(x).reshape(-1, 4)
        out = np.zeros((len(x), 19))
        self._ck(self.L.grmonty_b200_test_fluid_params(self.h, C.c_int64(len(x)), _ptr(x), _ptr(out)))
        return out

    def t_radiation(self, args):
        a = _arr(args).reshape(-1, 5)
        out = np.zeros((len(a), 5))
        self._ck(self.L.grmonty_b200_test_radiation(self.h, C.c_int64(len(a)), _ptr(a), _ptr(out)))
        return out

    def t_hotcross(self, args):
        a = _arr(args).reshape(-1, 2)
        out = np.zeros(len(a))
        self._ck(self.L.grmonty_b200_test_hotcross(self.h, C.c_int64(len(a)), _ptr(a), _ptr(out)))
        return out

    def t_angles(self, k, fluid):
        k, f = _arr(k).reshape(-1, 4), _arr(fluid).reshape(-1, 19)
        th, nu = np.zeros(len(k)), np.zeros(len(k))
        self._ck(self.L.grmonty_b200_test_angles(self.h, C.c_int64(len(k)), _ptr(k), _ptr(f), _ptr(th), _ptr(nu)))
        return th, nu

    def t_tetrad(self, rows24):
        a = _arr(rows24).reshape(-1, 24)
        ec, ev = np.zeros((len(a), 4, 4)), np.zeros((len(a), 4, 4))
        self._ck(self.L.grmonty_b200_test_tetrad(self.h, C.c_int64(len(a)), _ptr(a), _ptr(ec), _ptr(ev)))
        return ec, ev

    def t_zones(self):
        nz = self.n0 * self.n1
        a, b, c = np.zeros(nz), np.zeros(nz), np.zeros(nz, dtype=np.int64)
        self._ck(self.L.grmonty_b200_test_zones(self.h, _ptr(a), _ptr(b), _ptr(c, C.c_int64)))
        return a.reshape(self.n0, self.n1), b.reshape(self.n0, self.n1), c.reshape(self.n0, self.n1)

    def t_bias(self, args, max_tau, n_scatt, n_rec):
        a = _arr(args).reshape(-1, 2)
        out = np.zeros(len(a))
        self._ck(self.L.grmonty_b200_test_bias(self.h, C.c_int64(len(a)), _ptr(a), C.c_double(max_tau),
                                               C.c_double(n_scatt), C.c_double(n_rec), _ptr(out)))
        return out

    def t_make_primaries(self, idx):
        idx = _arr(idx, np.int64).reshape(-1)
        p = np.zeros((len(idx), 25))
        r = np.zeros((len(idx), 4), dtype=np.uint32)
        self._ck(self.L.grmonty_b200_test_make_primaries(self.h, C.c_int64(len(idx)), _ptr(idx, C.c_int64), _ptr(p),
                                                         _ptr(r, C.c_uint32)))
        return p, r

    def t_track(self, photons, rng, max_tau, n_scatt, n_rec):
        p = _arr(photons).reshape(-1, 25).copy()
        r = _arr(rng, np.uint32).reshape(-1, 4).copy()
        st = np.zeros(len(p), dtype=np.int32)
        self._ck(self.L.grmonty_b200_test_track(self.h, C.c_int64(len(p)), _ptr(p), _ptr(r, C.c_uint32),
                                                C.c_double(max_tau), C.c_double(n_scatt), C.c_double(n_rec),
                                                _ptr(st, C.c_int32)))
        return p, r, st

    def t_samplers(self, which, p0, p1, first_stream, n):
        out = np.zeros(n)
        self._ck(self.L.grmonty_b200_test_samplers(self.h, which, C.c_double(p0), C.c_double(p1),
                                                   C.c_int64(first_stream), C.c_int64(n), _ptr(out)))
        return out

    def t_bounds_violations(self) -> int:
        """pool accesses outside the pool counted by the checked build since the last call (test library only)"""
        v = C.c_uint32(0)
        self._ck(self.L.grmonty_b200_test_bounds_violations(self.h, C.byref(v)))
        return int(v.value)

    def t_philox(self, ctr, key):
        c, k = _arr(ctr, np.uint32).reshape(-1, 4), _arr(key, np.uint32).reshape(-1, 2)
        out = np.zeros((len(c), 4), dtype=np.uint32)
        self._ck(self.L.grmonty_b200_test_philox(self.h, C.c_int64(len(c)), _ptr(c, C.c_uint32),
                                                 _ptr(k, C.c_uint32), _ptr(out, C.c_uint32)))
        return out


def make_config(model: dict, **options):
    """grmonty_b200_config for `model` (see Context) + the arrays it points into (keep them alive as long as the
    config is in use).  `options`: the integer fields of the config tail (seed, rank, world, device, gen0, ...)."""
    cfg = Config()
    cfg.abi_version = ABI_VERSION
    cfg.struct_size = C.sizeof(Config)
    cfg.n0, cfg.n1 = int(model["n0"]), int(model["n1"])
    for k in Context.SCALARS:
        setattr(cfg, k, float(model[k]))
    keep = {}
    nz = cfg.n0 * cfg.n1
    for k in Context.GRIDS:
        a = _arr(model[k]).reshape(-1)
        assert a.size == nz, k
        keep[k] = a
        setattr(cfg, k, _ptr(a))
    for k, n in Context.TABLES.items():
        a = _arr(model[k]).reshape(-1)
        assert a.size == n, k
        keep[k] = a
        setattr(cfg, k, _ptr(a))
    cfg.world = 1
    for k, v in options.items():
        setattr(cfg, k, v)
    return cfg, keep


def _check_test_exports(L):
    if not hasattr(L, "grmonty_b200_test_geometry"):
        raise GrmontyError("this context lives in the product library; create it with test_exports=True")


def hotcross_table(device: int = 0) -> np.ndarray:
    """[221][81] log10 hot cross-section table built on the GPU (grmonty_b200_hotcross_table)"""
    t = np.zeros((221, 81))
    rc = lib().grmonty_b200_hotcross_table(device, _ptr(t))
    if rc != 0:
        raise GrmontyError(f"grmonty_b200_hotcross_table failed ({rc}): {lib().grmonty_b200_last_error(None).decode()}")
    return t


def init_tables(model: dict, device: int = 0) -> dict:
    """geom_det [n0][n1], weight [201], nint [20001], dndlnu_max [20001] built on the GPU (grmonty_b200_init_tables)
    from the grids, units, photon_n and the F(K) / K2 tables of `model`; also `device_ms` of the three kernels."""
    cfg = Config()
    cfg.abi_version = ABI_VERSION
    cfg.struct_size = C.sizeof(Config)
    cfg.n0, cfg.n1 = int(model["n0"]), int(model["n1"])
    for k in Context.SCALARS:
        if k in model:
            setattr(cfg, k, float(model[k]))
    keep = []
    for k in Context.GRIDS[:8] + ["f", "k2"]:
        a = _arr(model[k]).reshape(-1)
        keep.append(a)
        setattr(cfg, k, _ptr(a))
    cfg.device = device
    out = dict(geom_det=np.zeros((cfg.n0, cfg.n1)), weight=np.zeros(201), nint=np.zeros(20001),
               dndlnu_max=np.zeros(20001))
    ms = C.c_double()
    L = lib()
    rc = L.grmonty_b200_init_tables(C.byref(cfg), _ptr(out["geom_det"]), _ptr(out["weight"]), _ptr(out["nint"]),
                                    _ptr(out["dndlnu_max"]), C.byref(ms))
    if rc != 0:
        raise GrmontyError(f"grmonty_b200_init_tables failed ({rc}): {L.grmonty_b200_last_error(None).decode()}")
    out["device_ms"] = ms.value
    return out


# ---- host library binding (libgrmonty_b200_host.so): the reference's HARMModel surface -------------------------
_hlib = None


def host_lib():
    global _hlib
    if _hlib is None:
        if not os.path.exists(LIB_HOST):
            raise GrmontyError(f"{LIB_HOST} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        H = C.CDLL(LIB_HOST)
        H.gmh_last_error.restype = C.c_char_p
        H.gmh_create.restype = C.c_void_p
        H.gmh_create.argtypes = [C.c_int, C.c_double, C.c_int]
        H.gmh_destroy.argtypes = [C.c_void_p]
        H.gmh_read_file.argtypes = [C.c_void_p, C.c_char_p]
        H.gmh_init.argtypes = [C.c_void_p, C.c_int]
        H.gmh_init_stage.argtypes = [C.c_void_p, C.c_int, C.c_int]
        H.gmh_set_options.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                      C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_char_p]
        H.gmh_run_simulation.argtypes = [C.c_void_p]
        H.gmh_set_gpus.argtypes = [C.c_void_p, C.c_int]
        H.gmh_set_external_reduce.argtypes = [C.c_void_p, C.c_int]
        H.gmh_report_spectrum.argtypes = [C.c_void_p, C.c_char_p]
        H.gmh_report_spectrum_binary.argtypes = [C.c_void_p, C.c_char_p]
        H.gmh_set_dump_cache.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        H.gmh_read_from_cache.argtypes = [C.c_void_p]
        H.gmh_set_device_tables.argtypes = [C.c_void_p, C.c_int]
        H.gmh_set_hotcross_cache.argtypes = [C.c_void_p, C.c_char_p]
        H.gmh_hotcross_from_cache.argtypes = [C.c_void_p]
        for n in ("gmh_get_header", "gmh_get_header_raw", "gmh_get_units", "gmh_get_scalars", "gmh_get_spectrum",
                  "gmh_set_spectrum", "gmh_get_stats"):
            getattr(H, n).argtypes = [C.c_void_p, dp]
        H.gmh_get_grid.argtypes = [C.c_void_p, C.c_int, dp]
        H.gmh_get_table.argtypes = [C.c_void_p, C.c_int, dp]
        _hlib = H
    return _hlib


SPECTRUM_BIN_DTYPE = np.dtype([("magic", "S8"), ("version", "<u4"), ("n_th", "<u4"), ("n_e", "<u4"),
                               ("n_fields", "<u4"), ("created", "<u8"), ("scattered", "<u8"), ("recorded", "<u8"),
                               ("max_tau_scatt", "<f8"), ("mass_unit", "<f8"), ("photon_n", "<f8"), ("a", "<f8"),
                               ("h_slope", "<f8"), ("x_start2", "<f8"), ("x_stop2", "<f8")])


def read_spectrum_binary(path: str) -> dict:
    """Reader of HarmModel.report_spectrum_binary: header fields + `spectrum` [6][200][13] (harm::Spectrum order)."""
    with open(path, "rb") as f:
        raw = f.read()
    if len(raw) < SPECTRUM_BIN_DTYPE.itemsize:
        raise GrmontyError(f"{path}: not a grmonty_b200 binary spectrum")
    h = np.frombuffer(raw[:SPECTRUM_BIN_DTYPE.itemsize], dtype=SPECTRUM_BIN_DTYPE)[0]
    if h["magic"] != b"GMB2SPEC" or h["version"] != 1:
        raise GrmontyError(f"{path}: not a grmonty_b200 binary spectrum")
    shape = (int(h["n_th"]), int(h["n_e"]), int(h["n_fields"]))
    body = np.frombuffer(raw[SPECTRUM_BIN_DTYPE.itemsize:], dtype="<f8")
    if body.size != shape[0] * shape[1] * shape[2]:
        raise GrmontyError(f"{path}: truncated binary spectrum")
    out = {k: h[k].item() for k in SPECTRUM_BIN_DTYPE.names if k != "magic"}
    out["spectrum"] = body.reshape(shape).copy()
    return out


class HarmModel:
    """Python face of the C++ host object (cuda-grmonty_b200/host/harm_model.hpp), same call order as the
    reference main: HarmModel(photon_n, mass_unit) -> read_file -> init -> run_simulation -> report_spectrum."""

    STATS = ["created", "scattered", "recorded", "max_tau_scatt", "seconds", "kernel_ms", "transport_ms",
             "n_tracked", "n_steps", "n_push_attempts", "n_interactions", "n_scatter_events", "n_generations",
             "n_kernel_launches", "luminosity", "max_tau_scatt_reported"]

    def __init__(self, photon_n: int, mass_unit: float, verbosity: int = 6):
        self.H = host_lib()
        self.h = C.c_void_p(self.H.gmh_create(int(photon_n), float(mass_unit), verbosity))

    def _ck(self, rc):
        if rc != 0:
            raise GrmontyError(self.H.gmh_last_error().decode())

    def close(self):
        if self.h:
            self.H.gmh_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def read_file(self, path: str):
        self._ck(self.H.gmh_read_file(self.h, path.encode()))

    def set_dump_cache(self, on: bool = True, directory: str = ""):
        """Binary dump cache `<dump>.b200cache`: loaded by read_file when it matches the text dump,
        written after parsing otherwise."""
        self.H.gmh_set_dump_cache(self.h, int(bool(on)), directory.encode())

    def set_device_tables(self, on: bool = True):
        """init() builds the geometry / weight / nint / hot cross-section tables on the GPU."""
        self.H.gmh_set_device_tables(self.h, int(bool(on)))

    def set_hotcross_cache(self, path: str = ""):
        """on-disk copy of the grid-independent hot cross-section table (loaded if valid, else built and written)"""
        self.H.gmh_set_hotcross_cache(self.h, path.encode())

    def hotcross_from_cache(self) -> bool:
        return bool(self.H.gmh_hotcross_from_cache(self.h))

    def read_from_cache(self) -> bool:
        return bool(self.H.gmh_read_from_cache(self.h))

    def report_spectrum_binary(self, path: str):
        self._ck(self.H.gmh_report_spectrum_binary(self.h, path.encode()))

    def init(self, threads: int = 0):
        self._ck(self.H.gmh_init(self.h, threads))

    def init_stage(self, which: int, threads: int = 0):
        self._ck(self.H.gmh_init_stage(self.h, which, threads))

    def set_options(self, seed=123, rank=0, world=1, device=0, threads_per_block=0, blocks_per_sm=0,
                    queue_capacity=0, gen0=0, gen_cap=0, gen_budget=0, gen_fine_from=0, gen_fine_div=0,
                    nccl_comm=None, cuda_library=None):
        lib_path = (cuda_library or LIB_CUDA).encode()
        self.H.gmh_set_options(self.h, seed, rank, world, device, threads_per_block, blocks_per_sm, queue_capacity,
                               gen0, gen_cap, gen_budget, gen_fine_from, gen_fine_div, nccl_comm, lib_path)

    def set_gpus(self, gpus: int):
        """run_simulation() shards the run over GPUs 0..gpus-1 of this box (one host thread each) and all-reduces"""
        self.H.gmh_set_gpus(self.h, int(gpus))

    def set_external_reduce(self, on: bool = True):
        """world > 1 without an NCCL communicator: the caller sums the per-rank spectra itself"""
        self.H.gmh_set_external_reduce(self.h, int(bool(on)))

    def run_simulation(self):
        self._ck(self.H.gmh_run_simulation(self.h))

    def report_spectrum(self, path: str):
        self._ck(self.H.gmh_report_spectrum(self.h, path.encode()))

    def header_raw(self):
        a = np.zeros(26)
        self.H.gmh_get_header_raw(self.h, _ptr(a))
        return a

    def spectrum(self):
        a = np.zeros((N_TH, N_E, N_F))
        self.H.gmh_get_spectrum(self.h, _ptr(a))
        return a

    def set_spectrum(self, spec):
        a = _arr(spec).reshape(N_TH, N_E, N_F)
        self.H.gmh_set_spectrum(self.h, _ptr(a))

    def stats(self) -> dict:
        a = np.zeros(16)
        self.H.gmh_get_stats(self.h, _ptr(a))
        return dict(zip(self.STATS, a.tolist()))

    def model_dict(self) -> dict:
        """Everything grmonty_b200_create needs, as numpy arrays (same keys as tests/golden model files)."""
        h, u, s = np.zeros(13), np.zeros(8), np.zeros(3)
        self.H.gmh_get_header(self.h, _ptr(h))
        self.H.gmh_get_units(self.h, _ptr(u))
        self.H.gmh_get_scalars(self.h, _ptr(s))
        n0, n1 = int(h[0]), int(h[1])
        d = dict(n0=n0, n1=n1, x_start1=h[2], x_start2=h[3], dx1=h[4], dx2=h[5], dx3=h[6], x_stop1=h[7],
                 x_stop2=h[8], a=h[9], h_slope=h[10], r_0=h[11], gamma=h[12], mass_unit=u[0], l_unit=u[1],
                 t_unit=u[2], rho_unit=u[3], u_unit=u[4], b_unit=u[5], theta_e_unit=u[6], n_e_unit=u[7],
                 bias_norm=s[0], max_tau_scatt0=s[1], photon_n=s[2])
        c_me, c_cl, c_hbar = 9.1093826e-28, 2.99792458e10, 6.6260693e-27 / (2.0 * np.pi)
        d["d_tau_k"] = 2.0 * np.pi * d["l_unit"] / (c_me * c_cl * c_cl / c_hbar)
        d["x1_min"] = float(np.log(1.0 + np.sqrt(max(0.0, 1.0 - d["a"] ** 2))))
        for i, nm in enumerate(["k_rho", "u", "u_1", "u_2", "u_3", "b_1", "b_2", "b_3", "geom_det"]):
            a = np.zeros((n0, n1))
            if self.H.gmh_get_grid(self.h, i, _ptr(a)) == 0:
                d[nm] = a
        for i, (nm, n) in enumerate([("hotcross", 221 * 81), ("f", 201), ("k2", 201), ("weight", 201),
                                     ("nint", 20001), ("dndlnu_max", 20001)]):
            a = np.zeros(n)
            self.H.gmh_get_table(self.h, i, _ptr(a))
            d[nm] = a
        return d
