"""CPU tests of the C++ host (cuda-grmonty_b200/host): the parts of the reference surface that are kept --
HARM dump loader, table builders, spectrum file writer, CLI -- and the C-ABI library's exported symbols.

The dump-format test mirrors the reference's own harm_model_test.cpp (ReadFileHeader / ReadFileData) with the
same 2x3 fixture values (reference tests/harm_model_test.cpp:16-94, writers :224-262).
"""
import os
import subprocess

import numpy as np
import pytest

import cuda_grmonty_b200 as gm
from tools import make_harm_dump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    gm.build_host()


def write_fixture(path):
    """the reference test's sample header/data: header values 1, (2,3), 4, 5, 8, 9, 10 ... 28; data 111..823"""
    hdr = "1 2 3 4 5 8 9 10 11 12 13 14 15 16 17 18 19 20 21 22 23 24 25 26 27 28"
    lines = [hdr]
    for i in range(2):
        for j in range(3):
            prims = " ".join(str(100 * v + 10 * (i + 1) + (j + 1)) for v in range(1, 9))
            lines.append("0 0 0 0 " + prims + " 0 " + "0 0 0 0 0 0 0 0 " * 2 + "0 0 0 0 0")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def test_read_file_header_and_data(tmp_path):
    p = str(tmp_path / "harm_dump")
    write_fixture(p)
    m = gm.HarmModel(1000, 4e19)
    m.read_file(p)
    raw = m.header_raw()
    want = [1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28]
    assert np.array_equal(raw, np.array(want, dtype=float))
    d = m.model_dict()
    # derived header fields (reference harm_model.cpp:106-117)
    assert d["x_stop1"] == 4.0 + 2 * 8.0 and d["x_stop2"] == 5.0 + 3 * 9.0 and d["dx3"] == 2.0 * np.pi
    for v, nm in enumerate(["k_rho", "u", "u_1", "u_2", "u_3", "b_1", "b_2", "b_3"], start=1):
        want = np.array([[100 * v + 10 * (i + 1) + (j + 1) for j in range(3)] for i in range(2)], dtype=float)
        assert np.array_equal(d[nm], want), nm


def test_read_file_missing_raises(tmp_path):
    m = gm.HarmModel(1000, 4e19)
    with pytest.raises(gm.GrmontyError, match="File does not exist"):
        m.read_file(str(tmp_path / "nope"))


@pytest.fixture(scope="module")
def host48(tmp_path_factory):
    """host model of the same 48x48 synthetic dump the golden vectors were generated from"""
    p = str(tmp_path_factory.mktemp("dump") / "dump48.txt")
    header, table = make_harm_dump.make_dump(n0=48, n1=48)
    make_harm_dump.write_dump(p, header, table)
    m = gm.HarmModel(2000, 4e19)
    m.read_file(p)
    m.init()
    return m


def test_units_scalars_and_grids_match_reference(host48, golden_model):
    d = host48.model_dict()
    for k in ["mass_unit", "l_unit", "t_unit", "rho_unit", "u_unit", "b_unit", "theta_e_unit", "n_e_unit",
              "max_tau_scatt0", "d_tau_k", "x1_min", "x_start1", "dx1", "dx2", "x_stop1", "x_stop2", "a", "h_slope"]:
        assert d[k] == pytest.approx(golden_model[k], rel=1e-15), k
    assert d["bias_norm"] == pytest.approx(golden_model["bias_norm"], rel=1e-12)
    for k in ["k_rho", "u", "u_1", "u_2", "u_3", "b_1", "b_2", "b_3"]:
        # the dump text is regenerated here; values agree to the printed precision (%.15g)
        assert np.allclose(d[k], golden_model[k], rtol=1e-13, atol=0), k
    assert np.allclose(d["geom_det"], golden_model["geom_det"], rtol=1e-13)


def test_tables_match_reference(host48, golden_model):
    """table builders vs the reference's own tables (golden): hotcross and K2 restate the reference arithmetic
    (agreement ~1e-13); F(K) uses a different quadrature than the reference's adaptive GK61 (eps_rel 1e-6)."""
    d = host48.model_dict()
    assert np.max(np.abs(d["hotcross"] - golden_model["hotcross"])) < 1e-12      # log10 sigma
    assert np.max(np.abs(d["k2"] - golden_model["k2"])) < 1e-12                  # ln K2
    assert np.max(np.abs(d["f"] - golden_model["f"])) < 3e-6                     # ln F
    assert np.max(np.abs(d["weight"] - golden_model["weight"])) < 5e-6           # ln weight
    ok = np.isfinite(golden_model["nint"])
    assert np.array_equal(ok, np.isfinite(d["nint"]))
    assert np.max(np.abs(d["nint"][ok] - golden_model["nint"][ok])) < 1e-5
    ok = np.isfinite(golden_model["dndlnu_max"])
    assert np.max(np.abs(d["dndlnu_max"][ok] - golden_model["dndlnu_max"][ok])) < 1e-5


def test_tables_do_not_depend_on_thread_count(tmp_path):
    p = str(tmp_path / "d.txt")
    header, table = make_harm_dump.make_dump(n0=16, n1=16)
    make_harm_dump.write_dump(p, header, table)
    out = []
    for threads in (1, 5):
        m = gm.HarmModel(500, 4e19)
        m.read_file(p)
        for stage in (0, 2, 3, 4):  # skip the hot cross-section table (grid independent)
            m.init_stage(stage, threads)
        d = m.model_dict()
        out.append(np.concatenate([d["geom_det"].ravel(), d["f"], d["k2"], d["weight"], d["nint"]]))
    assert np.array_equal(out[0], out[1], equal_nan=True)


def test_spectrum_file_is_byte_identical_to_the_reference(host48, tmp_path):
    """report_spectrum on the same Spectrum array must produce the same bytes as the reference's
    (fixture: tests/golden/spectrum_file.npz, written by oracle/make_golden.py from the reference build)"""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "spectrum_file.npz"))
    host48.set_spectrum(fx["spectrum"])
    out = str(tmp_path / "spectrum.txt")
    host48.report_spectrum(out)
    got = open(out, "rb").read()
    assert got == fx["text"].tobytes()
    st = host48.stats()
    assert st["luminosity"] == pytest.approx(float(fx["luminosity"]), rel=1e-4)  # fixture value is from the 5-digit text


def test_cabi_library_exports_every_declared_symbol():
    """every function declared in include/grmonty_b200.h is exported by the CUDA library (no compute calls)"""
    import ctypes
    import re
    gm.build_cuda()
    for header, path, names in (("grmonty_b200.h", gm.LIB_CUDA, gm.ABI_SYMBOLS),
                                ("grmonty_b200_test.h", gm.LIB_CUDA_TEST, gm.TEST_ABI_SYMBOLS)):
        hdr = open(os.path.join(ROOT, "include", header)).read()
        declared = sorted(set(re.findall(r"\b(grmonty_b200_[a-z0-9_]+)\s*\(", hdr)))
        assert len(declared) >= 15
        L = ctypes.CDLL(path)
        missing = [s for s in declared if not hasattr(L, s)]
        assert not missing, missing
        assert sorted(names) == declared
    # the product library carries no test export; the test library is a superset of the product ABI
    prod, test = ctypes.CDLL(gm.LIB_CUDA), ctypes.CDLL(gm.LIB_CUDA_TEST)
    assert not [s for s in gm.TEST_ABI_SYMBOLS if hasattr(prod, s)]
    assert not [s for s in gm.ABI_SYMBOLS if not hasattr(test, s)]


def test_no_cpu_fallback(golden_model):
    """without a CUDA device the product fails loudly instead of computing on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gm.GrmontyError, match="no CUDA device"):
        gm.Context(golden_model)
    m = gm.HarmModel(10, 4e19)
    with pytest.raises(gm.GrmontyError):
        m.set_options()
        m.run_simulation()


def test_cli_flags(tmp_path):
    """the CLI accepts the reference's flag spellings; without a readable dump it exits non-zero"""
    r = subprocess.run([gm.CLI, "--harm_dump_path", str(tmp_path / "missing"), "--spectrum_path",
                        str(tmp_path / "s"), "-photon_n", "1000", "--mass_unit=4e19", "--verbosity", "error"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "File does not exist" in r.stderr


# ---- binary dump cache and the full-field binary spectrum --------------------------------------------

def _grids(m):
    d = m.model_dict()
    return {k: d[k] for k in ("k_rho", "u", "u_1", "u_2", "u_3", "b_1", "b_2", "b_3")}, d["bias_norm"], d["theta_e_unit"]


def test_dump_cache_round_trip_is_bit_exact(tmp_path):
    p = str(tmp_path / "dump32.txt")
    header, table = make_harm_dump.make_dump(n0=32, n1=24)
    make_harm_dump.write_dump(p, header, table)
    ref = gm.HarmModel(1000, 4e19)
    ref.read_file(p)                       # plain text parse, cache off (the default): no side file
    assert not os.path.exists(p + ".b200cache") and not ref.read_from_cache()
    a = gm.HarmModel(1000, 4e19)
    a.set_dump_cache(True)
    a.read_file(p)                         # parses, then writes the cache
    assert not a.read_from_cache() and os.path.exists(p + ".b200cache")
    b = gm.HarmModel(1000, 4e19)
    b.set_dump_cache(True)
    b.read_file(p)                         # loads the cache
    assert b.read_from_cache()
    assert np.array_equal(ref.header_raw(), b.header_raw())
    (g0, bn0, te0), (g1, bn1, te1) = _grids(ref), _grids(b)
    assert bn0 == bn1 and te0 == te1
    for k in g0:
        assert np.array_equal(g0[k], g1[k]), k
    # tables built from cached data are the same bits as from the parsed text
    ref.init_stage(0); b.init_stage(0)
    assert np.array_equal(ref.model_dict()["geom_det"], b.model_dict()["geom_det"])
    # size = header + 8 grids of raw doubles
    assert os.path.getsize(p + ".b200cache") == 8 + 8 + 8 * 5 + 8 + 26 * 8 + 8 * 32 * 24 * 8


def test_dump_cache_is_invalidated_by_a_changed_dump_or_a_bad_cache(tmp_path):
    p = str(tmp_path / "dump16.txt")
    cdir = tmp_path / "cache"
    cdir.mkdir()
    header, table = make_harm_dump.make_dump(n0=16, n1=16)
    make_harm_dump.write_dump(p, header, table)
    m = gm.HarmModel(1000, 4e19)
    m.set_dump_cache(True, str(cdir))
    m.read_file(p)
    cpath = str(cdir / "dump16.txt.b200cache")
    assert os.path.exists(cpath) and not os.path.exists(p + ".b200cache")
    m.read_file(p)
    assert m.read_from_cache()
    # another dump under the same name: different content (and size) -> the cache is ignored and rewritten
    table2 = table.copy()
    table2[:, 4] *= 2.0
    make_harm_dump.write_dump(p, header, table2)
    m.read_file(p)
    assert not m.read_from_cache()
    assert np.allclose(m.model_dict()["k_rho"].ravel(), table2[:, 4], rtol=1e-14)
    m.read_file(p)
    assert m.read_from_cache()
    # truncated / foreign cache files fall back to the text parser
    raw = open(cpath, "rb").read()
    flipped = bytearray(raw)
    flipped[-5] ^= 0x10                           # one bit of the payload: right size, wrong content
    for bad in (raw[:100], raw[:-8], b"XXXXXXXX" + raw[8:], b"", bytes(flipped)):
        with open(cpath, "wb") as f:
            f.write(bad)
        m.read_file(p)
        assert not m.read_from_cache()
        assert np.allclose(m.model_dict()["k_rho"].ravel(), table2[:, 4], rtol=1e-14)
        assert open(cpath, "rb").read() == raw      # and the cache is rewritten
    # an unwritable cache directory must not fail the read
    m.set_dump_cache(True, str(tmp_path / "does" / "not" / "exist"))
    m.read_file(p)
    assert not m.read_from_cache()


def test_binary_spectrum_keeps_all_thirteen_fields(host48, tmp_path):
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "spectrum_file.npz"))
    host48.set_spectrum(fx["spectrum"])
    p = str(tmp_path / "spectrum.bin")
    host48.report_spectrum_binary(p)
    got = gm.read_spectrum_binary(p)
    assert got["spectrum"].shape == (6, 200, 13) and np.array_equal(got["spectrum"], fx["spectrum"])
    assert got["n_th"] == 6 and got["n_e"] == 200 and got["n_fields"] == 13
    assert got["mass_unit"] == 4e19 and got["photon_n"] == 2000
    assert os.path.getsize(p) == gm.SPECTRUM_BIN_DTYPE.itemsize + 6 * 200 * 13 * 8
    with open(p, "r+b") as f:
        f.truncate(1000)
    with pytest.raises(gm.GrmontyError, match="truncated"):
        gm.read_spectrum_binary(p)
    with pytest.raises(gm.GrmontyError):
        host48.report_spectrum_binary(str(tmp_path / "no" / "such" / "dir" / "s.bin"))


def test_hotcross_table_disk_cache(host48, tmp_path):
    """the hot cross-section table depends only on the consts grid; a second model loads it from disk"""
    import time
    path = str(tmp_path / "hotcross.bin")
    a = gm.HarmModel(1000, 4e19)
    a.set_hotcross_cache(path)
    a.init_stage(1)
    assert not a.hotcross_from_cache() and os.path.getsize(path) == 8 + 16 + 32 + 8 + 221 * 81 * 8
    want = host48.model_dict()["hotcross"]
    b = gm.HarmModel(1000, 4e19)
    b.set_hotcross_cache(path)
    t0 = time.perf_counter()
    b.init_stage(1)
    assert b.hotcross_from_cache() and time.perf_counter() - t0 < 0.05
    assert np.array_equal(b.model_dict()["hotcross"], want) and np.array_equal(a.model_dict()["hotcross"], want)
    raw = open(path, "rb").read()
    flipped = bytearray(raw)
    flipped[-3] ^= 0x01
    # truncated, other version, foreign file, one flipped payload bit
    for bad in (raw[:-8], raw[:8] + b"\x07" + raw[9:], b"junk", bytes(flipped)):
        with open(path, "wb") as f:
            f.write(bad)
        c = gm.HarmModel(1000, 4e19)
        c.set_hotcross_cache(path)
        c.init_stage(1, 0)
        assert not c.hotcross_from_cache() and np.array_equal(c.model_dict()["hotcross"], want)
        assert open(path, "rb").read() == raw


def _read_with_cache(path):
    m = gm.HarmModel(1000, 4e19)
    m.set_dump_cache(True)
    m.read_file(path)
    d = m.model_dict()
    return float(d["bias_norm"]), float(np.asarray(d["b_3"]).sum()), float(np.asarray(d["k_rho"]).sum())


def test_dump_cache_with_concurrent_readers(tmp_path):
    """the ranks of a multi-GPU job read the same dump at the same time: each gets the right grids, the cache that
    is left behind is valid and no temporary file survives"""
    import multiprocessing as mp
    p = str(tmp_path / "dump64.txt")
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=64, n1=64))
    ref = gm.HarmModel(1000, 4e19)
    ref.read_file(p)
    d = ref.model_dict()
    want = (float(d["bias_norm"]), float(np.asarray(d["b_3"]).sum()), float(np.asarray(d["k_rho"]).sum()))
    with mp.get_context("spawn").Pool(6) as pool:
        got = pool.map(_read_with_cache, [p] * 12)
    assert all(g == want for g in got)
    assert sorted(os.listdir(tmp_path)) == ["dump64.txt", "dump64.txt.b200cache"]
    m = gm.HarmModel(1000, 4e19)
    m.set_dump_cache(True)
    m.read_file(p)
    assert m.read_from_cache()


def test_config_struct_matches_the_header(tmp_path):
    """the ctypes mirror of grmonty_b200_config / grmonty_b200_stats must have the size and ABI version a C compiler
    gives the structs of include/grmonty_b200.h (create() rejects a mismatch, but only on a GPU box)"""
    import ctypes as C
    import subprocess
    import cuda_grmonty_b200 as gm
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "grmonty_b200.h"\n'
                   'int main(void) { printf("%d %zu %zu\\n", GRMONTY_B200_ABI_VERSION, sizeof(grmonty_b200_config), '
                   'sizeof(grmonty_b200_stats)); return 0; }\n')
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    ver, cfg_size, stats_size = (int(v) for v in subprocess.check_output([str(exe)], text=True).split())
    assert ver == gm.ABI_VERSION
    assert cfg_size == C.sizeof(gm.Config)
    assert stats_size == C.sizeof(gm.Stats)


def test_bench_reads_the_roofline_traffic_from_the_committed_ncu_summary():
    """roofline.traffic is the dram__bytes_read + dram__bytes_write of the newest committed transport-kernel capture"""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    traffic, source = bench.ncu_traffic()
    assert traffic and traffic > 1e8
    assert source.startswith("profiles/r2_transport_ncu")
