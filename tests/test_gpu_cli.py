"""End-to-end run of the command-line driver (the reference's `main.cpp` surface) on a GPU box:
dump file in, spectrum file out, through libgrmonty_b200_host.so -> dlopen(libgrmonty_b200.so) -> C ABI."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_writes_the_reference_spectrum_format(tmp_path):
    import cuda_grmonty_b200 as gm
    from tools import make_harm_dump
    dump, spec = str(tmp_path / "dump48.txt"), str(tmp_path / "spectrum.txt")
    make_harm_dump.write_dump(dump, *make_harm_dump.make_dump(n0=48, n1=48))
    # flag spellings of the reference README: single-dash -photon_n, --flag=value and --flag value
    r = subprocess.run([gm.CLI, "--harm_dump_path", dump, f"--spectrum_path={spec}", "-photon_n", "2000",
                        "--mass_unit", "4e19", "--verbosity", "info"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    log = r.stdout + r.stderr
    assert "photon rate" in log.lower() or "rate" in log.lower()
    rows = np.loadtxt(spec)
    assert rows.shape == (200, 1 + 6 * 6)      # reference harm_model.cpp:433-457: 200 energy bins, 6 angle bins x 6 fields
    assert np.all(np.isfinite(rows))
    assert np.all(np.diff(rows[:, 0]) > 0)     # column 0: log10 of the bin energy, increasing
    nu_l_nu = rows[:, 1::6]
    assert (nu_l_nu >= 0).all() and nu_l_nu.sum() > 0
    # the same model through the Python binding gives the same luminosity scale (different generation of the run,
    # same seed 123 => identical counters)
    hm = gm.HarmModel(2000, 4e19)
    hm.read_file(dump)
    hm.init()
    hm.set_options(seed=123)
    hm.run_simulation()
    spec2 = str(tmp_path / "spectrum2.txt")
    hm.report_spectrum(spec2)
    assert open(spec).read() == open(spec2).read()
