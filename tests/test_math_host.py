"""Accuracy of the branch-free FP64 functions of cuda-grmonty_b200/csrc/gm_math.cuh, checked on the host:
the header is plain C++ when compiled by g++ (the two MUFU seed instructions are modelled by 21-bit truncations),
so the very code the kernels inline is compared with glibc's long-double functions on 2e6 samples per function."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# max error in ulp (sincos*: in units of 2^-53 absolute) the transport path's parity budget allows
LIMITS = {"rcp": 1.0, "div": 1.0, "sqrt": 1.0, "exp": 2.0, "exp10": 2.0, "log": 2.0, "sincospi": 2.0, "sincos": 2.0,
          "cbrt": 1.0,
          # bit-identical to exp_/exp10_ on their domain; fmin/fmax semantics (0 = no mismatch)
          "exp_bounded": 0.0, "exp10_bounded": 0.0, "minmax": 0.0,
          # sums of quotients over one common denominator vs the quotient-by-quotient forms (ulp of the result)
          "errnorm_onediv": 6.0, "stepsize_onediv": 6.0}


def test_math_header_accuracy(tmp_path):
    exe = str(tmp_path / "math_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "math_host_check.cpp"), "-lm"])
    out = subprocess.check_output([exe], text=True)
    seen = {}
    for line in out.splitlines():
        name, n, *errs = line.split()
        seen[name] = max(float(e) for e in errs)
        assert int(n) >= 1000000
    assert set(seen) == set(LIMITS)
    for name, err in seen.items():
        assert err <= LIMITS[name], (name, err)
