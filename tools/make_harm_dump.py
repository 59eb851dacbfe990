#!/usr/bin/env python
"""Synthetic dump019-shaped HARM dump generator (there is no network, so no real dump019).

Writes the text format the reference loader parses (reference harm_model.cpp:99-204 and the
fixture writers in tests/harm_model_test.cpp:224-262): one header line of 26 fields, then
n0*n1 lines (i outer, j inner) of 34 numbers

    x1 x2 r th rho u u1 u2 u3 B1 B2 B3 divB ucon[4] ucov[4] bcon[4] bcov[4] vmin0 vmax0 vmin1 vmax1 gdet

Shape defaults follow BASELINE.json: 192x192 axisymmetric grid, a = 0.9375, modified
Kerr-Schild coordinates x1 = ln r, theta = pi x2 + (1-h)/2 sin(2 pi x2), h = 0.3, r_0 = 0.

The fluid is a deterministic torus-like profile (hard part H2):
a cool dense torus that produces most of the emission plus a hot tenuous corona that fills most of
the volume (so that bias_norm, the gdet-weighted mean of theta_e^2, is dominated by the corona and
the scattering bias in the emitting torus stays O(1)).  The same file is read by the reference CPU
build and by the B200 path, so physical realism only matters for exercising all code paths.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

GAMMA = 13.0 / 9.0
TP_OVER_TE = 3.0
MP_OVER_ME = 1.67262171e-24 / 9.1093826e-28


def theta_e_unit(gamma: float = GAMMA) -> float:
    """reference harm_model.cpp:139-141"""
    two_temp_gamma = 0.5 * ((1.0 + 2.0 / 3.0 * (TP_OVER_TE + 1.0) / (TP_OVER_TE + 2.0)) + gamma)
    return (two_temp_gamma - 1.0) * MP_OVER_ME / (1.0 + TP_OVER_TE)


def mks_metric(x1, x2, a, hslope):
    """Covariant / contravariant MKS Kerr metric on arrays (r_0 = 0). Returns gcov[4][4], gcon[4][4], gdet."""
    r = np.exp(x1)
    th = np.pi * x2 + 0.5 * (1.0 - hslope) * np.sin(2.0 * np.pi * x2)
    sth = np.abs(np.sin(th)) + 1e-40
    cth = np.cos(th)
    s2 = sth * sth
    rho2 = r * r + a * a * cth * cth
    rfac = r
    hfac = np.pi + (1.0 - hslope) * np.pi * np.cos(2.0 * np.pi * x2)
    z = np.zeros_like(r)
    gcov = [[z] * 4 for _ in range(4)]
    gcov[0][0] = -1.0 + 2.0 * r / rho2
    gcov[0][1] = (2.0 * r / rho2) * rfac
    gcov[0][3] = -2.0 * a * r * s2 / rho2
    gcov[1][0] = gcov[0][1]
    gcov[1][1] = (1.0 + 2.0 * r / rho2) * rfac * rfac
    gcov[1][3] = -a * s2 * (1.0 + 2.0 * r / rho2) * rfac
    gcov[2][2] = rho2 * hfac * hfac
    gcov[3][0] = gcov[0][3]
    gcov[3][1] = gcov[1][3]
    gcov[3][3] = s2 * (rho2 + a * a * s2 * (1.0 + 2.0 * r / rho2))
    irho2 = 1.0 / rho2
    gcon = [[z] * 4 for _ in range(4)]
    gcon[0][0] = -1.0 - 2.0 * r * irho2
    gcon[0][1] = 2.0 * irho2
    gcon[1][0] = gcon[0][1]
    gcon[1][1] = irho2 * (r * (r - 2.0) + a * a) / (r * r)
    gcon[1][3] = a * irho2 / r
    gcon[3][1] = gcon[1][3]
    gcon[2][2] = irho2 / (hfac * hfac)
    gcon[3][3] = irho2 / s2
    gdet = rho2 * sth * hfac * r
    return r, th, gcov, gcon, gdet


def make_dump(n0=192, n1=192, a=0.9375, hslope=0.3, r_out=40.0, profile="torus_c"):
    rh = 1.0 + np.sqrt(1.0 - a * a)
    r_in = 0.98 * rh
    x1_start = np.log(r_in)
    dx1 = np.log(r_out / r_in) / n0
    dx2 = 1.0 / n1
    i = np.arange(n0)[:, None] * np.ones((1, n1))
    j = np.ones((n0, 1)) * np.arange(n1)[None, :]
    x1 = x1_start + (i + 0.5) * dx1
    x2 = (j + 0.5) * dx2
    r, th, gcov, gcon, gdet = mks_metric(x1, x2, a, hslope)
    cth = np.cos(th)

    teu = theta_e_unit()
    if profile == "torus_b":  # SURVEY Appendix D variant b: uniform theta_e ~ 11.2
        rho = np.exp(-np.log(r / 12.0) ** 2 / (2 * 0.5**2)) * np.exp(-((cth / 0.35) ** 2)) + 1e-6 * r**-1.5
        uu = 0.05 * rho
    elif profile == "torus_c":
        # cool dense torus + hot tenuous corona/funnel
        torus = np.exp(-np.log(r / 10.0) ** 2 / (2 * 0.55**2)) * np.exp(-((cth / 0.30) ** 2))
        corona = 2.0e-3 * r**-1.3
        rho = torus + corona
        theta_torus = 8.0 * (r / 10.0) ** -0.5
        theta_corona = 45.0 * (r / 10.0) ** -0.25
        theta = (torus * theta_torus + corona * theta_corona) / rho
        uu = theta * rho / teu
    else:
        raise ValueError(profile)

    # primitive velocities (relative to the normal observer); any value gives a valid timelike u^mu
    v1 = -0.03 * r**-1.5
    v2 = np.zeros_like(r)
    v3 = 0.8 / (r**1.5 + a)
    vcon = [np.zeros_like(r), v1, v2, v3]
    vdotv = sum(gcov[m][n] * vcon[m] * vcon[n] for m in range(1, 4) for n in range(1, 4))
    vfac = np.sqrt(-1.0 / gcon[0][0] * (1.0 + np.abs(vdotv)))
    ucon = [-vfac * gcon[0][0]] + [vcon[m] - vfac * gcon[0][m] for m in range(1, 4)]
    ucov = [sum(gcov[m][n] * ucon[n] for n in range(4)) for m in range(4)]

    # magnetic field primitives: plasma beta ~ 30 in the torus, mostly toroidal + some radial
    pgas = (GAMMA - 1.0) * uu
    beta = 30.0
    bsq_target = 2.0 * pgas / beta
    # All three components are non-zero everywhere, as in a turbulent GRMHD dump.  (A field that is purely toroidal
    # anywhere makes the reference's scattering tetrad degenerate there -- make_tetrad orthogonalises d/dphi
    # against u and b, tetrads.cpp:46-123 -- and the scattered wave-vectors come out off the light cone: with a
    # radial component that changed sign across the equator, 60 % of the scatterings in the torus mid-plane did,
    # and a few of those "photons" orbit the hole for 4e5 steps.)
    B3 = 0.85 * np.sqrt(bsq_target / gcov[3][3])
    B1 = 0.40 * np.sqrt(bsq_target / gcov[1][1])
    B2 = 0.33 * np.sqrt(bsq_target / gcov[2][2]) * np.cos(np.pi * x2) ** 2 + 0.1 * np.sqrt(bsq_target / gcov[2][2])
    Bp = [np.zeros_like(r), B1, B2, B3]
    udotb = sum(ucov[m] * Bp[m] for m in range(1, 4))
    bcon = [udotb] + [(Bp[m] + ucon[m] * udotb) / ucon[0] for m in range(1, 4)]
    bcov = [sum(gcov[m][n] * bcon[n] for n in range(4)) for m in range(4)]

    header = [
        2000.0, n0, n1, x1_start, 0.0, dx1, dx2, 2000.0, 100000, a, GAMMA, 0.9, 10.0, 2.0, 2.0, 1000,
        19, 0, 2, 0.01, 2, 0, r_in, r_out, hslope, 0.0,
    ]
    zeros = np.zeros_like(r)
    cols = [x1, x2, r, th, rho, uu, v1, v2, v3, B1, B2, B3, zeros] + ucon + ucov + bcon + bcov + [
        zeros, zeros, zeros, zeros, gdet,
    ]
    table = np.stack([c.reshape(-1) for c in cols], axis=1)
    return header, table


def write_dump(path, header, table):
    with open(path, "w") as f:
        f.write(" ".join(repr(int(h)) if isinstance(h, (int, np.integer)) else "%.17g" % h for h in header) + "\n")
        np.savetxt(f, table, fmt="%.15g")


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--out", required=True)
    ap.add_argument("--n0", type=int, default=192)
    ap.add_argument("--n1", type=int, default=192)
    ap.add_argument("--a", type=float, default=0.9375)
    ap.add_argument("--hslope", type=float, default=0.3)
    ap.add_argument("--r_out", type=float, default=40.0)
    ap.add_argument("--profile", default="torus_c")
    args = ap.parse_args(argv)
    header, table = make_dump(args.n0, args.n1, args.a, args.hslope, args.r_out, args.profile)
    write_dump(args.out, header, table)
    print(f"wrote {args.out}: {args.n0}x{args.n1}, a={args.a}, profile={args.profile}", file=sys.stderr)


if __name__ == "__main__":
    main()
