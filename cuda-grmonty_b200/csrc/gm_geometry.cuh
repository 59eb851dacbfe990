/*
 * gm_geometry.cuh -- modified Kerr-Schild geometry and the geodesic integrator, sm_100a device code.
 *
 * What the reference computes (CPU path, cuda_grmonty/harm_model.cpp):
 *   get_bl_coord :1632-1637   gcov_func :499-530   gcon_func :473-497   get_connection :1436-1569
 *   init_dkdlam :1571-1587    step_size :1620-1630 push_photon :1217-1289
 * How it is organised here (B200-first, not a translation):
 *   - one GeoPoint per position holds exp/sincos results shared by the connection, the metric row needed
 *     for the energy check and (in the interaction stage) the full metric: 1 exp + 2 sincos per position
 *     instead of 3 exp + 6 sincos;
 *   - the connection is kept as 4x10 packed symmetric coefficients contracted against the 10 products
 *     k^j k^k; divisions are replaced by 4 reciprocals (fac2 == 2 rho^2 identically);
 *   - push_photon's recursion (depth <= 7) is flattened into "attempts" driven by a (position, level) pair so
 *     that every lane of a warp executes the same attempt code each iteration (see gm_transport.cuh).
 */
#pragma once
#include "gm_math.cuh"

/* GM_FUSED_RHS (default on): the push attempt evaluates -Gamma^i_jk k^j k^k straight from the geometry, row by
 * row, in each fixed-point iteration, instead of holding the 36 connection components across the two iterations.
 * Measured on B200: t_push_kernel 194 -> 165 registers, transport kernel spills 102 -> 40 bytes, run time -2.6 %
 * in spite of the recomputation. */
/* GM_ERRNORM_ONEDIV / GM_STEP_ONEDIV: sums of quotients over one common denominator -- one FP64 division instead of
 * four.  The results differ from the quotient-by-quotient forms by rounding only (a few ulp; every product stays far
 * inside the double range, see the comments at the two sites); measured 633 -> 618 ms per step for the error norm
 * with not one accept / halve decision of a 2.7e9-attempt run changed (identical recorded count); 618 -> 613 for the
 * step size. */
#ifndef GM_ERRNORM_ONEDIV
#define GM_ERRNORM_ONEDIV 1
#endif
#ifndef GM_STEP_ONEDIV
#define GM_STEP_ONEDIV 1
#endif
#ifndef GM_FUSED_RHS
#define GM_FUSED_RHS 1
#endif
#include "gm_params.h"

namespace gm {

struct GeoPoint {
    double r;        /* exp(x1): radius used by the connection (reference ignores r_0 there, :1438) */
    double rm;       /* r + r_0: radius used by the metric (:1634) */
    double sx, cx;   /* sin, cos(2 pi x2) */
    double sth, cth; /* sin, cos(theta), theta = pi x2 + (1-h)/2 sin(2 pi x2) */
    double hfac;     /* dtheta/dx2 = pi (1 + (1-h) cos(2 pi x2)) */
};

__device__ __forceinline__ GeoPoint geo_point(const GmParams &P, double x1, double x2) {
    GeoPoint g;
    /* branch-free evaluations (gm_math.cuh): exp and sincospi interleave, sincos follows */
    g.r = fm::exp_bounded(x1); /* x1 = ln r of a point on or near the grid; NaN stays NaN */
    g.rm = g.r + P.r_0;
    fm::sincospi_(2.0 * x2, &g.sx, &g.cx);
    const double omh = 1.0 - P.h_slope;
    const double th = kPi * x2 + 0.5 * omh * g.sx;
    fm::sincos_(th, &g.sth, &g.cth);
    g.hfac = kPi * (1.0 + omh * g.cx);
    return g;
}

/* the seven independent non-zero covariant components */
struct MetricCov {
    double g00, g01, g03, g11, g13, g22, g33;
};

__device__ __forceinline__ MetricCov metric_cov(const GmParams &P, const GeoPoint &q) {
    const double s = fabs(q.sth) + kEps;
    const double s2 = s * s;
    const double a = P.a;
    const double r = q.rm;
    const double rho2 = r * r + a * a * q.cth * q.cth;
    const double rfac = r - P.r_0;
    const double tr = fm::div(2.0 * r, rho2);
    MetricCov m;
    m.g00 = -1.0 + tr;
    m.g01 = tr * rfac;
    m.g03 = -a * s2 * tr;
    m.g11 = (1.0 + tr) * rfac * rfac;
    m.g13 = -a * s2 * (1.0 + tr) * rfac;
    m.g22 = rho2 * q.hfac * q.hfac;
    m.g33 = s2 * (rho2 + a * a * s2 * (1.0 + tr));
    return m;
}

/* row 0 only: what push_photon needs for e = -k^mu g_{0 mu} (the reference GPU build has gcov_0_func) */
__device__ __forceinline__ void metric_cov_row0(const GmParams &P, const GeoPoint &q, double &g00, double &g01,
                                                double &g03) {
    const double s = fabs(q.sth) + kEps;
    const double r = q.rm;
    const double rho2 = r * r + P.a * P.a * q.cth * q.cth;
    const double tr = fm::div(2.0 * r, rho2);
    g00 = -1.0 + tr;
    g01 = tr * (r - P.r_0);
    g03 = -P.a * (s * s) * tr;
}

/* contravariant components g^00, g^01 (g^02 = g^03 = 0) needed by get_fluid_params; full set for tests */
struct MetricCon {
    double g00, g01, g11, g13, g22, g33;
};

__device__ __forceinline__ MetricCon metric_con(const GmParams &P, const GeoPoint &q) {
    const double s = fabs(q.sth) + kEps;
    const double r = q.rm;
    const double a = P.a;
    const double irho2 = fm::rcp(r * r + a * a * q.cth * q.cth);
    MetricCon m;
    m.g00 = -1.0 - 2.0 * r * irho2;
    m.g01 = 2.0 * irho2;
    m.g11 = fm::div(irho2 * (r * (r - 2.0) + a * a), r * r);
    m.g13 = fm::div(a * irho2, r);
    m.g22 = fm::div(irho2, q.hfac * q.hfac);
    m.g33 = fm::div(irho2, s * s);
    return m;
}

/* packed symmetric index of (j,k), j <= k: 00 01 02 03 11 12 13 22 23 33 */
enum { S00 = 0, S01, S02, S03, S11, S12, S13, S22, S23, S33 };

/* Gamma^i_{jk} in packed form G[i][S..]; G[1][S02] = G[1][S23] = G[2][S02] = G[2][S23] = 0 are not stored
 * as variables the contraction reads (they are skipped below). */
struct Connection {
    double G0[10], G1[10], G2[10], G3[10];
};

__device__ __forceinline__ void connection_eval(const GmParams &P, const GeoPoint &q, Connection &c) {
    const double r1 = q.r;
    const double r2 = r1 * r1, r3 = r2 * r1, r4 = r2 * r2;
    const double omh = 1.0 - P.h_slope;
    const double dth = q.hfac;
    const double d2th = -2.0 * kPi * kPi * omh * q.sx;
    const double dth2 = dth * dth;
    const double sth = q.sth, cth = q.cth;
    const double sth2 = sth * sth, cth2 = cth * cth;
    const double sth4 = sth2 * sth2, cth4 = cth2 * cth2;
    const double r1sth2 = r1 * sth2;
    const double cs = cth * sth;
    const double s2th = 2.0 * cs;
    const double a = P.a, a2 = a * a, a3 = a2 * a, a4 = a2 * a2;
    const double a2sth2 = a2 * sth2, a2cth2 = a2 * cth2, a4cth4 = a4 * cth4;
    const double rho2 = r2 + a2cth2;
    const double rho22 = rho2 * rho2, rho23 = rho22 * rho2;
    const double irho2 = fm::rcp(rho2);
    const double irho22 = irho2 * irho2, irho23 = irho22 * irho2;
    const double idth = fm::rcp(dth);
    const double ir1 = fm::rcp(r1);
    const double isth = fm::rcp(sth);
    const double irho23_dth = irho23 * idth;
    const double fac1 = r2 - a2cth2;
    const double fac1_rho23 = fac1 * irho23;
    const double fac2 = 2.0 * rho2; /* a^2 + 2 r^2 + a^2 cos(2 theta) */
    const double fac3 = a2 + r1 * (r1 - 2.0);

    c.G0[S00] = 2.0 * r1 * fac1_rho23;
    c.G0[S01] = r1 * (2.0 * r1 + rho2) * fac1_rho23;
    c.G0[S02] = -a2 * r1 * s2th * dth * irho22;
    c.G0[S03] = -2.0 * a * r1sth2 * fac1_rho23;
    c.G0[S11] = 2.0 * r2 * (r4 + r1 * fac1 - a4cth4) * irho23;
    c.G0[S12] = r1 * c.G0[S02];
    c.G0[S13] = a * r1 * (-r1 * (r3 + 2.0 * fac1) + a4cth4) * sth2 * irho23;
    c.G0[S22] = -2.0 * r2 * dth2 * irho2;
    c.G0[S23] = a3 * r1sth2 * s2th * dth * irho22;
    c.G0[S33] = 2.0 * r1sth2 * (-r1 * rho22 + a2sth2 * fac1) * irho23;

    c.G1[S00] = fac3 * fac1_rho23 * ir1;
    c.G1[S01] = fac1 * (-2.0 * r1 + a2sth2) * irho23;
    c.G1[S02] = 0.0;
    c.G1[S03] = -a * sth2 * c.G1[S00];
    c.G1[S11] = (r4 * (r1 - 2.0) * (1.0 + r1) +
                 a2 * (a2 * r1 * (1.0 + 3.0 * r1) * cth4 + a4cth4 * cth2 + r3 * sth2 +
                       r1 * cth2 * (2.0 * r1 + 3.0 * r3 - a2sth2))) *
                irho23;
    c.G1[S12] = -a2 * dth * s2th * (0.5 * irho2);
    c.G1[S13] = a * sth2 *
                (a4 * r1 * cth4 + r2 * (2.0 * r1 + r3 - a2sth2) + a2cth2 * (2.0 * r1 * (r2 - 1.0) + a2sth2)) *
                irho23;
    c.G1[S22] = -fac3 * dth2 * irho2;
    c.G1[S23] = 0.0;
    c.G1[S33] = -fac3 * sth2 * (r1 * rho22 - a2sth2 * fac1) * irho23 * ir1;

    c.G2[S00] = -a2 * r1 * s2th * irho23_dth;
    c.G2[S01] = r1 * c.G2[S00];
    c.G2[S02] = 0.0;
    c.G2[S03] = a * r1 * (a2 + r2) * s2th * irho23_dth;
    c.G2[S11] = r2 * c.G2[S00];
    c.G2[S12] = r2 * irho2;
    c.G2[S13] = (a * r1 * cs * (r3 * (2.0 + r1) + a2 * (2.0 * r1 * (1.0 + r1) * cth2 + a2 * cth4 + 2.0 * r1sth2))) *
                irho23_dth;
    c.G2[S22] = -a2 * cs * dth * irho2 + d2th * idth;
    c.G2[S23] = 0.0;
    c.G2[S33] = -cs * (rho23 + a2sth2 * rho2 * (r1 * (4.0 + r1) + a2cth2) + 2.0 * r1 * a4 * sth4) * irho23_dth;

    const double cot = cth * isth;
    c.G3[S00] = a * fac1_rho23;
    c.G3[S01] = r1 * c.G3[S00];
    c.G3[S02] = -2.0 * a * r1 * cot * dth * irho22;
    c.G3[S03] = -a2sth2 * fac1_rho23;
    c.G3[S11] = r2 * c.G3[S00];
    /* a^2 + 2 r (2 + r) + a^2 cos(2 theta) = fac2 + 4 r ; 1/fac2^2 = irho22 / 4 */
    c.G3[S12] = -2.0 * a * r1 * (fac2 + 4.0 * r1) * cot * dth * (0.25 * irho22);
    c.G3[S13] = r1 * (r1 * rho22 - a2sth2 * fac1) * irho23;
    c.G3[S22] = -a * r1 * dth2 * irho2;
    c.G3[S23] = dth * (0.25 * fac2 * fac2 * cot + a2 * r1 * s2th) * irho22;
    c.G3[S33] = (-a * r1sth2 * rho22 + a3 * sth4 * fac1) * irho23;
}

/* dk^i/dlambda = -Gamma^i_{jk} k^j k^k (reference harm_model.cpp:1255-1262, :1578-1586) */
__device__ __forceinline__ void geodesic_rhs(const Connection &c, const double k[4], double dk[4]) {
    const double k00 = k[0] * k[0], k01 = 2.0 * k[0] * k[1], k02 = 2.0 * k[0] * k[2], k03 = 2.0 * k[0] * k[3];
    const double k11 = k[1] * k[1], k12 = 2.0 * k[1] * k[2], k13 = 2.0 * k[1] * k[3];
    const double k22 = k[2] * k[2], k23 = 2.0 * k[2] * k[3], k33 = k[3] * k[3];
    dk[0] = -(c.G0[S00] * k00 + c.G0[S01] * k01 + c.G0[S02] * k02 + c.G0[S03] * k03 + c.G0[S11] * k11 +
              c.G0[S12] * k12 + c.G0[S13] * k13 + c.G0[S22] * k22 + c.G0[S23] * k23 + c.G0[S33] * k33);
    dk[1] = -(c.G1[S00] * k00 + c.G1[S01] * k01 + c.G1[S03] * k03 + c.G1[S11] * k11 + c.G1[S12] * k12 +
              c.G1[S13] * k13 + c.G1[S22] * k22 + c.G1[S33] * k33);
    dk[2] = -(c.G2[S00] * k00 + c.G2[S01] * k01 + c.G2[S03] * k03 + c.G2[S11] * k11 + c.G2[S12] * k12 +
              c.G2[S13] * k13 + c.G2[S22] * k22 + c.G2[S33] * k33);
    dk[3] = -(c.G3[S00] * k00 + c.G3[S01] * k01 + c.G3[S02] * k02 + c.G3[S03] * k03 + c.G3[S11] * k11 +
              c.G3[S12] * k12 + c.G3[S13] * k13 + c.G3[S22] * k22 + c.G3[S23] * k23 + c.G3[S33] * k33);
}

/* dk^i/dlambda straight from the geometry: the same expressions as connection_eval + geodesic_rhs, but every
 * component is consumed as soon as it is formed, one row at a time, so the 36 components never have to be live
 * together (GM_FUSED_RHS: -72 registers in the push phase, +~300 instructions for a second fixed-point iteration). */
__device__ __forceinline__ void geodesic_rhs_direct(const GmParams &P, const GeoPoint &q, const double k[4],
                                                    double dk[4]) {
    const double k00 = k[0] * k[0], k01 = 2.0 * k[0] * k[1], k02 = 2.0 * k[0] * k[2], k03 = 2.0 * k[0] * k[3];
    const double k11 = k[1] * k[1], k12 = 2.0 * k[1] * k[2], k13 = 2.0 * k[1] * k[3];
    const double k22 = k[2] * k[2], k23 = 2.0 * k[2] * k[3], k33 = k[3] * k[3];
    const double r1 = q.r;
    const double r2 = r1 * r1, r3 = r2 * r1, r4 = r2 * r2;
    const double omh = 1.0 - P.h_slope;
    const double dth = q.hfac;
    const double d2th = -2.0 * kPi * kPi * omh * q.sx;
    const double dth2 = dth * dth;
    const double sth = q.sth, cth = q.cth;
    const double sth2 = sth * sth, cth2 = cth * cth;
    const double sth4 = sth2 * sth2, cth4 = cth2 * cth2;
    const double r1sth2 = r1 * sth2;
    const double cs = cth * sth;
    const double s2th = 2.0 * cs;
    const double a = P.a, a2 = a * a, a3 = a2 * a, a4 = a2 * a2;
    const double a2sth2 = a2 * sth2, a2cth2 = a2 * cth2, a4cth4 = a4 * cth4;
    const double rho2 = r2 + a2cth2;
    const double rho22 = rho2 * rho2, rho23 = rho22 * rho2;
    const double irho2 = fm::rcp(rho2);
    const double irho22 = irho2 * irho2, irho23 = irho22 * irho2;
    const double idth = fm::rcp(dth);
    const double ir1 = fm::rcp(r1);
    const double isth = fm::rcp(sth);
    const double irho23_dth = irho23 * idth;
    const double fac1 = r2 - a2cth2;
    const double fac1_rho23 = fac1 * irho23;
    const double fac2 = 2.0 * rho2;
    const double fac3 = a2 + r1 * (r1 - 2.0);
    const double cot = cth * isth;
    {
        const double g02 = -a2 * r1 * s2th * dth * irho22;
        double acc = (2.0 * r1 * fac1_rho23) * k00;
        acc = fma(r1 * (2.0 * r1 + rho2) * fac1_rho23, k01, acc);
        acc = fma(g02, k02, acc);
        acc = fma(-2.0 * a * r1sth2 * fac1_rho23, k03, acc);
        acc = fma(2.0 * r2 * (r4 + r1 * fac1 - a4cth4) * irho23, k11, acc);
        acc = fma(r1 * g02, k12, acc);
        acc = fma(a * r1 * (-r1 * (r3 + 2.0 * fac1) + a4cth4) * sth2 * irho23, k13, acc);
        acc = fma(-2.0 * r2 * dth2 * irho2, k22, acc);
        acc = fma(a3 * r1sth2 * s2th * dth * irho22, k23, acc);
        acc = fma(2.0 * r1sth2 * (-r1 * rho22 + a2sth2 * fac1) * irho23, k33, acc);
        dk[0] = -acc;
    }
    {
        const double g00 = fac3 * fac1_rho23 * ir1;
        double acc = g00 * k00;
        acc = fma(fac1 * (-2.0 * r1 + a2sth2) * irho23, k01, acc);
        acc = fma(-a * sth2 * g00, k03, acc);
        acc = fma((r4 * (r1 - 2.0) * (1.0 + r1) +
                   a2 * (a2 * r1 * (1.0 + 3.0 * r1) * cth4 + a4cth4 * cth2 + r3 * sth2 +
                         r1 * cth2 * (2.0 * r1 + 3.0 * r3 - a2sth2))) *
                      irho23,
                  k11, acc);
        acc = fma(-a2 * dth * s2th * (0.5 * irho2), k12, acc);
        acc = fma(a * sth2 *
                      (a4 * r1 * cth4 + r2 * (2.0 * r1 + r3 - a2sth2) + a2cth2 * (2.0 * r1 * (r2 - 1.0) + a2sth2)) *
                      irho23,
                  k13, acc);
        acc = fma(-fac3 * dth2 * irho2, k22, acc);
        acc = fma(-fac3 * sth2 * (r1 * rho22 - a2sth2 * fac1) * irho23 * ir1, k33, acc);
        dk[1] = -acc;
    }
    {
        const double g00 = -a2 * r1 * s2th * irho23_dth;
        double acc = g00 * k00;
        acc = fma(r1 * g00, k01, acc);
        acc = fma(a * r1 * (a2 + r2) * s2th * irho23_dth, k03, acc);
        acc = fma(r2 * g00, k11, acc);
        acc = fma(r2 * irho2, k12, acc);
        acc = fma((a * r1 * cs * (r3 * (2.0 + r1) + a2 * (2.0 * r1 * (1.0 + r1) * cth2 + a2 * cth4 + 2.0 * r1sth2))) *
                      irho23_dth,
                  k13, acc);
        acc = fma(-a2 * cs * dth * irho2 + d2th * idth, k22, acc);
        acc = fma(-cs * (rho23 + a2sth2 * rho2 * (r1 * (4.0 + r1) + a2cth2) + 2.0 * r1 * a4 * sth4) * irho23_dth, k33,
                  acc);
        dk[2] = -acc;
    }
    {
        const double g00 = a * fac1_rho23;
        double acc = g00 * k00;
        acc = fma(r1 * g00, k01, acc);
        acc = fma(-2.0 * a * r1 * cot * dth * irho22, k02, acc);
        acc = fma(-a2sth2 * fac1_rho23, k03, acc);
        acc = fma(r2 * g00, k11, acc);
        acc = fma(-2.0 * a * r1 * (fac2 + 4.0 * r1) * cot * dth * (0.25 * irho22), k12, acc);
        acc = fma(r1 * (r1 * rho22 - a2sth2 * fac1) * irho23, k13, acc);
        acc = fma(-a * r1 * dth2 * irho2, k22, acc);
        acc = fma(dth * (0.25 * fac2 * fac2 * cot + a2 * r1 * s2th) * irho22, k23, acc);
        acc = fma((-a * r1sth2 * rho22 + a3 * sth4 * fac1) * irho23, k33, acc);
        dk[3] = -acc;
    }
}

__device__ __forceinline__ void init_dkdlam(const GmParams &P, const double x[4], const double k[4], double dk[4]) {
    const GeoPoint q = geo_point(P, x[1], x[2]);
    Connection c;
    connection_eval(P, q, c);
    geodesic_rhs(c, k, dk);
}

/* reference harm_model.cpp:1620-1630; 1/(A/B + eps) is evaluated as B/(A + eps B): 4 divisions instead of 7 */
__device__ __forceinline__ double step_size(const GmParams &P, const double x[4], const double k[4]) {
    const double b1 = fabs(k[1]) + kEps, b2 = fabs(k[2]) + kEps, b3 = fabs(k[3]) + kEps;
    const double a1 = fabs(kStepEps * x[1]);
    const double a2 = fabs(kStepEps * fm::min_(x[2], P.x_stop2 - x[2]));
    const double a3 = kStepEps;
    const double d1 = a1 + kEps * b1, d2 = a2 + kEps * b2, d3 = a3 + kEps * b3;
#if GM_STEP_ONEDIV
    /* 1 / (b1/d1 + b2/d2 + b3/d3) = d1 d2 d3 / (b1 d2 d3 + b2 d1 d3 + b3 d1 d2); d_k >= 1e-80, so the products stay
     * above 1e-240 */
    const double d23 = d2 * d3;
    return fm::div(d1 * d23, fma(b1, d23, d1 * fma(b2, d3, b3 * d2)));
#else
    const double i1 = fm::div(b1, d1);
    const double i2 = fm::div(b2, d2);
    const double i3 = fm::div(b3, d3);
    return fm::rcp(i1 + i2 + i3);
#endif
}

/* |a - b| / |b + eps| for the fixed-point convergence test */
__device__ __forceinline__ double rel_change(double a, double b) { return fabs(fm::div(a - b, b + kEps)); }

/* The error norm sum_i |kp_i - kn_i| / |kn_i + eps| of the fixed-point iteration only feeds the threshold e_tol.
 * Measured alternatives that were rejected (profiles/r1_ab_microopts.txt): approximate reciprocals
 * (rcp.approx.ftz.f64) are 4 % faster but flip about one accept / halve decision in 1e8 against the oracle; the same
 * with an exact redo near the threshold is slower than the exact form (spills). */
__device__ __forceinline__ double err_norm(const double kp[4], const double kn[4]) {
#if GM_ERRNORM_ONEDIV
    /* sum_i |n_i| / |d_i| over one common denominator: one division instead of four (differs from the four-quotient
     * form by rounding only) */
    const double d0 = fabs(kn[0] + kEps), d1 = fabs(kn[1] + kEps), d2 = fabs(kn[2] + kEps), d3 = fabs(kn[3] + kEps);
    const double n0 = fabs(kp[0] - kn[0]), n1 = fabs(kp[1] - kn[1]), n2 = fabs(kp[2] - kn[2]), n3 = fabs(kp[3] - kn[3]);
    const double d01 = d0 * d1, d23 = d2 * d3;
    const double num = fma(fma(n0, d1, n1 * d0), d23, fma(n2, d3, n3 * d2) * d01);
    return fm::div(num, d01 * d23);
#else
    return ((rel_change(kp[0], kn[0]) + rel_change(kp[1], kn[1])) + rel_change(kp[2], kn[2])) +
           rel_change(kp[3], kn[3]);
#endif
}

/* One push_photon attempt of size dl from (x,k,dk) (reference harm_model.cpp:1230-1277): half kick, drift,
 * connection at the new point, <= 2 fixed-point iterations, energy check.  Returns true if the attempt must
 * be rejected and halved (the caller applies the depth limit).  Outputs are written to xn/kn/dkn/e1. */
__device__ __forceinline__ bool push_attempt(const GmParams &P, const double x[4], const double k[4],
                                             const double dk[4], double dl, double e_0_s, double xn[4],
                                             double kn[4], double dkn[4], double &e1, GeoPoint &q) {
    const double dl_2 = 0.5 * dl;
    double kh[4], kp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double d = dk[i] * dl_2;
        kh[i] = k[i] + d;
        kp[i] = kh[i] + d;
        xn[i] = x[i] + kh[i] * dl;
    }
    q = geo_point(P, xn[1], xn[2]);
#if !GM_FUSED_RHS
    Connection c;
    connection_eval(P, q, c);
#endif
    /* <= kMaxIter = 2 fixed-point iterations (reference :1247-1267) as a real loop: one copy of the contraction and
     * of the error norm in the instruction stream (the loop body is the largest piece of the hot code, which
     * competes for the 32 KB instruction cache) */
    double err = 0.0;
#pragma unroll 1
    for (int it = 0; it < kMaxIter; ++it) {
#if GM_FUSED_RHS
        GeoPoint qq = q;
        asm volatile("" : "+d"(qq.r), "+d"(qq.sth), "+d"(qq.cth)); /* opaque per iteration: no hoisting */
        geodesic_rhs_direct(P, qq, kp, dkn);
#else
        geodesic_rhs(c, kp, dkn);
#endif
#pragma unroll
        for (int i = 0; i < 4; ++i)
            kn[i] = kh[i] + dl_2 * dkn[i];
        err = err_norm(kp, kn);
        if (!(err > kETol))
            break;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            kp[i] = kn[i];
    }
    double g00, g01, g03;
    metric_cov_row0(P, q, g00, g01, g03);
    e1 = -(kn[0] * g00 + kn[1] * g01 + kn[3] * g03);
    const double err_e = fabs(fm::div(e1 - e_0_s, e_0_s));
    /* !(err <= tol) also catches NaN; isinf(err) > tol anyway (reference :1279) */
    return (err_e > 1.0e-4) || !(err <= kETol);
}

/* Halving bookkeeping shared by the flattened transport loop and the stand-alone full push.
 * A step of size dl is a binary tree of depth <= 7; `pos` counts completed 1/128ths, `level` is the depth
 * of the attempt to make at `pos`.  After an accepted attempt at (pos, level) the recursion of the reference
 * returns to the nearest ancestor whose second half has not run: level' = 7 - ctz(pos'). */
__device__ __forceinline__ int halving_next_level(int pos) { return kMaxHalvings - (__ffs(pos) - 1); }

/* complete push_photon (used for the scatter back-up and by tests); returns the number of attempts */
__device__ __forceinline__ int push_photon_full(const GmParams &P, double x[4], double k[4], double dk[4],
                                                double &e_0_s, double dl) {
    int pos = 0, level = 0, attempts = 0;
    while (pos < 128) {
        if (x[1] < P.x_start1) { /* reference :1218-1220: silent no-op */
            pos += 128 >> level;
            level = (pos < 128) ? halving_next_level(pos) : 0;
            continue;
        }
        double xn[4], kn[4], dkn[4], e1;
        GeoPoint q;
        const bool fail = push_attempt(P, x, k, dk, ldexp(dl, -level), e_0_s, xn, kn, dkn, e1, q);
        ++attempts;
        if (fail && level < kMaxHalvings) {
            ++level;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                x[i] = xn[i];
                k[i] = kn[i];
                dk[i] = dkn[i];
            }
            e_0_s = e1;
            pos += 128 >> level;
            level = (pos < 128) ? halving_next_level(pos) : 0;
        }
    }
    return attempts;
}

} /* namespace gm */
