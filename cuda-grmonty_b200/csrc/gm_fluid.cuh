/*
 * gm_fluid.cuh -- HARM grid interpolation and fluid-frame state, sm_100a device code.
 *
 * Reference: get_fluid_params harm_model.cpp:595-671, x_to_ij :1406-1434, interp_scalar :1646-1656,
 * get_fluid_zone :538-593.
 * Layout: the eight primitives of a zone are interleaved ([n0][n1][8] doubles = one 64-byte record per zone,
 * 16-byte vector loads through the read-only path) instead of the reference's eight separate arrays: a
 * bilinear lookup touches 4 records = 2 x 128 contiguous bytes instead of 32 scattered sectors.  At 192^2
 * the grid is 2.4 MB and at 1024^2 67 MB, both L2-resident on B200 (126 MB); the API additionally pins it
 * with an access-policy window.
 */
#pragma once
#include "gm_geometry.cuh"
#include "gm_params.h"

namespace gm {

struct Fluid {
    double n_e, theta_e, b;
    double u_con[4], u_cov[4], b_con[4], b_cov[4];
};

__device__ __forceinline__ void lower_sparse(const MetricCov &g, const double v[4], double o[4]) {
    o[0] = g.g00 * v[0] + g.g01 * v[1] + g.g03 * v[3];
    o[1] = g.g01 * v[0] + g.g11 * v[1] + g.g13 * v[3];
    o[2] = g.g22 * v[2];
    o[3] = g.g03 * v[0] + g.g13 * v[1] + g.g33 * v[3];
}

/* primitives (v^i, B^i) -> u^mu, u_mu, b^mu, b_mu, |B| ; reference harm_model.cpp:638-668 */
__device__ __forceinline__ void prims_to_fluid(const GmParams &P, const MetricCov &g, double gcon00,
                                               double gcon01, const double v[4], const double bp[4], Fluid &f) {
    const double v_dot_v =
        g.g11 * v[1] * v[1] + 2.0 * g.g13 * v[1] * v[3] + g.g22 * v[2] * v[2] + g.g33 * v[3] * v[3];
    const double v_fac = fm::sqrt_(-fm::rcp(gcon00) * (1.0 + fabs(v_dot_v)));
    f.u_con[0] = -v_fac * gcon00;
    f.u_con[1] = v[1] - v_fac * gcon01;
    f.u_con[2] = v[2];
    f.u_con[3] = v[3];
    lower_sparse(g, f.u_con, f.u_cov);
    const double u_dot_bp = f.u_cov[1] * bp[1] + f.u_cov[2] * bp[2] + f.u_cov[3] * bp[3];
    const double iu0 = fm::rcp(f.u_con[0]);
    f.b_con[0] = u_dot_bp;
    f.b_con[1] = (bp[1] + f.u_con[1] * u_dot_bp) * iu0;
    f.b_con[2] = (bp[2] + f.u_con[2] * u_dot_bp) * iu0;
    f.b_con[3] = (bp[3] + f.u_con[3] * u_dot_bp) * iu0;
    lower_sparse(g, f.b_con, f.b_cov);
    f.b = fm::sqrt_(f.b_con[0] * f.b_cov[0] + f.b_con[1] * f.b_cov[1] + f.b_con[2] * f.b_cov[2] +
                    f.b_con[3] * f.b_cov[3]) *
          P.b_unit;
}

/* Bilinear, cell-centred, clamped interpolation of the 8 primitives at (x1,x2).
 * Returns false (and f.n_e = 0) outside [x_start, x_stop] (reference :599-603). */
__device__ __forceinline__ bool fluid_params(const GmParams &P, double x1, double x2, const MetricCov &g,
                                             const GeoPoint &q, Fluid &f) {
    if (x1 < P.x_start1 || x1 > P.x_stop1 || x2 < P.x_start2 || x2 > P.x_stop2) {
        f.n_e = 0.0;
        return false;
    }
    const double qi = (x1 - P.x_start1) * P.inv_dx1, qj = (x2 - P.x_start2) * P.inv_dx2;
    int i = (int)(qi - 0.5 + 1000) - 1000;
    int j = (int)(qj - 0.5 + 1000) - 1000;
    double del_i, del_j;
    if (i < 0) {
        i = 0;
        del_i = 0.0;
    } else if (i > P.n0 - 2) {
        i = P.n0 - 2;
        del_i = 1.0;
    } else {
        del_i = (x1 - ((i + 0.5) * P.dx1 + P.x_start1)) * P.inv_dx1;
    }
    if (j < 0) {
        j = 0;
        del_j = 0.0;
    } else if (j > P.n1 - 2) {
        j = P.n1 - 2;
        del_j = 1.0;
    } else {
        del_j = (x2 - ((j + 0.5) * P.dx2 + P.x_start2)) * P.inv_dx2;
    }
    const double c00 = (1.0 - del_i) * (1.0 - del_j), c01 = (1.0 - del_i) * del_j;
    const double c10 = del_i * (1.0 - del_j), c11 = del_i * del_j;
    const double2 *z00 = reinterpret_cast<const double2 *>(P.grid + ((size_t)i * P.n1 + j) * 8);
    const double2 *z10 = z00 + (size_t)P.n1 * 4;
    double pr[8];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const double2 a = __ldg(z00 + v), b = __ldg(z00 + 4 + v), c = __ldg(z10 + v), d = __ldg(z10 + 4 + v);
        pr[2 * v] = a.x * c00 + b.x * c01 + c.x * c10 + d.x * c11;
        pr[2 * v + 1] = a.y * c00 + b.y * c01 + c.y * c10 + d.y * c11;
    }
    f.n_e = pr[0] * P.n_e_unit;
    f.theta_e = fm::div(pr[1], pr[0]) * P.theta_e_unit;
    const double v[4] = {0.0, pr[2], pr[3], pr[4]};
    const double bp[4] = {0.0, pr[5], pr[6], pr[7]};
    const MetricCon gc = metric_con(P, q);
    prims_to_fluid(P, g, gc.g00, gc.g01, v, bp, f);
    return true;
}

/* zone-centre state (reference get_fluid_zone :538-593 with get_coord :1639-1644) */
__device__ __forceinline__ void fluid_zone(const GmParams &P, int i, int j, double x[4], MetricCov &g, Fluid &f) {
    x[0] = 0.0;
    x[1] = P.x_start1 + (i + 0.5) * P.dx1;
    x[2] = P.x_start2 + (j + 0.5) * P.dx2;
    x[3] = 0.0;
    const GeoPoint q = geo_point(P, x[1], x[2]);
    g = metric_cov(P, q);
    const MetricCon gc = metric_con(P, q);
    const double *z = P.grid + ((size_t)i * P.n1 + j) * 8;
    f.n_e = z[0] * P.n_e_unit;
    f.theta_e = (z[1] / f.n_e) * P.n_e_unit * P.theta_e_unit;
    const double v[4] = {0.0, z[2], z[3], z[4]};
    const double bp[4] = {0.0, z[5], z[6], z[7]};
    prims_to_fluid(P, g, gc.g00, gc.g01, v, bp, f);
}

} /* namespace gm */
