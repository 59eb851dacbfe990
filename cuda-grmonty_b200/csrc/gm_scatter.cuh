/*
 * gm_scatter.cuh -- orthonormal tetrads, electron / Klein-Nishina / Thomson samplers and the Compton
 * scattering kernel body; sm_100a device code.
 *
 * Reference: tetrads.cpp:46-194 (make_tetrad, coordinate_to_tetrad, tetrad_to_coordinate, lower),
 * proba.cpp:30-215 (sample_electron_distr_p, sample_beta_distr, sample_y_distr, sample_mu_distr,
 * sample_klein_nishina, sample_thomson, sample_rand_dir), harm_model.cpp:1071-1215 (scatter_super_photon,
 * sample_scattered_photon), :1658-1671 (boost).
 * All random numbers come from the scattering photon's own Philox stream (gm_rng.cuh) in FP64 -- the
 * reference GPU build draws float uniforms/normals (proba.cuh:129,215,231).
 */
#pragma once
#include "gm_fluid.cuh"
#include "gm_params.h"
#include "gm_rng.cuh"

namespace gm {

__device__ __forceinline__ double dot_sparse(const MetricCov &g, const double a[4], const double b[4]) {
    return g.g00 * a[0] * b[0] + g.g01 * (a[0] * b[1] + a[1] * b[0]) + g.g03 * (a[0] * b[3] + a[3] * b[0]) +
           g.g11 * a[1] * b[1] + g.g13 * (a[1] * b[3] + a[3] * b[1]) + g.g22 * a[2] * b[2] + g.g33 * a[3] * b[3];
}

__device__ __forceinline__ void normalize_vec(const MetricCov &g, double v[4]) {
    const double inorm = 1.0 / sqrt(fabs(dot_sparse(g, v, v)));
#pragma unroll
    for (int i = 0; i < 4; ++i)
        v[i] *= inorm;
}

__device__ __forceinline__ void project_out(const MetricCov &g, double a[4], const double b[4]) {
    const double fac = dot_sparse(g, a, b) / dot_sparse(g, b, b);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        a[i] -= b[i] * fac;
}

/* Gram-Schmidt tetrad: e0 = u, e1 ~ trial (d/dx1 if trial is null), e2 ~ d/dx2, e3 ~ d/dx3
 * (reference make_tetrad, tetrads.cpp:68-124).  e_cov[0] carries the extra minus sign of :121-123. */
__device__ __forceinline__ void make_tetrad(const MetricCov &g, const double u_con[4], const double trial_in[4],
                                            double e_con[4][4], double e_cov[4][4]) {
    double trial[4] = {trial_in[0], trial_in[1], trial_in[2], trial_in[3]};
#pragma unroll
    for (int i = 0; i < 4; ++i)
        e_con[0][i] = u_con[i];
    normalize_vec(g, e_con[0]);
    if (dot_sparse(g, trial, trial) < 1.0e-30) {
        trial[0] = 0.0;
        trial[1] = 1.0;
        trial[2] = 0.0;
        trial[3] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        e_con[1][i] = trial[i];
        e_con[2][i] = (i == 2) ? 1.0 : 0.0;
        e_con[3][i] = (i == 3) ? 1.0 : 0.0;
    }
    project_out(g, e_con[1], e_con[0]);
    normalize_vec(g, e_con[1]);
    project_out(g, e_con[2], e_con[0]);
    project_out(g, e_con[2], e_con[1]);
    normalize_vec(g, e_con[2]);
    project_out(g, e_con[3], e_con[0]);
    project_out(g, e_con[3], e_con[1]);
    project_out(g, e_con[3], e_con[2]);
    normalize_vec(g, e_con[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        lower_sparse(g, e_con[i], e_cov[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        e_cov[0][i] = -e_cov[0][i];
}

__device__ __forceinline__ void coordinate_to_tetrad(const double e_cov[4][4], const double k[4], double kt[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        kt[i] = e_cov[i][0] * k[0] + e_cov[i][1] * k[1] + e_cov[i][2] * k[2] + e_cov[i][3] * k[3];
}

__device__ __forceinline__ void tetrad_to_coordinate(const double e[4][4], const double kt[4], double k[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        k[i] = e[0][i] * kt[0] + e[1][i] * kt[1] + e[2][i] * kt[2] + e[3][i] * kt[3];
}

/* ---- samplers ---------------------------------------------------------------------------------------- */
__device__ __forceinline__ void sample_rand_dir(const GmParams &P, Rng &r, double &x, double &y, double &z) {
    z = rng_uniform(P, r) * 2.0 - 1.0;
    const double u = rng_uniform(P, r);
    double s, c;
    sincospi(2.0 * u, &s, &c);
    const double t = sqrt(1.0 - z * z);
    x = t * c;
    y = t * s;
}

/* reference sample_y_distr, proba.cpp:123-166 */
__device__ __forceinline__ double sample_y(const GmParams &P, Rng &r, double theta_e) {
    const double sq = sqrt(0.5 * theta_e);
    double pi_3 = 0.44311346272637900682 /* sqrt(pi)/4 */;
    double pi_4 = sq / 2.0;
    double pi_5 = 3.0 * 1.77245385090551602730 * theta_e / 8.0;
    double pi_6 = theta_e * sq;
    const double is = 1.0 / (pi_3 + pi_4 + pi_5 + pi_6);
    pi_3 *= is;
    pi_4 *= is;
    pi_5 *= is;
    double y, x2, prob;
    do {
        const double x1 = rng_uniform(P, r);
        int dof;
        if (x1 < pi_3)
            dof = 3;
        else if (x1 < pi_3 + pi_4)
            dof = 4;
        else if (x1 < pi_3 + pi_4 + pi_5)
            dof = 5;
        else
            dof = 6;
        const double x = rng_chi_sq(P, r, dof);
        y = sqrt(x / 2.0);
        x2 = rng_uniform(P, r);
        prob = sqrt(1.0 + 0.5 * theta_e * y * y) / (1.0 + y * sq);
    } while (x2 >= prob);
    return y;
}

/* reference sample_mu_distr, proba.cpp:168-172 */
__device__ __forceinline__ double sample_mu(const GmParams &P, Rng &r, double beta_e) {
    const double x1 = rng_uniform(P, r);
    const double det = 1.0 + 2.0 * beta_e + beta_e * beta_e - 4.0 * beta_e * x1;
    return (1.0 - sqrt(det)) / beta_e;
}

/* orthonormal frame (v0 given, v1 from a random direction, v2 = v0 x v1) and a direction at polar cosine
 * c_th / azimuth phi in it; shared by the electron and the photon sampling (proba.cpp:67-106,
 * harm_model.cpp:1166-1204) */
__device__ __forceinline__ void direction_about(const GmParams &P, Rng &r, double v0x, double v0y, double v0z,
                                                double c_th, double s_th, double &d1, double &d2, double &d3) {
    double n0x, n0y, n0z;
    sample_rand_dir(P, r, n0x, n0y, n0z);
    const double n0dotv0 = v0x * n0x + v0y * n0y + v0z * n0z;
    double v1x = n0x - n0dotv0 * v0x, v1y = n0y - n0dotv0 * v0y, v1z = n0z - n0dotv0 * v0z;
    const double iv1 = 1.0 / sqrt(v1x * v1x + v1y * v1y + v1z * v1z);
    v1x *= iv1;
    v1y *= iv1;
    v1z *= iv1;
    const double v2x = v0y * v1z - v0z * v1y, v2y = v0z * v1x - v0x * v1z, v2z = v0x * v1y - v0y * v1x;
    double s_phi, c_phi;
    sincospi(2.0 * rng_uniform(P, r), &s_phi, &c_phi);
    d1 = c_th * v0x + s_th * (c_phi * v1x + s_phi * v2x);
    d2 = c_th * v0y + s_th * (c_phi * v1y + s_phi * v2y);
    d3 = c_th * v0z + s_th * (c_phi * v1z + s_phi * v2z);
}

/* reference sample_electron_distr_p, proba.cpp:30-112 */
__device__ __forceinline__ void sample_electron(const GmParams &P, Rng &r, const double k[4], double theta_e,
                                                double p[4]) {
    double x1, sigma_kn, gamma_e, beta_e, mu;
    do {
        const double y = sample_y(P, r, theta_e);
        gamma_e = y * y * theta_e + 1.0;
        beta_e = sqrt(1.0 - 1.0 / (gamma_e * gamma_e));
        mu = sample_mu(P, r, beta_e);
        mu = fmin(fmax(mu, -1.0), 1.0);
        const double k_ = gamma_e * (1.0 - beta_e * mu) * k[0];
        if (k_ < 1.0e-3) {
            sigma_kn = 1.0 - 2.0 * k_;
        } else {
            const double t = 1.0 + 2.0 * k_;
            sigma_kn = (3.0 / (4.0 * k_ * k_)) *
                       (2.0 + k_ * k_ * (1.0 + k_) / (t * t) + (k_ * k_ - 2.0 * k_ - 2.0) / (2.0 * k_) * log(t));
        }
        x1 = rng_uniform(P, r);
    } while (x1 >= sigma_kn);
    const double iv0 = 1.0 / sqrt(k[1] * k[1] + k[2] * k[2] + k[3] * k[3]);
    double d1, d2, d3;
    direction_about(P, r, k[1] * iv0, k[2] * iv0, k[3] * iv0, mu, sqrt(1. - mu * mu), d1, d2, d3);
    p[0] = gamma_e;
    p[1] = gamma_e * beta_e * d1;
    p[2] = gamma_e * beta_e * d2;
    p[3] = gamma_e * beta_e * d3;
}

/* reference sample_klein_nishina + klein_nishina, proba.cpp:174-189,212-215 */
__device__ __forceinline__ double sample_klein_nishina(const GmParams &P, Rng &r, double k0) {
    const double k0pmin = k0 / (1.0 + 2.0 * k0);
    const double env = 2.0 * (1.0 + 2.0 * k0 + 2.0 * k0 * k0) / (k0 * k0 * (1.0 + 2.0 * k0));
    const double ik0 = 1.0 / k0;
    double x1, k0p, kn;
    do {
        k0p = k0pmin + (k0 - k0pmin) * rng_uniform(P, r);
        x1 = env * rng_uniform(P, r);
        const double ik0p = 1.0 / k0p;
        const double ch = 1.0 + ik0 - ik0p;
        kn = (k0 * ik0p + k0p * ik0 - 1.0 + ch * ch) * (ik0 * ik0);
    } while (x1 >= kn);
    return k0p;
}

/* reference sample_thomson, proba.cpp:191-200 */
__device__ __forceinline__ double sample_thomson(const GmParams &P, Rng &r) {
    double x1, x2;
    do {
        x1 = 2.0 * rng_uniform(P, r) - 1.0;
        x2 = (3.0 / 4.0) * rng_uniform(P, r);
    } while (x2 >= (3.0 / 8.0) * (1.0 + x1 * x1));
    return x1;
}

/* general Lorentz boost of v into the frame moving with 4-velocity u (reference boost, :1658-1671) */
__device__ __forceinline__ void boost(const double v[4], const double u[4], double vp[4]) {
    const double g = u[0];
    const double v_ = sqrt(fabs(1.0 - 1.0 / (g * g)));
    const double in = 1.0 / (g * v_ + kEps);
    const double n1 = u[1] * in, n2 = u[2] * in, n3 = u[3] * in;
    const double gm1 = g - 1.0;
    vp[0] = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
    vp[1] = -u[1] * v[0] + (1.0 + n1 * n1 * gm1) * v[1] + n1 * n2 * gm1 * v[2] + n1 * n3 * gm1 * v[3];
    vp[2] = -u[2] * v[0] + n2 * n1 * gm1 * v[1] + (1.0 + n2 * n2 * gm1) * v[2] + n2 * n3 * gm1 * v[3];
    vp[3] = -u[3] * v[0] + n3 * n1 * gm1 * v[1] + n3 * n2 * gm1 * v[2] + (1.0 + n3 * n3 * gm1) * v[3];
}

/* reference sample_scattered_photon, harm_model.cpp:1147-1215 (p is consumed, its spatial part is flipped) */
__device__ __forceinline__ void sample_scattered_photon(const GmParams &P, Rng &r, const double k[4], double p[4],
                                                        double kp[4]) {
    double ke[4];
    boost(k, p, ke);
    double k0p, c_th;
    if (ke[0] > 1.0e-4) {
        k0p = sample_klein_nishina(P, r, ke[0]);
        c_th = 1.0 - 1.0 / k0p + 1.0 / ke[0];
    } else {
        k0p = ke[0];
        c_th = sample_thomson(P, r);
    }
    const double s_th = sqrt(fabs(1.0 - c_th * c_th));
    const double ike0 = 1.0 / ke[0];
    double d1, d2, d3;
    direction_about(P, r, ke[1] * ike0, ke[2] * ike0, ke[3] * ike0, c_th, s_th, d1, d2, d3);
    p[1] = -p[1];
    p[2] = -p[2];
    p[3] = -p[3];
    const double kpe[4] = {k0p, k0p * d1, k0p * d2, k0p * d3};
    boost(kpe, p, kp);
}

/* what a scattering produces for the new (scattered) superphoton */
struct ScatterChild {
    double k[4];
    double e, l;
};

/* reference scatter_super_photon, harm_model.cpp:1071-1145.
 * Returns true if a valid child was produced.  `w` is the scattering photon's weight: set to 0 when its
 * wave-vector is off the light cone (:1076-1081); the caller then drops it (:1018-1021).
 * Children the reference would leave uninitialised (:1106-1108) or zero-weighted (:1118-1121) are dropped. */
__device__ __forceinline__ bool scatter_super_photon(const GmParams &P, Rng &r, double k[4], double &w,
                                                     const Fluid &f, const MetricCov &g, ScatterChild &ch) {
    if (k[0] > 1.0e5 || k[0] < 0.0 || isnan(k[0]) || isnan(k[1]) || isnan(k[3])) {
        k[0] = fabs(k[0]);
        w = 0.0;
        return false;
    }
    double b_hat[4];
    if (f.b > 0.0) {
        const double ib = 1.0 / (f.b / P.b_unit);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            b_hat[i] = f.b_con[i] * ib;
    } else {
        b_hat[0] = 0.0;
        b_hat[1] = 1.0;
        b_hat[2] = 0.0;
        b_hat[3] = 0.0;
    }
    double e_con[4][4], e_cov[4][4];
    make_tetrad(g, f.u_con, b_hat, e_con, e_cov);
    double kt[4];
    coordinate_to_tetrad(e_cov, k, kt);
    if (kt[0] > 1.0e5 || kt[0] < 0.0 || isnan(kt[1]))
        return false;
    double p[4], ktp[4];
    sample_electron(P, r, kt, f.theta_e, p);
    sample_scattered_photon(P, r, kt, p, ktp);
    tetrad_to_coordinate(e_con, ktp, ch.k);
    if (isnan(ch.k[1]))
        return false;
    ktp[0] = -ktp[0];
    /* only components 0 and 3 of the covariant vector are needed: e = -k_t, l = k_phi */
    ch.e = -(e_cov[0][0] * ktp[0] + e_cov[1][0] * ktp[1] + e_cov[2][0] * ktp[2] + e_cov[3][0] * ktp[3]);
    ch.l = e_cov[0][3] * ktp[0] + e_cov[1][3] * ktp[1] + e_cov[2][3] * ktp[2] + e_cov[3][3] * ktp[3];
    return true;
}

} /* namespace gm */
