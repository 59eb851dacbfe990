/*
 * oracle/grmonty_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See grmonty_oracle.h.
 *
 * Scalar, single-threaded, straightforward: clarity over speed.  Arithmetic follows the
 * reference expression by expression (build with -ffp-contract=off) so that function-level
 * results agree with the reference CPU build to the last few ulps.
 *
 * "ref:" comments give the reference file:line (under /root/reference/cuda_grmonty/) restated.
 */
#define _DEFAULT_SOURCE
#define _USE_MATH_DEFINES
#include "grmonty_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants, ref: consts.hpp:14-171 (values must match exactly) ------------------------- */
#define C_EPS 1.0e-40
#define C_NU_MIN 1.0e9
#define C_NU_MAX 1.0e16
#define C_THETA_E_MIN 0.3
#define C_TP_OVER_TE 3.0
#define C_WEIGHT_MIN 1.0e31
#define C_ROULETTE 1.0e4
#define C_R_MAX 100.0
#define C_STEP_EPS 0.04
#define C_E_TOL 1.0e-3
#define C_MAX_ITER 2
#define C_MAX_N_STEP 1280000
#define C_EE 4.80320680e-10
#define C_CL 2.99792458e10
#define C_ME 9.1093826e-28
#define C_MP 1.67262171e-24
#define C_HPL 6.6260693e-27
#define C_SIGMA_THOMSON 0.665245873e-24
#define C_BTHSQ_MIN 1.0e-4
#define C_BTHSQ_MAX 1.0e8
#define C_HC_MIN_W 1.0e-12
#define C_HC_MAX_W 1.0e6
#define C_HC_MIN_T 1.0e-4
#define C_HC_MAX_T 1.0e4
#define C_HC_MAX_GAMMA 12.0
#define C_HC_D_MU_E 0.05
#define C_HC_D_GAMMA_E 0.05
#define C_JNU_MIN_K 0.002
#define C_JNU_MAX_K 1.0e7
#define C_JNU_MAX_T 1.0e2
#define C_JNU_CST 1.88774862536
#define C_SPEC_D_L_E 0.25

static double c_l_nu_min(void) { return log(C_NU_MIN); }
static double c_n_l_n(void) { return log(C_NU_MAX) - log(C_NU_MIN); }
static double c_d_l_nu(void) { return (log(C_NU_MAX) - log(C_NU_MIN)) / ORC_N_E_SAMP; }
static double c_x1_max(void) { return log(C_R_MAX); }
static double c_l_b_min(void) { return log(C_BTHSQ_MIN); }
static double c_d_l_b(void) { return log(C_BTHSQ_MAX / C_BTHSQ_MIN) / ORC_NINT; }
static double c_hc_l_min_w(void) { return log10(C_HC_MIN_W); }
static double c_hc_l_min_t(void) { return log10(C_HC_MIN_T); }
static double c_hc_d_l_w(void) { return log10(C_HC_MAX_W / C_HC_MIN_W) / ORC_HC_NW; }
static double c_hc_d_l_t(void) { return log10(C_HC_MAX_T / C_HC_MIN_T) / ORC_HC_NT; }
static double c_jnu_l_min_k(void) { return log(C_JNU_MIN_K); }
static double c_jnu_d_l_k(void) { return log(C_JNU_MAX_K / C_JNU_MIN_K) / ORC_N_E_SAMP; }
static double c_jnu_l_min_t(void) { return log(C_THETA_E_MIN); }
static double c_jnu_d_l_t(void) { return log(C_JNU_MAX_T / C_THETA_E_MIN) / ORC_N_E_SAMP; }
static double c_jnu_k_fac(void) { return 9 * M_PI * C_ME * C_CL / C_EE; }
static double c_spec_l_e_0(void) { return log(1.0e-12); }

/* ============================================================================================
 * RNG: Philox4x32-10 (Salmon et al. 2011), replaces ref: monty_rand.cpp:19-31.
 * stream = (id[0..2]) in counter words 1..3, draw index in counter word 0, key = 64-bit seed.
 * ============================================================================================ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

void orc_rng_primary(orc_rng *r, uint64_t primary_index) {
    r->id[0] = (uint32_t)primary_index;
    r->id[1] = (uint32_t)(primary_index >> 32);
    r->id[2] = 0u;
    r->ctr = 0u;
}

void orc_rng_zone(orc_rng *r, uint64_t zone_index) {
    r->id[0] = (uint32_t)zone_index;
    r->id[1] = (uint32_t)(zone_index >> 32);
    r->id[2] = 0x40000000u;
    r->ctr = 0u;
}

static void rng_block(const orc_model *m, orc_rng *r, uint32_t out[4]) {
    uint32_t ctr[4] = {r->ctr, r->id[0], r->id[1], r->id[2]};
    uint32_t key[2] = {(uint32_t)m->seed, (uint32_t)(m->seed >> 32)};
    orc_philox4x32_10(ctr, key, out);
    r->ctr += 1u;
}

/* uniform in the open interval (0,1): (53-bit integer + 0.5) * 2^-53 */
double orc_uniform(const orc_model *m, orc_rng *r) {
    uint32_t o[4];
    rng_block(m, r, o);
    uint64_t bits = ((uint64_t)o[1] << 32) | o[0];
    return ((double)(bits >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

/* a scattered photon gets a fresh 95-bit identity drawn from its parent's stream */
void orc_rng_child(const orc_model *m, orc_rng *parent, orc_rng *child) {
    uint32_t o[4];
    rng_block(m, parent, o);
    child->id[0] = o[0];
    child->id[1] = o[1];
    child->id[2] = o[2] | 0x80000000u;
    child->ctr = 0u;
}

/* chi-square with 3..6 degrees of freedom from uniforms (exact):
 *   chi2_4 = -2 ln(U1 U2), chi2_6 = -2 ln(U1 U2 U3), chi2_3 = -2 ln U1 + N^2, chi2_5 = -2 ln(U1 U2) + N^2,
 *   N^2 = -2 ln(Ua) cos^2(2 pi Ub) (Box-Muller).  Replaces ref: monty_rand.cpp:28-31. */
double orc_chi_sq(const orc_model *m, orc_rng *r, int dof) {
    double s = 0.0;
    int n_exp = dof / 2;
    for (int i = 0; i < n_exp; ++i) {
        s += -2.0 * log(orc_uniform(m, r));
    }
    if (dof & 1) {
        double ua = orc_uniform(m, r);
        double ub = orc_uniform(m, r);
        double c = cos(2.0 * M_PI * ub);
        s += -2.0 * log(ua) * c * c;
    }
    return s;
}

/* ============================================================================================
 * Geometry
 * ============================================================================================ */
/* ref: harm_model.cpp:1632-1637 */
static void bl_coord(const orc_model *m, const double x[4], double *r, double *th) {
    *r = exp(x[1]) + m->r_0;
    *th = M_PI * x[2] + ((1.0 - m->h_slope) / 2.0) * sin(2.0 * M_PI * x[2]);
}

/* ref: harm_model.cpp:499-530 */
void orc_gcov(const orc_model *m, const double x[4], double g[4][4]) {
    memset(g, 0, sizeof(double) * 16);
    double r, th;
    bl_coord(m, x, &r, &th);
    double sin_theta = fabs(sin(th)) + C_EPS;
    double cos_theta = cos(th);
    double s2 = sin_theta * sin_theta;
    double rho2 = r * r + m->a * m->a * cos_theta * cos_theta;
    double tfac = 1.0;
    double rfac = r - m->r_0;
    double hfac = M_PI + (1.0 - m->h_slope) * M_PI * cos(2.0 * M_PI * x[2]);
    double pfac = 1.0;
    g[0][0] = (-1.0 + 2.0 * r / rho2) * tfac * tfac;
    g[0][1] = (2.0 * r / rho2) * tfac * rfac;
    g[0][3] = (-2.0 * m->a * r * s2 / rho2) * tfac * pfac;
    g[1][0] = g[0][1];
    g[1][1] = (1.0 + 2.0 * r / rho2) * rfac * rfac;
    g[1][3] = (-m->a * s2 * (1.0 + 2.0 * r / rho2)) * rfac * pfac;
    g[2][2] = rho2 * hfac * hfac;
    g[3][0] = g[0][3];
    g[3][1] = g[1][3];
    g[3][3] = s2 * (rho2 + m->a * m->a * s2 * (1.0 + 2.0 * r / rho2)) * pfac * pfac;
}

/* ref: harm_model.cpp:473-497 */
void orc_gcon(const orc_model *m, const double x[4], double g[4][4]) {
    memset(g, 0, sizeof(double) * 16);
    double r, th;
    bl_coord(m, x, &r, &th);
    double sin_theta = fabs(sin(th)) + C_EPS;
    double cos_theta = cos(th);
    double irho2 = 1.0 / (r * r + m->a * m->a * cos_theta * cos_theta);
    double hfac = M_PI + (1.0 - m->h_slope) * M_PI * cos(2.0 * M_PI * x[2]);
    g[0][0] = -1.0 - 2.0 * r * irho2;
    g[0][1] = 2.0 * irho2;
    g[1][0] = g[0][1];
    g[1][1] = irho2 * (r * (r - 2.0) + m->a * m->a) / (r * r);
    g[1][3] = m->a * irho2 / r;
    g[2][2] = irho2 / (hfac * hfac);
    g[3][1] = g[1][3];
    g[3][3] = irho2 / (sin_theta * sin_theta);
}

/* ref: harm_model.cpp:1436-1569.  Closed-form MKS Kerr connection; r = exp(x1) (r_0 ignored, Appendix A.1).
 * Only [i][j][k] with j <= k are filled by the reference; we fill both triangles. */
void orc_get_connection(const orc_model *m, const double x[4], double lc[4][4][4]) {
    double r1 = exp(x[1]);
    double r2 = r1 * r1, r3 = r2 * r1, r4 = r3 * r1;
    double s_x = sin(2.0 * M_PI * x[2]);
    double c_x = cos(2.0 * M_PI * x[2]);
    double th = M_PI * x[2] + 0.5 * (1.0 - m->h_slope) * s_x;
    double dthdx2 = M_PI * (1.0 + (1.0 - m->h_slope) * c_x);
    double d2thdx22 = -2.0 * M_PI * M_PI * (1.0 - m->h_slope) * s_x;
    double dthdx22 = dthdx2 * dthdx2;
    double sth = sin(th), cth = cos(th);
    double sth2 = sth * sth, r1sth2 = r1 * sth2, sth4 = sth2 * sth2;
    double cth2 = cth * cth, cth4 = cth2 * cth2;
    double s2th = 2.0 * sth * cth, c2th = 2.0 * cth2 - 1.0;
    double a = m->a, a2 = a * a, a3 = a2 * a, a4 = a3 * a;
    double a2sth2 = a2 * sth2, a2cth2 = a2 * cth2, a4cth4 = a4 * cth4;
    double rho2 = r2 + a2cth2, rho22 = rho2 * rho2, rho23 = rho22 * rho2;
    double irho2 = 1.0 / rho2, irho22 = irho2 * irho2, irho23 = irho22 * irho2;
    double irho23_dthdx2 = irho23 / dthdx2;
    double fac1 = r2 - a2cth2, fac1_rho23 = fac1 * irho23;
    double fac2 = a2 + 2.0 * r2 + a2 * c2th;
    double fac3 = a2 + r1 * (-2.0 + r1);

    lc[0][0][0] = 2.0 * r1 * fac1_rho23;
    lc[0][0][1] = r1 * (2.0 * r1 + rho2) * fac1_rho23;
    lc[0][0][2] = -a2 * r1 * s2th * dthdx2 * irho22;
    lc[0][0][3] = -2.0 * a * r1sth2 * fac1_rho23;
    lc[0][1][1] = 2.0 * r2 * (r4 + r1 * fac1 - a4cth4) * irho23;
    lc[0][1][2] = -a2 * r2 * s2th * dthdx2 * irho22;
    lc[0][1][3] = a * r1 * (-r1 * (r3 + 2.0 * fac1) + a4cth4) * sth2 * irho23;
    lc[0][2][2] = -2.0 * r2 * dthdx22 * irho2;
    lc[0][2][3] = a3 * r1sth2 * s2th * dthdx2 * irho22;
    lc[0][3][3] = 2.0 * r1sth2 * (-r1 * rho22 + a2sth2 * fac1) * irho23;

    lc[1][0][0] = fac3 * fac1 / (r1 * rho23);
    lc[1][0][1] = fac1 * (-2.0 * r1 + a2sth2) * irho23;
    lc[1][0][2] = 0.0;
    lc[1][0][3] = -a * sth2 * fac3 * fac1 / (r1 * rho23);
    lc[1][1][1] = (r4 * (-2.0 + r1) * (1.0 + r1) + a2 * (a2 * r1 * (1.0 + 3.0 * r1) * cth4 + a4cth4 * cth2 +
                                                          r3 * sth2 + r1 * cth2 * (2.0 * r1 + 3.0 * r3 - a2sth2))) *
                  irho23;
    lc[1][1][2] = -a2 * dthdx2 * s2th / fac2;
    lc[1][1][3] = a * sth2 *
                  (a4 * r1 * cth4 + r2 * (2.0 * r1 + r3 - a2sth2) + a2cth2 * (2.0 * r1 * (-1.0 + r2) + a2sth2)) *
                  irho23;
    lc[1][2][2] = -fac3 * dthdx22 * irho2;
    lc[1][2][3] = 0.0;
    lc[1][3][3] = -fac3 * sth2 * (r1 * rho22 - a2 * fac1 * sth2) / (r1 * rho23);

    lc[2][0][0] = -a2 * r1 * s2th * irho23_dthdx2;
    lc[2][0][1] = r1 * lc[2][0][0];
    lc[2][0][2] = 0.0;
    lc[2][0][3] = a * r1 * (a2 + r2) * s2th * irho23_dthdx2;
    lc[2][1][1] = r2 * lc[2][0][0];
    lc[2][1][2] = r2 * irho2;
    lc[2][1][3] =
        (a * r1 * cth * sth * (r3 * (2.0 + r1) + a2 * (2.0 * r1 * (1.0 + r1) * cth2 + a2 * cth4 + 2.0 * r1sth2))) *
        irho23_dthdx2;
    lc[2][2][2] = -a2 * cth * sth * dthdx2 * irho2 + d2thdx22 / dthdx2;
    lc[2][2][3] = 0.0;
    lc[2][3][3] =
        -cth * sth * (rho23 + a2sth2 * rho2 * (r1 * (4.0 + r1) + a2cth2) + 2.0 * r1 * a4 * sth4) * irho23_dthdx2;

    lc[3][0][0] = a * fac1_rho23;
    lc[3][0][1] = r1 * lc[3][0][0];
    lc[3][0][2] = -2.0 * a * r1 * cth * dthdx2 / (sth * rho22);
    lc[3][0][3] = -a2sth2 * fac1_rho23;
    lc[3][1][1] = a * r2 * fac1_rho23;
    lc[3][1][2] = -2 * a * r1 * (a2 + 2.0 * r1 * (2.0 + r1) + a2 * c2th) * cth * dthdx2 / (sth * fac2 * fac2);
    lc[3][1][3] = r1 * (r1 * rho22 - a2sth2 * fac1) * irho23;
    lc[3][2][2] = -a * r1 * dthdx22 * irho2;
    lc[3][2][3] = dthdx2 * (0.25 * fac2 * fac2 * cth / sth + a2 * r1 * s2th) * irho22;
    lc[3][3][3] = (-a * r1sth2 * rho22 + a3 * sth4 * fac1) * irho23;

    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            for (int k = 0; k < j; ++k)
                lc[i][j][k] = lc[i][k][j];
}

/* dk^i/dlam = -Gamma^i_{jk} k^j k^k using the upper triangle; ref: harm_model.cpp:1578-1586, 1255-1262 */
static double geodesic_rhs(double lc[4][4][4], int i, const double k[4]) {
    double d = -2.0 * (k[0] * (lc[i][0][1] * k[1] + lc[i][0][2] * k[2] + lc[i][0][3] * k[3]) +
                       k[1] * (lc[i][1][2] * k[2] + lc[i][1][3] * k[3]) + lc[i][2][3] * k[2] * k[3]);
    d -= (lc[i][0][0] * k[0] * k[0] + lc[i][1][1] * k[1] * k[1] + lc[i][2][2] * k[2] * k[2] +
          lc[i][3][3] * k[3] * k[3]);
    return d;
}

/* ref: harm_model.cpp:1571-1587 */
void orc_init_dkdlam(const orc_model *m, const double x[4], const double k[4], double dk[4]) {
    double lc[4][4][4];
    orc_get_connection(m, x, lc);
    for (int i = 0; i < 4; ++i)
        dk[i] = geodesic_rhs(lc, i, k);
}

/* ref: harm_model.cpp:1620-1630 */
double orc_step_size(const orc_model *m, const double x[4], const double k[4]) {
    double dl_x_1 = C_STEP_EPS * x[1] / (fabs(k[1]) + C_EPS);
    double dl_x_2 = C_STEP_EPS * fmin(x[2], m->x_stop2 - x[2]) / (fabs(k[2]) + C_EPS);
    double dl_x_3 = C_STEP_EPS / (fabs(k[3]) + C_EPS);
    double i1 = 1.0 / (fabs(dl_x_1) + C_EPS);
    double i2 = 1.0 / (fabs(dl_x_2) + C_EPS);
    double i3 = 1.0 / (fabs(dl_x_3) + C_EPS);
    return 1.0 / (i1 + i2 + i3);
}

/* ref: harm_model.cpp:1217-1289.  Half-kick / drift / fixed-point kick with recursive halving. */
void orc_push_photon(orc_model *m, orc_photon *ph, double dl, int n) {
    if (ph->x[1] < m->x_start1)
        return;
    double x0[4], k0[4], dk0[4];
    memcpy(x0, ph->x, sizeof(x0));
    memcpy(k0, ph->k, sizeof(k0));
    memcpy(dk0, ph->dkdlam, sizeof(dk0));
    m->n_push_attempts++;

    double dl_2 = 0.5 * dl;
    double k[4];
    for (int i = 0; i < 4; ++i) {
        double dk = ph->dkdlam[i] * dl_2;
        ph->k[i] += dk;
        k[i] = ph->k[i] + dk;
        ph->x[i] += ph->k[i] * dl;
    }
    double lc[4][4][4];
    orc_get_connection(m, ph->x, lc);
    double err;
    int iter = 0;
    do {
        ++iter;
        double kc[4];
        memcpy(kc, k, sizeof(kc));
        err = 0.0;
        for (int i = 0; i < 4; ++i) {
            ph->dkdlam[i] = geodesic_rhs(lc, i, kc);
            k[i] = ph->k[i] + dl_2 * ph->dkdlam[i];
            err += fabs((kc[i] - k[i]) / (k[i] + C_EPS));
        }
    } while (err > C_E_TOL && iter < C_MAX_ITER);
    memcpy(ph->k, k, sizeof(k));

    double g[4][4];
    orc_gcov(m, ph->x, g);
    double e_1 = -(ph->k[0] * g[0][0] + ph->k[1] * g[0][1] + ph->k[2] * g[0][2] + ph->k[3] * g[0][3]);
    double err_e = fabs((e_1 - ph->e_0_s) / ph->e_0_s);
    if (n < 7 && (err_e > 1.0e-4 || err > C_E_TOL || isnan(err) || isinf(err))) {
        memcpy(ph->x, x0, sizeof(x0));
        memcpy(ph->k, k0, sizeof(k0));
        memcpy(ph->dkdlam, dk0, sizeof(dk0));
        orc_push_photon(m, ph, 0.5 * dl, n + 1);
        orc_push_photon(m, ph, 0.5 * dl, n + 1);
        e_1 = ph->e_0_s;
    }
    ph->e_0_s = e_1;
}

/* ref: harm_model.cpp:1589-1616 (uniform() <= 1/roulette survives) */
int orc_stop_criterion(const orc_model *m, orc_photon *ph) {
    if (ph->x[1] < m->x1_min)
        return 1;
    if (ph->x[1] > c_x1_max()) {
        if (ph->w < C_WEIGHT_MIN) {
            if (orc_uniform(m, &ph->rng) <= 1.0 / C_ROULETTE)
                ph->w *= C_ROULETTE;
            else
                ph->w = 0.0;
        }
        return 1;
    }
    if (ph->w < C_WEIGHT_MIN) {
        if (orc_uniform(m, &ph->rng) <= 1.0 / C_ROULETTE) {
            ph->w *= C_ROULETTE;
        } else {
            ph->w = 0.0;
            return 1;
        }
    }
    return 0;
}

/* ============================================================================================
 * Fluid
 * ============================================================================================ */
/* ref: tetrads.cpp:132-160 */
static void lower(const double v_con[4], double g[4][4], double v_cov[4]) {
    for (int i = 0; i < 4; ++i)
        v_cov[i] = g[i][0] * v_con[0] + g[i][1] * v_con[1] + g[i][2] * v_con[2] + g[i][3] * v_con[3];
}

/* ref: harm_model.cpp:1406-1434 */
static void x_to_ij(const orc_model *m, const double x[4], int *pi, int *pj, double *pdi, double *pdj) {
    int i = (int)((x[1] - m->x_start1) / m->dx1 - 0.5 + 1000) - 1000;
    int j = (int)((x[2] - m->x_start2) / m->dx2 - 0.5 + 1000) - 1000;
    double del_i, del_j;
    if (i < 0) {
        i = 0;
        del_i = 0.0;
    } else if (i > m->n0 - 2) {
        i = m->n0 - 2;
        del_i = 1.0;
    } else {
        del_i = (x[1] - ((i + 0.5) * m->dx1 + m->x_start1)) / m->dx1;
    }
    if (j < 0) {
        j = 0;
        del_j = 0.0;
    } else if (j > m->n1 - 2) {
        j = m->n1 - 2;
        del_j = 1.0;
    } else {
        del_j = (x[2] - ((j + 0.5) * m->dx2 + m->x_start2)) / m->dx2;
    }
    *pi = i;
    *pj = j;
    *pdi = del_i;
    *pdj = del_j;
}

/* ref: harm_model.cpp:1646-1656 */
static double interp_scalar(const orc_model *m, const double *var, int i, int j, const double coeff[4]) {
    int n1 = m->n1;
    return var[i * n1 + j] * coeff[0] + var[i * n1 + j + 1] * coeff[1] + var[(i + 1) * n1 + j] * coeff[2] +
           var[(i + 1) * n1 + j + 1] * coeff[3];
}

/* shared tail of get_fluid_params / get_fluid_zone: primitives -> u^mu, b^mu, |B|;
 * ref: harm_model.cpp:638-668 and :560-590 */
static void prims_to_fluid(const orc_model *m, const double v_con[4], const double bp[4], double gcov[4][4],
                           const double gcon0[4], orc_fluid *f) {
    double v_dot_v = 0.0;
    for (int i = 1; i < 4; ++i)
        for (int j = 1; j < 4; ++j)
            v_dot_v += gcov[i][j] * v_con[i] * v_con[j];
    double v_fac = sqrt(-1.0 / gcon0[0] * (1.0 + fabs(v_dot_v)));
    f->u_con[0] = -v_fac * gcon0[0];
    for (int i = 1; i < 4; ++i)
        f->u_con[i] = v_con[i] - v_fac * gcon0[i];
    lower(f->u_con, gcov, f->u_cov);
    double u_dot_bp = 0.0;
    for (int i = 1; i < 4; ++i)
        u_dot_bp += f->u_cov[i] * bp[i];
    f->b_con[0] = u_dot_bp;
    for (int i = 1; i < 4; ++i)
        f->b_con[i] = (bp[i] + f->u_con[i] * u_dot_bp) / f->u_con[0];
    lower(f->b_con, gcov, f->b_cov);
    f->b = sqrt(f->b_con[0] * f->b_cov[0] + f->b_con[1] * f->b_cov[1] + f->b_con[2] * f->b_cov[2] +
                f->b_con[3] * f->b_cov[3]) *
           m->b_unit;
}

/* ref: harm_model.cpp:595-671.  Outside the grid only n_e = 0 is defined; we zero the rest (Appendix A.15). */
void orc_get_fluid_params(const orc_model *m, const double x[4], double gcov[4][4], orc_fluid *f) {
    memset(f, 0, sizeof(*f));
    if (x[1] < m->x_start1 || x[1] > m->x_stop1 || x[2] < m->x_start2 || x[2] > m->x_stop2) {
        f->n_e = 0.0;
        return;
    }
    int i, j;
    double del_i, del_j;
    x_to_ij(m, x, &i, &j, &del_i, &del_j);
    double coeff[4] = {(1.0 - del_i) * (1.0 - del_j), (1.0 - del_i) * del_j, del_i * (1.0 - del_j), del_i * del_j};
    double rho = interp_scalar(m, m->k_rho, i, j, coeff);
    double uu = interp_scalar(m, m->u, i, j, coeff);
    f->n_e = rho * m->n_e_unit;
    f->theta_e = uu / rho * m->theta_e_unit;
    double bp[4] = {0.0, interp_scalar(m, m->b_1, i, j, coeff), interp_scalar(m, m->b_2, i, j, coeff),
                    interp_scalar(m, m->b_3, i, j, coeff)};
    double v_con[4] = {0.0, interp_scalar(m, m->u_1, i, j, coeff), interp_scalar(m, m->u_2, i, j, coeff),
                       interp_scalar(m, m->u_3, i, j, coeff)};
    double gcon[4][4];
    orc_gcon(m, x, gcon);
    prims_to_fluid(m, v_con, bp, gcov, gcon[0], f);
}

/* ref: harm_model.cpp:1639-1644 */
static void get_coord(const orc_model *m, int i, int j, double x[4]) {
    x[0] = 0.0;
    x[1] = m->x_start1 + (i + 0.5) * m->dx1;
    x[2] = m->x_start2 + (j + 0.5) * m->dx2;
    x[3] = 0.0;
}

/* ref: harm_model.cpp:538-593 (zone-centre state; u_cov/b_cov are filled as a by-product) */
void orc_get_fluid_zone(const orc_model *m, int i, int j, orc_fluid *f) {
    memset(f, 0, sizeof(*f));
    double x[4], gcov[4][4], gcon[4][4];
    get_coord(m, i, j, x);
    orc_gcov(m, x, gcov);
    orc_gcon(m, x, gcon);
    int z = i * m->n1 + j;
    double v_con[4] = {0.0, m->u_1[z], m->u_2[z], m->u_3[z]};
    double bp[4] = {0.0, m->b_1[z], m->b_2[z], m->b_3[z]};
    f->n_e = m->k_rho[z] * m->n_e_unit;
    f->theta_e = (m->u[z] / f->n_e) * m->n_e_unit * m->theta_e_unit;
    prims_to_fluid(m, v_con, bp, gcov, gcon[0], f);
}

/* ============================================================================================
 * Radiation
 * ============================================================================================ */
/* ref: jnu_mixed.cpp:150-158 (+ clamp of the upper index, Appendix A.17) */
static double interp_exp_table(const double *tab, double lx, double l_min, double d_l) {
    double d_i = (lx - l_min) / d_l;
    int i = (int)d_i;
    if (i > ORC_N_E_SAMP - 1)
        i = ORC_N_E_SAMP - 1;
    d_i -= i;
    return exp((1.0 - d_i) * tab[i] + d_i * tab[i + 1]);
}

/* ref: jnu_mixed.cpp:102-111 */
double orc_k2_eval(const orc_model *m, double theta_e) {
    if (theta_e < C_THETA_E_MIN)
        return 0.0;
    if (theta_e > C_JNU_MAX_T)
        return 2.0 * theta_e * theta_e;
    return interp_exp_table(m->k2, log(theta_e), c_jnu_l_min_t(), c_jnu_d_l_t());
}

/* ref: jnu_mixed.cpp:113-125 */
double orc_f_eval(const orc_model *m, double theta_e, double b_mag, double nu) {
    double k = c_jnu_k_fac() * nu / (b_mag * theta_e * theta_e);
    if (k > C_JNU_MAX_K)
        return 0.0;
    if (k < C_JNU_MIN_K) {
        double x = pow(k, 1.0 / 3.0);
        return x * (37.67503800178 + 2.240274341836 * x);
    }
    return interp_exp_table(m->f, log(k), c_jnu_l_min_k(), c_jnu_d_l_k());
}

/* ref: jnu_mixed.cpp:75-100 */
double orc_synch(const orc_model *m, double nu, double n_e, double theta_e, double b, double theta) {
    if (theta_e < C_THETA_E_MIN)
        return 0.0;
    double k2 = orc_k2_eval(m, theta_e);
    double nu_c = C_EE * b / (2.0 * M_PI * C_ME * C_CL);
    double sin_th = sin(theta);
    double nu_s = (2.0 / 9.0) * nu_c * theta_e * theta_e * sin_th;
    if (nu > 1.0e12 * nu_s)
        return 0.0;
    double x = nu / nu_s;
    double xp = pow(x, 1.0 / 3.0);
    double xx = sqrt(x) + C_JNU_CST * sqrt(xp);
    double f = xx * xx;
    return (M_SQRT2 * M_PI * C_EE * C_EE * n_e * nu_s / (3.0 * C_CL * k2)) * f * exp(-xp);
}

/* ref: hotcross.cpp:144-152 */
static double hc_klein_nishina(double w) {
    if (w < 1.0e-3)
        return (1.0 - 2.0 * w);
    return (3.0 / 4.0) * (2.0 / (w * w) + (1.0 / (2.0 * w) - (1.0 + w) / (w * w * w)) * log(1.0 + 2.0 * w) +
                          (1.0 + w) / ((1.0 + 2.0 * w) * (1.0 + 2.0 * w)));
}

/* ref: hotcross.cpp:108-142 with dnd_gamma_e (:154-164) and boostcross (:166-179) inlined.
 * k2f = K2(1/theta_e) exp(1/theta_e) (or sqrt(pi theta_e/2) when theta_e <= 1e-2) is passed in because
 * C99 has no modified Bessel function; it does not depend on the integration variables. */
double orc_hotcross_num(double w, double theta_e, double k2f) {
    if (isnan(w))
        return 0.0;
    if (theta_e < C_HC_MIN_T && w < C_HC_MIN_W)
        return C_SIGMA_THOMSON;
    if (theta_e < C_HC_MIN_T)
        return hc_klein_nishina(w) * C_SIGMA_THOMSON;
    double cross = 0.0;
    for (double mu_e = -1.0 + 0.5 * C_HC_D_MU_E; mu_e < 1.0; mu_e += C_HC_D_MU_E) {
        for (double gamma_e = 1.0 + 0.5 * theta_e * C_HC_D_GAMMA_E; gamma_e < 1.0 + C_HC_MAX_GAMMA * theta_e;
             gamma_e += theta_e * C_HC_D_GAMMA_E) {
            double dnd =
                ((gamma_e * sqrt(gamma_e * gamma_e - 1.) / (theta_e * k2f)) * exp(-(gamma_e - 1.) / theta_e));
            double f = 0.5 * dnd;
            double v = sqrt(gamma_e * gamma_e - 1.0) / gamma_e;
            double we = w * gamma_e * (1.0 - mu_e * v);
            double boostcross = hc_klein_nishina(we) * (1.0 - mu_e * v);
            cross += theta_e * C_HC_D_MU_E * C_HC_D_GAMMA_E * boostcross * f;
        }
    }
    return cross * C_SIGMA_THOMSON;
}

/* ref: hotcross.cpp:81-106.  The out-of-table numeric fall-back (:90-93) needs K2 and is out of the
 * oracle's reach in C99; it returns NaN there so that a test hitting it fails loudly. */
/* Diagnostics (read by tools through ctypes): out-of-table cross-section calls, the longest-lived photon, and how far
 * from the light cone scattered wave-vectors come out (a degenerate scattering tetrad shows up here). */
double orc_diag_max_offcone = 0.0;
uint64_t orc_diag_n_offcone = 0;
int orc_diag_max_n_step = 0;
double orc_diag_longest[10] = {0};
uint64_t orc_n_out_of_table = 0;
double orc_out_of_table_args[2] = {0.0, 0.0};
double orc_hotcross_lkup(const orc_model *m, double w, double theta_e) {
    if (w * theta_e < 1.0e-6)
        return C_SIGMA_THOMSON;
    if (theta_e < C_HC_MIN_T)
        return hc_klein_nishina(w) * C_SIGMA_THOMSON;
    if (w <= C_HC_MIN_W || w >= C_HC_MAX_W || theta_e <= C_HC_MIN_T || theta_e >= C_HC_MAX_T) {
        if (orc_n_out_of_table++ == 0) {
            orc_out_of_table_args[0] = w;
            orc_out_of_table_args[1] = theta_e;
        }
        return NAN;
    }
    const double l_w = log10(w);
    const double l_t = log10(theta_e);
    int i = (int)((l_w - c_hc_l_min_w()) / c_hc_d_l_w());
    int j = (int)((l_t - c_hc_l_min_t()) / c_hc_d_l_t());
    double d_i = (l_w - c_hc_l_min_w()) / c_hc_d_l_w() - i;
    double d_j = (l_t - c_hc_l_min_t()) / c_hc_d_l_t() - j;
    const double *t = m->hotcross;
    const int nt = ORC_HC_NT + 1;
    double l_cross = (1.0 - d_i) * (1.0 - d_j) * t[i * nt + j] + d_i * (1.0 - d_j) * t[(i + 1) * nt + j] +
                     (1.0 - d_i) * d_j * t[i * nt + j + 1] + d_i * d_j * t[(i + 1) * nt + j + 1];
    return pow(10, l_cross);
}

/* ref: radiation.cpp:59-87 */
double orc_bk_angle(const orc_model *m, const double k[4], const double u_cov[4], const double b_cov[4], double b) {
    if (b == 0.0)
        return M_PI / 2.0;
    double k_ = fabs(k[0] * u_cov[0] + k[1] * u_cov[1] + k[2] * u_cov[2] + k[3] * u_cov[3]);
    double mu = (k[0] * b_cov[0] + k[1] * b_cov[1] + k[2] * b_cov[2] + k[3] * b_cov[3]) / (k_ * b / m->b_unit);
    if (mu < -1.0)
        mu = -1.0;
    if (mu > 1.0)
        mu = 1.0;
    return acos(mu);
}

/* ref: radiation.cpp:89-101 */
double orc_fluid_nu(const double k[4], const double u_cov[4]) {
    double energy = -(k[0] * u_cov[0] + k[1] * u_cov[1] + k[2] * u_cov[2] + k[3] * u_cov[3]);
    return energy * C_ME * C_CL * C_CL / C_HPL;
}

/* ref: radiation.cpp:103-107, 142-146 */
double orc_alpha_inv_scatt(const orc_model *m, double nu, double theta_e, double n_e) {
    double e_g = C_HPL * nu / (C_ME * C_CL * C_CL);
    double kappa = orc_hotcross_lkup(m, e_g, theta_e) / C_MP;
    return nu * kappa * n_e * C_MP;
}

/* ref: radiation.cpp:120-128 */
static double b_nu_inv(double nu, double theta_e) {
    double x = C_HPL * nu / (C_ME * C_CL * C_CL * theta_e);
    if (x < 1.0e-3)
        return (2.0 * C_HPL / (C_CL * C_CL)) / (x / 24.0 * (24.0 + x * (12.0 + x * (4.0 + x))));
    return (2.0 * C_HPL / (C_CL * C_CL)) / (exp(x) - 1.0);
}

/* ref: radiation.cpp:109-118, 130-140 */
double orc_alpha_inv_abs(const orc_model *m, double nu, double theta_e, double n_e, double b, double theta) {
    double j = orc_synch(m, nu, n_e, theta_e, b, theta) / (nu * nu);
    double b_nu = b_nu_inv(nu, theta_e);
    return j / (b_nu + 1.0e-100);
}

/* ref: harm_model.cpp:1391-1404.  Reads the bias_* statistics (frozen or live, see header). */
double orc_bias_func(const orc_model *m, double t_e, double w) {
    double max = 0.5 * w / C_WEIGHT_MIN;
    double avg_num_scatt = m->bias_n_scatt / (1.0 * m->bias_n_recorded + 1.0);
    double bias = 100.0 * t_e * t_e / (m->bias_norm * m->bias_max_tau_scatt * (avg_num_scatt + 2.0));
    if (bias < C_TP_OVER_TE)
        bias = C_TP_OVER_TE;
    if (bias > max)
        bias = max;
    return bias / C_TP_OVER_TE;
}

/* ============================================================================================
 * Tetrads, ref: tetrads.cpp:46-194
 * ============================================================================================ */
static double dot_g(const double a[4], const double b[4], double g[4][4]) {
    double s = 0.0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            s += a[i] * b[j] * g[i][j];
    return s;
}
/* ref: tetrads.cpp:162-176 */
static void normalize(double v[4], double g[4][4]) {
    double norm = sqrt(fabs(dot_g(v, v, g)));
    for (int i = 0; i < 4; ++i)
        v[i] /= norm;
}
/* ref: tetrads.cpp:178-194 */
static void project_out(double a[4], const double b[4], double g[4][4]) {
    double b_sq = dot_g(b, b, g);
    double a_dot_b = dot_g(a, b, g);
    for (int i = 0; i < 4; ++i)
        a[i] -= b[i] * a_dot_b / b_sq;
}
/* ref: tetrads.cpp:68-124 */
void orc_make_tetrad(const double u_con[4], double trial[4], double g[4][4], double e_con[4][4],
                     double e_cov[4][4]) {
    for (int i = 0; i < 4; ++i)
        e_con[0][i] = u_con[i];
    normalize(e_con[0], g);
    double norm = dot_g(trial, trial, g);
    if (norm < 1.0e-30) {
        for (int i = 0; i < 4; ++i)
            trial[i] = (i == 1) ? 1.0 : 0.0;
    }
    for (int i = 0; i < 4; ++i)
        e_con[1][i] = trial[i];
    project_out(e_con[1], e_con[0], g);
    normalize(e_con[1], g);
    for (int i = 0; i < 4; ++i)
        e_con[2][i] = (i == 2) ? 1.0 : 0.0;
    project_out(e_con[2], e_con[0], g);
    project_out(e_con[2], e_con[1], g);
    normalize(e_con[2], g);
    for (int i = 0; i < 4; ++i)
        e_con[3][i] = (i == 3) ? 1.0 : 0.0;
    project_out(e_con[3], e_con[0], g);
    project_out(e_con[3], e_con[1], g);
    project_out(e_con[3], e_con[2], g);
    normalize(e_con[3], g);
    for (int i = 0; i < 4; ++i)
        lower(e_con[i], g, e_cov[i]);
    for (int i = 0; i < 4; ++i)
        e_cov[0][i] *= -1.0;
}
/* ref: tetrads.cpp:46-55 */
static void coordinate_to_tetrad(double e_cov[4][4], const double k[4], double kt[4]) {
    for (int i = 0; i < 4; ++i) {
        kt[i] = 0.0;
        for (int j = 0; j < 4; ++j)
            kt[i] += e_cov[i][j] * k[j];
    }
}
/* ref: tetrads.cpp:57-66 */
static void tetrad_to_coordinate(double e_con[4][4], const double kt[4], double k[4]) {
    for (int i = 0; i < 4; ++i) {
        k[i] = 0.0;
        for (int j = 0; j < 4; ++j)
            k[i] += e_con[j][i] * kt[j];
    }
}

/* ============================================================================================
 * Samplers, ref: proba.cpp:30-215
 * ============================================================================================ */
/* ref: proba.cpp:202-210 */
static void sample_rand_dir(const orc_model *m, orc_rng *r, double *x, double *y, double *z) {
    *z = orc_uniform(m, r) * 2.0 - 1.0;
    double phi = orc_uniform(m, r) * 2.0 * M_PI;
    *x = sqrt(1.0 - *z * *z) * cos(phi);
    *y = sqrt(1.0 - *z * *z) * sin(phi);
}

/* ref: proba.cpp:123-166 */
double orc_sample_y(const orc_model *m, orc_rng *r, double theta_e) {
    double pi_3 = sqrt(M_PI) / 4.0;
    double pi_4 = sqrt(0.5 * theta_e) / 2.0;
    double pi_5 = 3.0 * sqrt(M_PI) * theta_e / 8.0;
    double pi_6 = theta_e * sqrt(0.5 * theta_e);
    double s_3 = pi_3 + pi_4 + pi_5 + pi_6;
    pi_3 /= s_3;
    pi_4 /= s_3;
    pi_5 /= s_3;
    pi_6 /= s_3;
    double y, x2, prob;
    do {
        double x1 = orc_uniform(m, r);
        int dof;
        if (x1 < pi_3)
            dof = 3;
        else if (x1 < pi_3 + pi_4)
            dof = 4;
        else if (x1 < pi_3 + pi_4 + pi_5)
            dof = 5;
        else
            dof = 6;
        double x = orc_chi_sq(m, r, dof);
        y = sqrt(x / 2.0);
        x2 = orc_uniform(m, r);
        double num = sqrt(1.0 + 0.5 * theta_e * y * y);
        double den = (1.0 + y * sqrt(0.5 * theta_e));
        prob = num / den;
    } while (x2 >= prob);
    return y;
}

/* ref: proba.cpp:168-172 */
double orc_sample_mu(const orc_model *m, orc_rng *r, double beta_e) {
    double x1 = orc_uniform(m, r);
    double det = 1.0 + 2.0 * beta_e + beta_e * beta_e - 4.0 * beta_e * x1;
    return (1.0 - sqrt(det)) / beta_e;
}

/* ref: proba.cpp:30-112 (the 1e7-trial escape hatch at :61-64 is dead code, Appendix A.9) */
void orc_sample_electron(const orc_model *m, orc_rng *r, const double k[4], double theta_e, double p[4]) {
    double x1, sigma_kn, gamma_e, beta_e, mu;
    do {
        double y = orc_sample_y(m, r, theta_e); /* ref: proba.cpp:114-121 */
        gamma_e = y * y * theta_e + 1.0;
        beta_e = sqrt(1.0 - 1.0 / (gamma_e * gamma_e));
        mu = orc_sample_mu(m, r, beta_e);
        if (mu > 1.0)
            mu = 1.0;
        else if (mu < -1.0)
            mu = -1.0;
        double k_ = gamma_e * (1.0 - beta_e * mu) * k[0];
        if (k_ < 1.0e-3) {
            sigma_kn = 1.0 - 2.0 * k_;
        } else {
            sigma_kn = (3.0 / (4.0 * k_ * k_)) * (2.0 + k_ * k_ * (1.0 + k_) / ((1.0 + 2.0 * k_) * (1.0 + 2.0 * k_)) +
                                                  (k_ * k_ - 2.0 * k_ - 2.0) / (2.0 * k_) * log(1.0 + 2.0 * k_));
        }
        x1 = orc_uniform(m, r);
    } while (x1 >= sigma_kn);

    double v0x = k[1], v0y = k[2], v0z = k[3];
    double v0 = sqrt(v0x * v0x + v0y * v0y + v0z * v0z);
    v0x /= v0;
    v0y /= v0;
    v0z /= v0;
    double n0x, n0y, n0z;
    sample_rand_dir(m, r, &n0x, &n0y, &n0z);
    double n0dotv0 = v0x * n0x + v0y * n0y + v0z * n0z;
    double v1x = n0x - (n0dotv0)*v0x;
    double v1y = n0y - (n0dotv0)*v0y;
    double v1z = n0z - (n0dotv0)*v0z;
    double v1 = sqrt(v1x * v1x + v1y * v1y + v1z * v1z);
    v1x /= v1;
    v1y /= v1;
    v1z /= v1;
    double v2x = v0y * v1z - v0z * v1y;
    double v2y = v0z * v1x - v0x * v1z;
    double v2z = v0x * v1y - v0y * v1x;
    double phi = orc_uniform(m, r) * 2.0 * M_PI;
    double s_phi = sin(phi), c_phi = cos(phi);
    double c_th = mu;
    double s_th = sqrt(1. - mu * mu);
    p[0] = gamma_e;
    p[1] = gamma_e * beta_e * (c_th * v0x + s_th * (c_phi * v1x + s_phi * v2x));
    p[2] = gamma_e * beta_e * (c_th * v0y + s_th * (c_phi * v1y + s_phi * v2y));
    p[3] = gamma_e * beta_e * (c_th * v0z + s_th * (c_phi * v1z + s_phi * v2z));
}

/* ref: proba.cpp:212-215 */
static double klein_nishina(double a, double ap) {
    double ch = 1.0 + 1.0 / a - 1.0 / ap;
    return (a / ap + ap / a - 1.0 + ch * ch) / (a * a);
}

/* ref: proba.cpp:174-189 */
double orc_sample_klein_nishina(const orc_model *m, orc_rng *r, double k0) {
    double k0pmin = k0 / (1.0 + 2.0 * k0);
    double k0pmax = k0;
    double x1, k0p_tent;
    do {
        k0p_tent = k0pmin + (k0pmax - k0pmin) * orc_uniform(m, r);
        x1 = 2.0 * (1.0 + 2.0 * k0 + 2.0 * k0 * k0) / (k0 * k0 * (1.0 + 2.0 * k0));
        x1 *= orc_uniform(m, r);
    } while (x1 >= klein_nishina(k0, k0p_tent));
    return k0p_tent;
}

/* ref: proba.cpp:191-200 */
double orc_sample_thomson(const orc_model *m, orc_rng *r) {
    double x1, x2;
    do {
        x1 = 2.0 * orc_uniform(m, r) - 1.0;
        x2 = (3.0 / 4.0) * orc_uniform(m, r);
    } while (x2 >= (3.0 / 8.0) * (1.0 + x1 * x1));
    return x1;
}

/* ref: harm_model.cpp:1658-1671 */
static void boost(const double v[4], const double u[4], double vp[4]) {
    double g = u[0];
    double v_ = sqrt(fabs(1.0 - 1.0 / (g * g)));
    double n1 = u[1] / (g * v_ + C_EPS);
    double n2 = u[2] / (g * v_ + C_EPS);
    double n3 = u[3] / (g * v_ + C_EPS);
    double gm1 = g - 1.0;
    vp[0] = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
    vp[1] = -u[1] * v[0] + (1.0 + n1 * n1 * gm1) * v[1] + n1 * n2 * gm1 * v[2] + n1 * n3 * gm1 * v[3];
    vp[2] = -u[2] * v[0] + n2 * n1 * gm1 * v[1] + (1.0 + n2 * n2 * gm1) * v[2] + n2 * n3 * gm1 * v[3];
    vp[3] = -u[3] * v[0] + n3 * n1 * gm1 * v[1] + n3 * n2 * gm1 * v[2] + (1.0 + n3 * n3 * gm1) * v[3];
}

/* ref: harm_model.cpp:1147-1215 */
void orc_sample_scattered_photon(const orc_model *m, orc_rng *r, const double k[4], double p[4], double kp[4]) {
    double ke[4];
    boost(k, p, ke);
    double k0p, c_th;
    if (ke[0] > 1.0e-4) {
        k0p = orc_sample_klein_nishina(m, r, ke[0]);
        c_th = 1.0 - 1.0 / k0p + 1.0 / ke[0];
    } else {
        k0p = ke[0];
        c_th = orc_sample_thomson(m, r);
    }
    double s_th = sqrt(fabs(1.0 - c_th * c_th));
    double v0x = ke[1] / ke[0], v0y = ke[2] / ke[0], v0z = ke[3] / ke[0];
    double n0x, n0y, n0z;
    sample_rand_dir(m, r, &n0x, &n0y, &n0z);
    double n0dotv0 = v0x * n0x + v0y * n0y + v0z * n0z;
    double v1x = n0x - (n0dotv0)*v0x;
    double v1y = n0y - (n0dotv0)*v0y;
    double v1z = n0z - (n0dotv0)*v0z;
    double v1 = sqrt(v1x * v1x + v1y * v1y + v1z * v1z);
    v1x /= v1;
    v1y /= v1;
    v1z /= v1;
    double v2x = v0y * v1z - v0z * v1y;
    double v2y = v0z * v1x - v0x * v1z;
    double v2z = v0x * v1y - v0y * v1x;
    double phi = 2.0 * M_PI * orc_uniform(m, r);
    double s_phi = sin(phi), c_phi = cos(phi);
    p[1] *= -1.;
    p[2] *= -1.;
    p[3] *= -1.;
    double dir1 = c_th * v0x + s_th * (c_phi * v1x + s_phi * v2x);
    double dir2 = c_th * v0y + s_th * (c_phi * v1y + s_phi * v2y);
    double dir3 = c_th * v0z + s_th * (c_phi * v1z + s_phi * v2z);
    double kpe[4] = {k0p, k0p * dir1, k0p * dir2, k0p * dir3};
    boost(kpe, p, kp);
}

/* ref: harm_model.cpp:1071-1145.  Returns 1 if the child photon is valid.
 * Where the reference leaves the child uninitialised (early return at :1106-1108, Appendix A.14) or sets
 * its weight to zero (:1118-1121) we return 0: the child is dropped. */
int orc_scatter_super_photon(const orc_model *m, orc_photon *ph, orc_photon *php, const orc_fluid *f,
                             double gcov[4][4]) {
    if (ph->k[0] > 1.0e5 || ph->k[0] < 0.0 || isnan(ph->k[0]) || isnan(ph->k[1]) || isnan(ph->k[3])) {
        ph->k[0] = fabs(ph->k[0]);
        ph->w = 0.0;
        return 0;
    }
    double b_hat[4];
    if (f->b > 0.0) {
        for (int i = 0; i < 4; ++i)
            b_hat[i] = f->b_con[i] / (f->b / m->b_unit);
    } else {
        for (int i = 0; i < 4; ++i)
            b_hat[i] = 0.0;
        b_hat[1] = 1.0;
    }
    double e_con[4][4], e_cov[4][4];
    orc_make_tetrad(f->u_con, b_hat, gcov, e_con, e_cov);
    double kt[4];
    coordinate_to_tetrad(e_cov, ph->k, kt);
    if (kt[0] > 1.0e5 || kt[0] < 0.0 || isnan(kt[1]))
        return 0;
    double p[4];
    orc_sample_electron(m, &ph->rng, kt, f->theta_e, p);
    double ktp[4];
    orc_sample_scattered_photon(m, &ph->rng, kt, p, ktp);
    tetrad_to_coordinate(e_con, ktp, php->k);
    if (isnan(php->k[1])) {
        php->w = 0.0;
        return 0;
    }
    { /* diagnostics: |k.k| relative to the size of its terms */
        double kk = 0.0, sz = 0.0;
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) {
                kk += gcov[a][b] * php->k[a] * php->k[b];
                sz += fabs(gcov[a][b] * php->k[a] * php->k[b]);
            }
        const double rel = fabs(kk) / (sz + 1e-300);
        if (rel > orc_diag_max_offcone)
            orc_diag_max_offcone = rel;
        if (rel > 1e-6)
            orc_diag_n_offcone++;
    }
    double tmp_k[4];
    ktp[0] *= -1.0;
    tetrad_to_coordinate(e_cov, ktp, tmp_k);
    php->e = -tmp_k[0];
    php->e_0_s = -tmp_k[0];
    php->l = tmp_k[3];
    php->tau_abs = 0.0;
    php->tau_scatt = 0.0;
    php->b_0 = f->b;
    php->x1i = ph->x[1];
    php->x2i = ph->x[2];
    for (int i = 0; i < 4; ++i)
        php->x[i] = ph->x[i];
    php->n_e_0 = ph->n_e_0;
    php->theta_e_0 = ph->theta_e_0;
    php->e_0 = ph->e_0;
    php->n_scatt = ph->n_scatt + 1;
    return 1;
}

/* ============================================================================================
 * Generation
 * ============================================================================================ */
/* ref: harm_model.cpp:1337-1389 */
void orc_init_zone(const orc_model *m, int i, int j, double *nz_out, double *dn_max_out) {
    *nz_out = 0.0;
    *dn_max_out = 0.0;
    orc_fluid fz;
    orc_get_fluid_zone(m, i, j, &fz);
    if (fz.n_e == 0.0 || fz.theta_e < C_THETA_E_MIN)
        return;
    double l_bth = log(fz.b * fz.theta_e * fz.theta_e);
    double d_l = (l_bth - c_l_b_min()) / c_d_l_b();
    int l = (int)d_l;
    d_l -= l;
    if (l < 0)
        return;
    double ninterp = 0.0, dn_max = 0.0;
    if (l >= ORC_NINT) {
        /* out-of-table branch, ref :1358-1369, including its index slip (x_2 where i was meant, Appendix A.9) */
        for (int s = 0; s <= ORC_N_E_SAMP; ++s) {
            double dn = orc_f_eval(m, fz.theta_e, fz.b, exp(j * c_d_l_nu() + c_l_nu_min())) /
                        (exp(m->weight[s]) + 1.0e-100);
            if (dn > dn_max)
                dn_max = dn;
            ninterp += c_d_l_nu() * dn;
        }
    } else if (!isinf(m->nint[l]) && !isinf(m->nint[l + 1])) {
        ninterp = exp((1.0 - d_l) * m->nint[l] + d_l * m->nint[l + 1]);
        dn_max = exp((1.0 - d_l) * m->dndlnu_max[l] + d_l * m->dndlnu_max[l + 1]);
    }
    double k2 = orc_k2_eval(m, fz.theta_e);
    if (k2 == 0.0)
        return;
    double nz = m->geom_det[i * m->n1 + j] * fz.n_e * fz.b * fz.theta_e * fz.theta_e * ninterp / k2;
    if (nz > m->photon_n * log(C_NU_MAX / C_NU_MIN))
        return;
    *nz_out = nz;
    *dn_max_out = dn_max;
}

/* ref: harm_model.cpp:673-704 -- one uniform per zone decides the rounding; here it comes from the zone's
 * own Philox stream instead of the global stream. */
uint64_t orc_zone_counts(const orc_model *m, int64_t *num_to_gen, double *dn_max) {
    uint64_t total = 0;
    for (int i = 0; i < m->n0; ++i) {
        for (int j = 0; j < m->n1; ++j) {
            int z = i * m->n1 + j;
            double nz, dm;
            orc_init_zone(m, i, j, &nz, &dm);
            orc_rng r;
            orc_rng_zone(&r, (uint64_t)z);
            double u = orc_uniform(m, &r);
            int64_t n = (int64_t)nz;
            if (fmod(nz, 1.0) > u)
                n += 1;
            num_to_gen[z] = n;
            dn_max[z] = dm;
            total += (uint64_t)n;
        }
    }
    return total;
}

/* ref: harm_model.cpp:784-792 */
static double linear_interp_weight(const orc_model *m, double nu) {
    double l_nu = log(nu);
    double d_i = (l_nu - c_l_nu_min()) / c_d_l_nu();
    int i = (int)d_i;
    if (i > ORC_N_E_SAMP - 1)
        i = ORC_N_E_SAMP - 1;
    d_i -= i;
    return exp((1.0 - d_i) * m->weight[i] + d_i * m->weight[i + 1]);
}

/* ref: harm_model.cpp:706-782 (+ the InitPhoton -> Photon copy at :373-391) */
void orc_sample_zone_photon(const orc_model *m, int i, int j, double dn_max, orc_rng *r, orc_photon *ph) {
    memset(ph, 0, sizeof(*ph));
    get_coord(m, i, j, ph->x);
    orc_fluid fz;
    orc_get_fluid_zone(m, i, j, &fz);
    double b_hat[4];
    if (fz.b > 0.0) {
        for (int d = 0; d < 4; ++d)
            b_hat[d] = fz.b_con[d] * m->b_unit / fz.b;
    } else {
        for (int d = 1; d < 4; ++d)
            b_hat[d] = 0.0;
        b_hat[0] = 1.0;
    }
    double gcov[4][4], e_con[4][4], e_cov[4][4];
    orc_gcov(m, ph->x, gcov);
    orc_make_tetrad(fz.u_con, b_hat, gcov, e_con, e_cov);

    double nu, weight;
    do {
        nu = exp(orc_uniform(m, r) * c_n_l_n() + c_l_nu_min());
        weight = linear_interp_weight(m, nu);
    } while (orc_uniform(m, r) > (orc_f_eval(m, fz.theta_e, fz.b, nu) / (weight + 1.0e-100)) / dn_max);
    ph->w = weight;
    double j_max = orc_synch(m, nu, fz.n_e, fz.theta_e, fz.b, M_PI / 2.0);
    double cos_th, th;
    do {
        cos_th = 2.0 * orc_uniform(m, r) - 1.0;
        th = acos(cos_th);
    } while (orc_uniform(m, r) > (orc_synch(m, nu, fz.n_e, fz.theta_e, fz.b, th) / j_max));
    double sin_th = sqrt(1.0 - cos_th * cos_th);
    double phi = 2.0 * M_PI * orc_uniform(m, r);
    double cos_phi = cos(phi), sin_phi = sin(phi);
    double e = nu * C_HPL / (C_ME * C_CL * C_CL);
    double kt[4] = {e, e * cos_th, e * sin_th * cos_phi, e * sin_th * sin_phi};
    tetrad_to_coordinate(e_con, kt, ph->k);
    kt[0] *= -1.0;
    double tmp_k[4];
    tetrad_to_coordinate(e_cov, kt, tmp_k);
    ph->e = -tmp_k[0];
    ph->e_0 = -tmp_k[0];
    ph->e_0_s = ph->e;
    ph->l = tmp_k[3];
    ph->n_e_0 = fz.n_e;
    ph->theta_e_0 = fz.theta_e;
    ph->b_0 = fz.b;
    ph->n_scatt = 0;
    ph->tau_abs = 0.0;
    ph->tau_scatt = 0.0;
    ph->x1i = ph->x[1];
    ph->x2i = ph->x[2];
    ph->rng = *r;
}

/* ============================================================================================
 * Record, ref: harm_model.cpp:1291-1335
 * ============================================================================================ */
enum { F_DN_DLE, F_DE_DLE, F_NPH, F_NSCATT, F_X1I_AV, F_X2I_SQ, F_X3F_SQ, F_TAU_ABS, F_TAU_SCATT, F_NE_0,
       F_THETA_E_0, F_B_0, F_E_0 };

void orc_record_super_photon(orc_model *m, const orc_photon *ph) {
    if (isnan(ph->w) || isnan(ph->e))
        return;
    if (ph->tau_scatt > m->acc_max_tau_scatt)
        m->acc_max_tau_scatt = ph->tau_scatt;
    if (m->stats_mode == ORC_STATS_LIVE)
        m->bias_max_tau_scatt = m->acc_max_tau_scatt;
    double dx2 = (m->x_stop2 - m->x_start2) / (2.0 * ORC_N_TH_BINS);
    int ix2;
    if (ph->x[2] < 0.5 * (m->x_start2 + m->x_stop2))
        ix2 = (int)(ph->x[2] / dx2);
    else
        ix2 = (int)((m->x_stop2 - ph->x[2]) / dx2);
    if (ix2 < 0 || ix2 >= ORC_N_TH_BINS)
        return;
    double l_e = log(ph->e);
    int i_e = (int)((l_e - c_spec_l_e_0()) / C_SPEC_D_L_E + 2.5) - 2;
    if (i_e < 0 || i_e >= ORC_N_E_BINS)
        return;
    m->acc_n_recorded += 1;
    m->acc_n_scatt += (uint64_t)ph->n_scatt;
    if (m->stats_mode == ORC_STATS_LIVE) {
        m->bias_n_recorded = (double)m->acc_n_recorded;
        m->bias_n_scatt = (double)m->acc_n_scatt;
    }
    double *s = m->spectrum[ix2][i_e];
    s[F_DN_DLE] += ph->w;
    s[F_DE_DLE] += ph->w * ph->e;
    s[F_TAU_ABS] += ph->w * ph->tau_abs;
    s[F_TAU_SCATT] += ph->w * ph->tau_scatt;
    s[F_X1I_AV] += ph->w * ph->x1i;
    s[F_X2I_SQ] += ph->w * (ph->x2i * ph->x2i);
    s[F_X3F_SQ] += ph->w * (ph->x[3] * ph->x[3]);
    s[F_NE_0] += ph->w * (ph->n_e_0);
    s[F_B_0] += ph->w * (ph->b_0);
    s[F_THETA_E_0] += ph->w * (ph->theta_e_0);
    s[F_NSCATT] += ph->n_scatt;
    s[F_NPH] += 1.0;
}

/* ============================================================================================
 * Transport, ref: harm_model.cpp:894-1069
 * ============================================================================================ */
static void eval_opacities(const orc_model *m, const orc_photon *ph, const orc_fluid *f, double *alpha_scatt,
                           double *alpha_abs, double *nu_out) {
    double theta = orc_bk_angle(m, ph->k, f->u_cov, f->b_cov, f->b);
    double nu = orc_fluid_nu(ph->k, f->u_cov);
    *nu_out = nu;
    *alpha_scatt = orc_alpha_inv_scatt(m, nu, f->theta_e, f->n_e);
    *alpha_abs = orc_alpha_inv_abs(m, nu, f->theta_e, f->n_e, f->b, theta);
}

/* Resumable form of track_super_photon.  The loop variables the reference keeps on the stack
 * (alpha_scatti, alpha_absi, bi, the n_e > 0 latch, n_step) live in orc_track_state so that a photon can be
 * SUSPENDED at a step boundary and continued later: the CUDA path bounds the tail of each generation by
 * letting every photon lineage make at most `budget` push attempts per generation ("generation clock",
 * DESIGN.md); a suspended photon continues in the next generation, with that generation's frozen bias
 * statistics.  With budget = INT_MAX this is exactly the reference's control flow. */
static void carry_push(orc_model *m, const orc_track_state *s) {
    if (m->n_carry == m->cap_carry) {
        m->cap_carry = m->cap_carry ? 2 * m->cap_carry : 1024;
        m->carry = (orc_track_state *)realloc(m->carry, m->cap_carry * sizeof(orc_track_state));
    }
    m->carry[m->n_carry] = *s;
    m->carry[m->n_carry].clock = 0;
    m->n_carry++;
}

/* reference :894-917: validation and start-of-track quantities; returns 0 if the photon is invalid */
static int track_begin(orc_model *m, const orc_photon *ph_in, int clock, orc_track_state *s) {
    s->ph = *ph_in;
    orc_photon *ph = &s->ph;
    for (int i = 0; i < 4; ++i) {
        if (isnan(ph->x[i]) || isnan(ph->k[i]))
            return 0;
    }
    if (ph->w == 0.0)
        return 0;
    m->n_tracked++;
    double gcov[4][4];
    orc_fluid f;
    orc_gcov(m, ph->x, gcov);
    orc_get_fluid_params(m, ph->x, gcov, &f);
    s->alpha_scatti = s->alpha_absi = s->bi = 0.0;
    double nu = 0.0;
    if (f.n_e > 0.0) {
        eval_opacities(m, ph, &f, &s->alpha_scatti, &s->alpha_absi, &nu);
        if (nu < 0.0 || isnan(nu)) /* unphysical wave-vector: the CUDA path defines alpha = 0 here */
            s->alpha_scatti = s->alpha_absi = 0.0;
        s->bi = orc_bias_func(m, f.theta_e, ph->w);
    }
    /* (photons are born/scattered inside the grid, so the reference's unguarded evaluation at :907-913
     *  always sees valid fluid data; outside we define alpha = 0.) */
    s->ne_pos = f.n_e > 0.0;
    orc_init_dkdlam(m, ph->x, ph->k, ph->dkdlam);
    s->n_step = 0;
    s->clock = clock;
    return 1;
}

/* reference :919-1068.  Returns 1 if the photon was suspended (it is then on m->carry), else 0. */
static int track_loop(orc_model *m, orc_track_state *s) {
    orc_photon *ph = &s->ph;
    double gcov[4][4];
    orc_fluid f;
    double nu = 0.0;
    for (;;) {
        if (m->budget > 0 && s->clock >= m->budget) {
            carry_push(m, s);
            return 1;
        }
        if (orc_stop_criterion(m, ph))
            break;
        double x2[4], k2[4], dk2[4], e0s2;
        memcpy(x2, ph->x, sizeof(x2));
        memcpy(k2, ph->k, sizeof(k2));
        memcpy(dk2, ph->dkdlam, sizeof(dk2));
        e0s2 = ph->e_0_s;
        double dl = orc_step_size(m, ph->x, ph->k);
        uint64_t att0 = m->n_push_attempts;
        orc_push_photon(m, ph, dl, 0);
        s->clock += (int)(m->n_push_attempts - att0);
        m->n_steps++;
        if (orc_stop_criterion(m, ph))
            break;

        if (s->alpha_absi > 0.0 || s->alpha_scatti > 0.0 || s->ne_pos) {
            m->n_interactions++;
            orc_gcov(m, ph->x, gcov);
            orc_get_fluid_params(m, ph->x, gcov, &f);
            s->ne_pos = f.n_e > 0.0;
            int bound_flag = (f.n_e == 0.0);
            double theta = 0.0;
            if (!bound_flag) {
                theta = orc_bk_angle(m, ph->k, f.u_cov, f.b_cov, f.b);
                nu = orc_fluid_nu(ph->k, f.u_cov);
            }
            double d_tau_scatt, d_tau_abs, bias;
            if (bound_flag || nu < 0.0) {
                d_tau_scatt = 0.5 * s->alpha_scatti * m->d_tau_k * dl;
                d_tau_abs = 0.5 * s->alpha_absi * m->d_tau_k * dl;
                s->alpha_scatti = 0.0;
                s->alpha_absi = 0.0;
                bias = 0.0;
                s->bi = 0.0;
            } else {
                double alpha_scattf = orc_alpha_inv_scatt(m, nu, f.theta_e, f.n_e);
                d_tau_scatt = 0.5 * (s->alpha_scatti + alpha_scattf) * m->d_tau_k * dl;
                s->alpha_scatti = alpha_scattf;
                double alpha_absf = orc_alpha_inv_abs(m, nu, f.theta_e, f.n_e, f.b, theta);
                d_tau_abs = 0.5 * (s->alpha_absi + alpha_absf) * m->d_tau_k * dl;
                s->alpha_absi = alpha_absf;
                double bf = orc_bias_func(m, f.theta_e, ph->w);
                bias = 0.5 * (s->bi + bf);
                s->bi = bf;
            }
            double x1 = -log(orc_uniform(m, &ph->rng));
            orc_photon php;
            memset(&php, 0, sizeof(php));
            php.w = ph->w / bias;
            if (bias * d_tau_scatt > x1 && php.w > C_WEIGHT_MIN) {
                orc_rng_child(m, &ph->rng, &php.rng);
                double frac = x1 / (bias * d_tau_scatt);
                d_tau_abs *= frac;
                if (d_tau_abs > 100)
                    return 0; /* absorbed before scattering */
                d_tau_scatt *= frac;
                double d_tau = d_tau_abs + d_tau_scatt;
                if (d_tau_abs < 1.0e-3)
                    ph->w *= (1.0 - d_tau / 24.0 * (24.0 - d_tau * (12.0 - d_tau * (4.0 - d_tau))));
                else
                    ph->w *= exp(-d_tau);
                /* back up to the scattering point: re-push the pre-step snapshot by dl*frac */
                memcpy(ph->x, x2, sizeof(x2));
                memcpy(ph->k, k2, sizeof(k2));
                memcpy(ph->dkdlam, dk2, sizeof(dk2));
                ph->e_0_s = e0s2;
                att0 = m->n_push_attempts;
                orc_push_photon(m, ph, dl * frac, 0);
                s->clock += (int)(m->n_push_attempts - att0);
                orc_gcov(m, ph->x, gcov);
                orc_get_fluid_params(m, ph->x, gcov, &f);
                s->ne_pos = f.n_e > 0.0;
                if (f.n_e > 0.0) {
                    m->n_scatter_events++;
                    int child_ok = orc_scatter_super_photon(m, ph, &php, &f, gcov);
                    if (ph->w < 1.0e-100)
                        return 0;
                    if (child_ok) {
                        /* the child inherits its lineage's generation clock */
                        orc_track_state cs;
                        if (track_begin(m, &php, s->clock, &cs))
                            track_loop(m, &cs);
                    }
                    theta = orc_bk_angle(m, ph->k, f.u_cov, f.b_cov, f.b);
                    nu = orc_fluid_nu(ph->k, f.u_cov);
                    if (nu < 0.0) {
                        s->alpha_scatti = 0.0;
                        s->alpha_absi = 0.0;
                    } else {
                        s->alpha_scatti = orc_alpha_inv_scatt(m, nu, f.theta_e, f.n_e);
                        s->alpha_absi = orc_alpha_inv_abs(m, nu, f.theta_e, f.n_e, f.b, theta);
                    }
                    s->bi = orc_bias_func(m, f.theta_e, ph->w);
                } else {
                    /* left the grid while backing up: the reference reads uninitialised fluid data here
                     * (Appendix A.15); we define alpha = 0, bias = 0 (the gate then latches off). */
                    s->alpha_scatti = 0.0;
                    s->alpha_absi = 0.0;
                    s->bi = 0.0;
                }
            } else {
                if (d_tau_abs > 100)
                    return 0; /* absorbed */
                double d_tau = d_tau_abs + d_tau_scatt;
                if (d_tau < 1.0e-3)
                    ph->w *= (1. - d_tau / 24. * (24. - d_tau * (12. - d_tau * (4. - d_tau))));
                else
                    ph->w *= exp(-d_tau);
            }
            ph->tau_abs += d_tau_abs;
            ph->tau_scatt += d_tau_scatt;
        }
        ++s->n_step;
        if (s->n_step > C_MAX_N_STEP)
            break;
    }
    if (s->n_step > orc_diag_max_n_step) { /* diagnostics: the longest-lived photon seen and where it ended */
        orc_diag_max_n_step = s->n_step;
        for (int i = 0; i < 4; ++i) {
            orc_diag_longest[i] = ph->x[i];
            orc_diag_longest[4 + i] = ph->k[i];
        }
        orc_diag_longest[8] = ph->w;
        orc_diag_longest[9] = (double)ph->n_scatt;
    }
    if (ph->x[1] > c_x1_max() && s->n_step <= C_MAX_N_STEP)
        orc_record_super_photon(m, ph);
    return 0;
}

/* track one photon (and all its descendants); photons that run out of attempts end up on m->carry */
void orc_track_super_photon(orc_model *m, orc_photon *ph) {
    orc_track_state s;
    if (!track_begin(m, ph, 0, &s))
        return;
    track_loop(m, &s);
    *ph = s.ph;
}

/* continue the photons suspended so far (they may be suspended again) */
static void run_carried(orc_model *m, int clock0) {
    size_t n = m->n_carry;
    if (!n)
        return;
    orc_track_state *list = (orc_track_state *)malloc(n * sizeof(orc_track_state));
    memcpy(list, m->carry, n * sizeof(orc_track_state));
    m->n_carry = 0;
    for (size_t i = 0; i < n; ++i) {
        list[i].clock = clock0;
        track_loop(m, &list[i]);
    }
    free(list);
}

/* ============================================================================================
 * Driver: generations of primaries (shared schedule with the CUDA path)
 * ============================================================================================ */
int64_t orc_generation_size(int64_t g_start, int64_t gen0, int64_t gen_cap, int64_t fine_from, int64_t fine_div,
                            int64_t ramp) {
    /* size of the generation that starts at run position g_start: gen0 first, then (ramp - 1) times the positions
     * started so far (the cumulative count grows ramp-fold) until fine_from, then 1/fine_div of the cumulative count
     * (growth factor 1 + 1/fine_div per generation), never more than gen_cap */
    int64_t s;
    if (g_start < gen0)
        s = gen0;
    else if (g_start < fine_from || fine_div <= 1)
        s = g_start * ((ramp > 2 ? ramp : 2) - 1);
    else
        s = g_start / fine_div;
    if (s < gen0)
        s = gen0;
    return s < gen_cap ? s : gen_cap;
}

static void idx_to_zone(const orc_model *m, const int64_t *prefix, int64_t idx, int *pi, int *pj) {
    /* prefix[z] = number of primaries in zones < z; find z with prefix[z] <= idx < prefix[z+1] */
    int64_t lo = 0, hi = (int64_t)m->n0 * m->n1;
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) / 2;
        if (prefix[mid] <= idx)
            lo = mid;
        else
            hi = mid;
    }
    *pi = (int)(lo / m->n1);
    *pj = (int)(lo % m->n1);
}

void orc_make_primary(const orc_model *m, const int64_t *prefix, const double *dn_max, int64_t idx,
                      orc_photon *ph) {
    int i, j;
    idx_to_zone(m, prefix, idx, &i, &j);
    orc_rng r;
    orc_rng_primary(&r, (uint64_t)idx);
    orc_sample_zone_photon(m, i, j, dn_max[i * m->n1 + j], &r, ph);
}

void orc_run_primary(orc_model *m, const int64_t *prefix, const double *dn_max, int64_t idx, int clock0) {
    orc_photon ph;
    orc_make_primary(m, prefix, dn_max, idx, &ph);
    orc_track_state s;
    if (track_begin(m, &ph, clock0, &s))
        track_loop(m, &s);
    m->n_created++;
}

/* Processing order.  Position j of the run handles primary pi(j) = (j * mult) mod total with mult ~ total/phi
 * coprime to total (a Weyl sequence): every contiguous range of positions samples the whole zone-ordered
 * index range evenly, so the bias statistics frozen at a generation start are representative of all zones.
 * (The reference walks the zones in order, harm_model.cpp:673-704, and updates the statistics after every
 * photon; frozen statistics in zone order would see only the horizon-crossing inner zones at first.)
 * The photon itself -- its Philox stream and birth zone -- depends only on the primary index, not on j. */
static int64_t gcd64(int64_t a, int64_t b) {
    while (b) {
        int64_t t = a % b;
        a = b;
        b = t;
    }
    return a;
}
int64_t orc_perm_multiplier(int64_t total) {
    if (total < 3)
        return 1;
    int64_t mult = (int64_t)((double)total * 0.6180339887498949);
    if (mult < 1)
        mult = 1;
    while (gcd64(mult, total) != 1)
        ++mult;
    return mult;
}
int64_t orc_permute(int64_t j, int64_t mult, int64_t total) {
    return (int64_t)(((unsigned __int128)(uint64_t)j * (uint64_t)mult) % (uint64_t)total);
}

void orc_run(orc_model *m, int64_t first, int64_t last, int rank, int world, int64_t gen0, int64_t gen_cap) {
    int64_t nz = (int64_t)m->n0 * m->n1;
    int64_t *num = (int64_t *)malloc(sizeof(int64_t) * nz);
    int64_t *prefix = (int64_t *)malloc(sizeof(int64_t) * (nz + 1));
    double *dn_max = (double *)malloc(sizeof(double) * nz);
    uint64_t total = orc_zone_counts(m, num, dn_max);
    prefix[0] = 0;
    for (int64_t z = 0; z < nz; ++z)
        prefix[z + 1] = prefix[z] + num[z];
    if (last < 0 || last > (int64_t)total)
        last = (int64_t)total;
    const int64_t mult = orc_perm_multiplier((int64_t)total);
    /* generations partition the global index range [0,total); statistics freeze at generation starts */
    int64_t g_start = 0;
    const int budget = m->budget > 0 ? m->budget : INT_MAX;
    double lag_tau = m->acc_max_tau_scatt, lag_scatt = (double)m->acc_n_scatt, lag_rec = (double)m->acc_n_recorded;
    for (int64_t g = 0; g_start < last; ++g) {
        /* rank-local schedule, like the CUDA path (gm_api.cu grmonty_b200_run_range) */
        int64_t g_end = g_start + world * orc_generation_size(g_start / world, gen0, gen_cap, m->gen_fine_from,
                                                              m->gen_fine_div, m->gen_ramp);
        const int64_t lo = g_start > first ? g_start : first, hi = g_end < last ? g_end : last;
        if (lo < hi) {
            if (m->stats_mode == ORC_STATS_FROZEN) {
                /* stats_lag 1: every generation uses the snapshot taken one generation ago (statistics of the
                 * generations <= g - 2); 2: only the generations at the size cap do (the pipelined scheduler's default,
                 * gm_pipeline.cuh); 0: none */
                const int64_t size = orc_generation_size(g_start / world, gen0, gen_cap, m->gen_fine_from,
                                                         m->gen_fine_div, m->gen_ramp);
                const int lagged = m->stats_lag == 1 || (m->stats_lag == 2 && size >= gen_cap);
                if (lagged) {
                    m->bias_max_tau_scatt = lag_tau;
                    m->bias_n_scatt = lag_scatt;
                    m->bias_n_recorded = lag_rec;
                } else {
                    m->bias_max_tau_scatt = m->acc_max_tau_scatt;
                    m->bias_n_scatt = (double)m->acc_n_scatt;
                    m->bias_n_recorded = (double)m->acc_n_recorded;
                }
                lag_tau = m->acc_max_tau_scatt; /* the snapshot the next generation may lag to */
                lag_scatt = (double)m->acc_n_scatt;
                lag_rec = (double)m->acc_n_recorded;
            }
            m->budget = budget;
            /* Attempt budgets (same rule as the CUDA path, gm_api.cu run_batch): the t-th of the generation's
             * `count` primaries may make budget + (count - 1 - t) / spread attempts, carried lineages
             * budget + count / spread -- lanes take the primaries in order, so a lineage that starts early has the
             * rest of the generation to run without delaying its end.  A negative start value of the lineage
             * clock implements it. */
            int64_t f0 = lo + ((rank - lo % world) % world + world) % world;
            const int64_t count = f0 < hi ? (hi - f0 + world - 1) / world : 0;
            const int64_t spread = (budget != INT_MAX && m->gen_budget_spread > 0) ? m->gen_budget_spread : 0;
            run_carried(m, spread ? -(int)(count / spread) : 0); /* suspended photons: this generation's statistics */
            int64_t t = 0;
            for (int64_t j = f0; j < hi; j += world, ++t)
                orc_run_primary(m, prefix, dn_max, m->zone_order ? j : orc_permute(j, mult, (int64_t)total),
                                spread ? -(int)((count - 1 - t) / spread) : 0);
        }
        g_start = g_end;
    }
    /* drain: photons still suspended after the last generation run to completion */
    while (m->n_carry > 0) {
        if (m->stats_mode == ORC_STATS_FROZEN) {
            /* the drain generation starts when the last generation is complete: latest statistics */
            m->bias_max_tau_scatt = m->acc_max_tau_scatt;
            m->bias_n_scatt = (double)m->acc_n_scatt;
            m->bias_n_recorded = (double)m->acc_n_recorded;
        }
        m->budget = INT_MAX;
        run_carried(m, 0);
    }
    m->budget = budget == INT_MAX ? 0 : budget;
    free(num);
    free(prefix);
    free(dn_max);
}

void orc_clear_outputs(orc_model *m) {
    memset(m->spectrum, 0, sizeof(m->spectrum));
    m->n_created = 0;
    m->n_steps = m->n_push_attempts = m->n_interactions = m->n_scatter_events = m->n_tracked = 0;
    m->acc_n_scatt = 0;
    m->acc_n_recorded = 0;
}

/* ---- ctypes helpers -------------------------------------------------------------------------- */
void orc_photon_from_flat(const double *p, orc_photon *ph) {
    for (int i = 0; i < 4; ++i) {
        ph->x[i] = p[i];
        ph->k[i] = p[4 + i];
        ph->dkdlam[i] = p[8 + i];
    }
    ph->w = p[12];
    ph->e = p[13];
    ph->l = p[14];
    ph->x1i = p[15];
    ph->x2i = p[16];
    ph->tau_abs = p[17];
    ph->tau_scatt = p[18];
    ph->n_e_0 = p[19];
    ph->theta_e_0 = p[20];
    ph->b_0 = p[21];
    ph->e_0 = p[22];
    ph->e_0_s = p[23];
    ph->n_scatt = (int)p[24];
}
void orc_photon_to_flat(const orc_photon *ph, double *p) {
    for (int i = 0; i < 4; ++i) {
        p[i] = ph->x[i];
        p[4 + i] = ph->k[i];
        p[8 + i] = ph->dkdlam[i];
    }
    p[12] = ph->w;
    p[13] = ph->e;
    p[14] = ph->l;
    p[15] = ph->x1i;
    p[16] = ph->x2i;
    p[17] = ph->tau_abs;
    p[18] = ph->tau_scatt;
    p[19] = ph->n_e_0;
    p[20] = ph->theta_e_0;
    p[21] = ph->b_0;
    p[22] = ph->e_0;
    p[23] = ph->e_0_s;
    p[24] = ph->n_scatt;
}
orc_model *orc_model_alloc(void) { return (orc_model *)calloc(1, sizeof(orc_model)); }
void orc_model_free(orc_model *m) {
    if (m)
        free(m->carry);
    free(m);
}
unsigned long orc_sizeof_model(void) { return (unsigned long)sizeof(orc_model); }
