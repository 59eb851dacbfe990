/*
 * gm_radiation.cuh -- emissivity, absorption, hot Compton cross-section, bias; sm_100a device code.
 *
 * Reference: radiation.cpp:59-146 (bk_angle, fluid_nu, alpha_inv_scatt, alpha_inv_abs, b_nu_inv, jnu_inv,
 * kappa_es), jnu_mixed.cpp:75-125,150-168 (synch, k2_eval, f_eval), hotcross.cpp:81-179
 * (total_compton_cross_lkup and its numeric fall-back), harm_model.cpp:1391-1404 (bias_func).
 *
 * B200 notes: the synchrotron formula only needs sin(theta), so the pitch angle is carried as its cosine
 * (sqrt(1-mu^2) instead of acos + sin); pow(x,1/3) is cbrt; the table look-ups are tiny (hotcross 143 KB,
 * k2/f 1.6 KB) and stay in L1/L2 via the read-only path.
 */
#pragma once
#include "gm_fluid.cuh"
#include "gm_params.h"

namespace gm {

/* exp of a linear interpolation in a 201-entry log table (reference jnu_mixed.cpp:150-168;
 * upper index clamped) */
__device__ __forceinline__ double interp_exp_table(const double *tab, double lx, double l_min, double d_l) {
    double d_i = fm::div(lx - l_min, d_l);
    int i = (int)d_i;
    i = max(0, min(i, kNESamp - 1));
    d_i -= i;
    return fm::exp_((1.0 - d_i) * __ldg(tab + i) + d_i * __ldg(tab + i + 1));
}

/* K2(1/theta_e) given l_theta = ln(theta_e) (reference jnu_mixed::k2_eval, jnu_mixed.cpp:102-111); selects only */
__device__ __forceinline__ double k2_eval_l(const GmParams &P, double theta_e, double l_theta) {
    double d_i = (l_theta - P.jnu_l_min_t) * P.inv_jnu_d_l_t;
    int i = (int)d_i;
    i = max(0, min(i, kNESamp - 1));
    d_i -= i;
    /* ln K2 table: finite entries */
    const double tab = fm::exp_bounded((1.0 - d_i) * __ldg(P.k2 + i) + d_i * __ldg(P.k2 + i + 1));
    const double v = theta_e > kJnuMaxT ? 2.0 * theta_e * theta_e : tab;
    return theta_e < kThetaEMin ? 0.0 : v;
}
__device__ __forceinline__ double k2_eval(const GmParams &P, double theta_e) {
    return k2_eval_l(P, theta_e, fm::log_(fm::max_(theta_e, 1.0e-300)));
}

__device__ __forceinline__ double f_eval(const GmParams &P, double theta_e, double b_mag, double nu) {
    const double k = kJnuKFac * nu / (b_mag * theta_e * theta_e);
    if (k > kJnuMaxK)
        return 0.0;
    if (k < kJnuMinK) {
        const double x = cbrt(k);
        return x * (37.67503800178 + 2.240274341836 * x);
    }
    return interp_exp_table(P.f, log(k), P.jnu_l_min_k, P.jnu_d_l_k);
}

/* thermal synchrotron emissivity given sin(theta) and ln(theta_e) (reference synch, jnu_mixed.cpp:75-100).
 * Branch-free: the two "return 0" conditions of the reference are applied as selects at the end, the arithmetic
 * in between runs on guarded arguments. */
__device__ __forceinline__ double synch_sin_l(const GmParams &P, double nu, double n_e, double theta_e, double b,
                                              double sin_th, double l_theta) {
    const double k2 = k2_eval_l(P, theta_e, l_theta);
    const double nu_c = b * (kEE / (2.0 * kPi * kME * kCL));
    const double nu_s = (2.0 / 9.0) * nu_c * theta_e * theta_e * sin_th;
    const bool zero = (theta_e < kThetaEMin) || (nu > 1.0e12 * nu_s) || !(nu_s > 0.0);
    const double nu_s_g = zero ? 1.0 : nu_s, nu_g = zero ? 1.0 : nu, k2_g = zero ? 1.0 : k2;
    const double x = fm::div(nu_g, nu_s_g);
    const double xp = fm::cbrt_(x);
    const double xx = fm::sqrt_(x) + kJnuCst * fm::sqrt_(xp);
    const double f = xx * xx;
    const double j = fm::div(1.41421356237309504880 * kPi * kEE * kEE * n_e * nu_s_g, 3.0 * kCL * k2_g) * f *
                     fm::exp_(-xp);
    return zero ? 0.0 : j;
}
__device__ __forceinline__ double synch_sin(const GmParams &P, double nu, double n_e, double theta_e, double b,
                                            double sin_th) {
    return synch_sin_l(P, nu, n_e, theta_e, b, sin_th, fm::log_(fm::max_(theta_e, 1.0e-300)));
}

/* Klein-Nishina total cross-section / sigma_T (reference hotcross.cpp:144-152) */
__device__ __forceinline__ double hc_klein_nishina(double w) {
    if (w < 1.0e-3)
        return (1.0 - 2.0 * w);
    const double iw = 1.0 / w;
    const double t = 1.0 + 2.0 * w;
    return (3.0 / 4.0) * (2.0 * iw * iw + (0.5 * iw - (1.0 + w) * iw * iw * iw) * log(t) + (1.0 + w) / (t * t));
}

/* K2(1/theta) exp(1/theta): trapezoid rule on K_nu(x) e^x = int_0^inf exp(-x (cosh t - 1)) cosh(nu t) dt,
 * which converges geometrically for this analytic integrand.  The reference calls std::cyl_bessel_k
 * (hotcross.cpp:157); only the cold out-of-table fall-back needs it on the device. */
__device__ __noinline__ double k2_scaled(double x) {
    /* step: the integrand is ~ exp(-x t^2 / 2) cosh(2t), of width 1/sqrt(x) for large x; the trapezoid error is
     * ~ exp(-2 pi^2 / (h^2 x)), so h = 0.25 / sqrt(x) (error < 1e-130) once that is below the 0.125 that suffices
     * for the broad small-x integrand */
    const double h = x > 4.0 ? 0.25 / sqrt(x) : 0.125;
    double s = 0.5;
#pragma unroll 1
    for (int n = 1; n < 4000; ++n) {
        const double t = n * h;
        const double arg = x * (cosh(t) - 1.0);
        if (arg > 745.0)
            break;
        s += exp(-arg) * cosh(2.0 * t);
    }
    return s * h;
}

/* numeric integral over the electron distribution (reference total_compton_cross_num, hotcross.cpp:108-142);
 * cold path: only reached when (w, theta_e) leaves the table, SURVEY hard part H4 */
__device__ __noinline__ double hotcross_num(double w, double theta_e) {
    if (isnan(w))
        return 0.0;
    if (theta_e < kHcMinT && w < kHcMinW)
        return kSigmaThomson;
    if (theta_e < kHcMinT)
        return hc_klein_nishina(w) * kSigmaThomson;
    const double k2f = (theta_e > 1.0e-2) ? k2_scaled(1.0 / theta_e) : sqrt(kPi * theta_e / 2.0);
    double cross = 0.0;
    /* gamma_e - 1 ~ theta_e is the difference of two numbers of order one: the loop variable and gamma_e^2 - 1 are
     * formed with explicitly rounded multiplies and adds (no FMA contraction), as the reference's host code does,
     * otherwise the table entries at theta_e ~ 1e-4 drift by 5e-10 */
    const double d_gamma = __dmul_rn(theta_e, kHcDGammaE);
    const double gamma_0 = __dadd_rn(1.0, __dmul_rn(__dmul_rn(0.5, theta_e), kHcDGammaE));
    const double gamma_max = __dadd_rn(1.0, __dmul_rn(kHcMaxGamma, theta_e));
#pragma unroll 1
    for (double mu_e = -1.0 + 0.5 * kHcDMuE; mu_e < 1.0; mu_e += kHcDMuE) {
#pragma unroll 1
        for (double gamma_e = gamma_0; gamma_e < gamma_max; gamma_e = __dadd_rn(gamma_e, d_gamma)) {
            const double sq = sqrt(__dadd_rn(__dmul_rn(gamma_e, gamma_e), -1.));
            const double dnd = (gamma_e * sq / (theta_e * k2f)) * exp(-(gamma_e - 1.) / theta_e);
            const double v = sq / gamma_e;
            const double one_m_muv = __dadd_rn(1.0, -__dmul_rn(mu_e, v));
            const double we = w * gamma_e * one_m_muv;
            cross = __dadd_rn(cross, __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(theta_e, kHcDMuE), kHcDGammaE),
                                                         __dmul_rn(hc_klein_nishina(we), one_m_muv)),
                                               __dmul_rn(0.5, dnd)));
        }
    }
    return cross * kSigmaThomson;
}

/* the rare regimes of total_compton_cross_lkup that are not a table look-up (reference hotcross.cpp:86-93) */
__device__ __noinline__ double hotcross_cold(double w, double theta_e) {
    if (theta_e < kHcMinT)
        return hc_klein_nishina(w) * kSigmaThomson;
    return hotcross_num(w, theta_e);
}

/* reference total_compton_cross_lkup, hotcross.cpp:81-106, given l_w = ln(w) and l_theta = ln(theta_e).
 * The table look-up runs for every lane on clamped coordinates (its loads and exp10 cost the same for one lane
 * as for 32); the Thomson regime is a select, the out-of-table regimes a rarely taken call. */
__device__ __forceinline__ double hotcross_lkup_l(const GmParams &P, double w, double theta_e, double l_w,
                                                  double l_theta) {
    const bool thomson = w * theta_e < 1.0e-6;
    const bool in_table = !(w <= kHcMinW || w >= kHcMaxW || theta_e <= kHcMinT || theta_e >= kHcMaxT);
    const double kLog10E = 0.43429448190325182765;
    const double qw = (l_w * kLog10E - P.hc_l_min_w) * P.inv_hc_d_l_w;
    const double qt = (l_theta * kLog10E - P.hc_l_min_t) * P.inv_hc_d_l_t;
    /* the interpolant is used only for in_table points, where 0 < qw < n_w and 0 < qt < n_t: clamping the cell index
     * (two integer min/max) keeps the loads in bounds for all other lanes, whose result is replaced below */
    const int i = max(0, min((int)qw, kHcNW - 1)), j = max(0, min((int)qt, kHcNT - 1));
    const double d_i = qw - i, d_j = qt - j;
    const double *t = P.hotcross + i * (kHcNT + 1) + j;
    const double t00 = __ldg(t), t10 = __ldg(t + kHcNT + 1), t01 = __ldg(t + 1), t11 = __ldg(t + kHcNT + 2);
    const double l_cross = (1.0 - d_i) * (1.0 - d_j) * t00 + d_i * (1.0 - d_j) * t10 + (1.0 - d_i) * d_j * t01 +
                           d_i * d_j * t11;
    double sigma = fm::exp10_bounded(l_cross); /* log10 sigma table: finite entries */
    sigma = thomson ? kSigmaThomson : sigma;
    if (!thomson && !in_table)
        sigma = hotcross_cold(w, theta_e);
    return sigma;
}
__device__ __forceinline__ double hotcross_lkup(const GmParams &P, double w, double theta_e) {
    return hotcross_lkup_l(P, w, theta_e, fm::log_(fm::max_(w, 1.0e-300)), fm::log_(fm::max_(theta_e, 1.0e-300)));
}

/* fluid-frame photon energy (units of m_e c^2) and cosine of the angle between k and b
 * (reference bk_angle radiation.cpp:59-87, fluid_nu :89-101) */
__device__ __forceinline__ void fluid_frame(const GmParams &P, const double k[4], const Fluid &f, double &e_fluid,
                                            double &mu) {
    const double ku = k[0] * f.u_cov[0] + k[1] * f.u_cov[1] + k[2] * f.u_cov[2] + k[3] * f.u_cov[3];
    e_fluid = -ku;
    const double kb = k[0] * f.b_cov[0] + k[1] * f.b_cov[1] + k[2] * f.b_cov[2] + k[3] * f.b_cov[3];
    const bool no_b = (f.b == 0.0);
    const double den = fabs(ku) * (no_b ? 1.0 : f.b) * P.inv_b_unit;
    mu = fm::div(kb, den);
    mu = fm::min_(fm::max_(mu, -1.0), 1.0);
    mu = no_b ? 0.0 : mu; /* theta = pi/2 */
}

/* invariant scattering opacity nu * sigma_hot * n_e (reference alpha_inv_scatt / kappa_es,
 * radiation.cpp:103-107,142-146; the m_p factors cancel) */
__device__ __forceinline__ double alpha_inv_scatt_l(const GmParams &P, double nu, double theta_e, double n_e,
                                                    double l_nu, double l_theta) {
    const double e_g = nu * (kHPL / (kME * kCL * kCL));
    const double kLnHOverMc2 = -46.263250426746548; /* ln(h / (m_e c^2)) in cgs: ln(e_g) = ln(nu) + this */
    return nu * hotcross_lkup_l(P, e_g, theta_e, l_nu + kLnHOverMc2, l_theta) * n_e;
}
__device__ __forceinline__ double alpha_inv_scatt(const GmParams &P, double nu, double theta_e, double n_e) {
    return alpha_inv_scatt_l(P, nu, theta_e, n_e, fm::log_(fm::max_(nu, 1.0e-300)), fm::log_(fm::max_(theta_e, 1.0e-300)));
}

/* reference b_nu_inv, radiation.cpp:120-128 (series below x = 1e-3), as a select between the two forms */
__device__ __forceinline__ double b_nu_inv(double nu, double theta_e) {
    const double x = fm::div(nu * (kHPL / (kME * kCL * kCL)), theta_e);
    const double c = 2.0 * kHPL / (kCL * kCL);
    double den = x * (1.0 / 24.0) * (24.0 + x * (12.0 + x * (4.0 + x)));
    /* x >= 1e-3 means h nu >= 1e-3 k T_e: only up-scattered X-ray photons get here, so the exponential is a
     * rarely taken branch instead of a select that every lane pays for */
    if (!(x < 1.0e-3))
        den = fm::exp_(fm::min_(x, 700.0)) - 1.0;
    return fm::div(c, den);
}

/* invariant absorption opacity by Kirchhoff's law (reference alpha_inv_abs, radiation.cpp:109-118):
 * (j / nu^2) / (B + 1e-100) evaluated as j / (nu^2 (B + 1e-100)): one division, no denormal intermediate */
__device__ __forceinline__ double alpha_inv_abs_sin_l(const GmParams &P, double nu, double theta_e, double n_e,
                                                      double b, double sin_th, double l_theta) {
    const double j = synch_sin_l(P, nu, n_e, theta_e, b, sin_th, l_theta);
    return fm::div(j, nu * nu * (b_nu_inv(nu, theta_e) + 1.0e-100));
}
__device__ __forceinline__ double alpha_inv_abs_sin(const GmParams &P, double nu, double theta_e, double n_e,
                                                    double b, double sin_th) {
    return alpha_inv_abs_sin_l(P, nu, theta_e, n_e, b, sin_th, fm::log_(fm::max_(theta_e, 1.0e-300)));
}

/* reference bias_func, harm_model.cpp:1391-1404, with the generation's frozen statistics: `bias_den` is
 * bias_norm * max_tau_scatt * (n_scatt / (n_recorded + 1) + 2), the same for every photon of a generation */
__device__ __forceinline__ double bias_func_den(double theta_e, double w, double bias_den) {
    const double mx = w * (0.5 / kWeightMin);
    double bias = fm::div(100.0 * theta_e * theta_e, bias_den);
    bias = fm::max_(bias, kTpOverTe);
    bias = fm::min_(bias, mx);
    return bias * (1.0 / kTpOverTe);
}
__device__ __forceinline__ double bias_func(const GmParams &P, const GmBiasStats &s, double theta_e, double w) {
    (void)P;
    return bias_func_den(theta_e, w, s.bias_den);
}

/* both opacities at once for a photon with wave-vector k in fluid f */
__device__ __forceinline__ void opacities(const GmParams &P, const double k[4], const Fluid &f, double &nu,
                                          double &alpha_scatt, double &alpha_abs) {
    double e_fluid, mu;
    fluid_frame(P, k, f, e_fluid, mu);
    nu = e_fluid * (kME * kCL * kCL / kHPL);
    if (nu < 0.0 || isnan(nu)) {
        alpha_scatt = 0.0;
        alpha_abs = 0.0;
        return;
    }
    const double l_nu = fm::log_(nu), l_theta = fm::log_(f.theta_e);
    alpha_scatt = alpha_inv_scatt_l(P, nu, f.theta_e, f.n_e, l_nu, l_theta);
    alpha_abs = alpha_inv_abs_sin_l(P, nu, f.theta_e, f.n_e, f.b, fm::sqrt_(1.0 - mu * mu), l_theta);
}

} /* namespace gm */
