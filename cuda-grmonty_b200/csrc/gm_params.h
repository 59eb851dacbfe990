/*
 * gm_params.h -- constants and the parameter block shared by host and device code of the B200 path.
 *
 * Constant VALUES are those of the reference (cuda_grmonty/consts.hpp:14-171); derived logarithms are
 * computed once on the host with the same std::log / std::log10 expressions and passed in GmParams so the
 * device never re-derives them.
 */
#pragma once
#include <cstdint>

namespace gm {

/* reference consts.hpp:21-55 */
constexpr double kEps = 1.0e-40;
constexpr int kNESamp = 200;
constexpr int kNEBins = 200;
constexpr int kNThBins = 6;
constexpr int kSpecFields = 13;
constexpr double kNuMin = 1.0e9;
constexpr double kNuMax = 1.0e16;
constexpr double kThetaEMin = 0.3;
constexpr double kTpOverTe = 3.0;
constexpr double kWeightMin = 1.0e31;
constexpr double kRoulette = 1.0e4;
constexpr double kRMax = 100.0;
constexpr double kStepEps = 0.04;
constexpr double kETol = 1.0e-3;
constexpr int kMaxIter = 2;
constexpr int kMaxNStep = 1280000;
constexpr int kMaxHalvings = 7; /* reference harm_model.cpp:1279 */
/* reference consts.hpp:58-83 (CGS) */
constexpr double kPi = 3.14159265358979323846;
constexpr double kEE = 4.80320680e-10;
constexpr double kCL = 2.99792458e10;
constexpr double kME = 9.1093826e-28;
constexpr double kMP = 1.67262171e-24;
constexpr double kHPL = 6.6260693e-27;
constexpr double kHBAR = kHPL / (2. * kPi);
constexpr double kSigmaThomson = 0.665245873e-24;
constexpr double kGNewt = 6.6742e-8;
constexpr double kMSun = 1.989e33;
constexpr double kLSun = 3.827e33;
constexpr double kMBH = 4.0e6 * kMSun;
/* reference consts.hpp:86-90 */
constexpr int kNint = 20000;
constexpr double kBthsqMin = 1.0e-4;
constexpr double kBthsqMax = 1.0e8;
/* reference consts.hpp:94-112 (hotcross) */
constexpr double kHcMinW = 1.0e-12;
constexpr double kHcMaxW = 1.0e6;
constexpr double kHcMinT = 1.0e-4;
constexpr double kHcMaxT = 1.0e4;
constexpr int kHcNW = 220;
constexpr int kHcNT = 80;
constexpr double kHcMaxGamma = 12.0;
constexpr double kHcDMuE = 0.05;
constexpr double kHcDGammaE = 0.05;
/* reference consts.hpp:118-139 (jnu) */
constexpr double kJnuMinK = 0.002;
constexpr double kJnuMaxK = 1.0e7;
constexpr double kJnuMaxT = 1.0e2;
constexpr double kJnuCst = 1.88774862536;
constexpr double kJnuKFac = 9 * kPi * kME * kCL / kEE;
constexpr double kJcst = 1.41421356237309504880 * kEE * kEE * kEE / (27.0 * kME * kCL * kCL);
/* reference consts.hpp:151-158 */
constexpr double kSpecDLE = 0.25;

/* spectrum field indices = member order of harm::Spectrum (reference harm_data.hpp:129-143) */
enum SpecField {
    F_DN_DLE = 0, F_DE_DLE, F_NPH, F_NSCATT, F_X1I_AV, F_X2I_SQ, F_X3F_SQ, F_TAU_ABS, F_TAU_SCATT, F_NE_0,
    F_THETA_E_0, F_B_0, F_E_0
};

/* Parameter block passed by value to every kernel (lives in the constant bank). */
struct GmParams {
    int n0, n1;
    double x_start1, x_start2, dx1, dx2, x_stop1, x_stop2;
    double a, h_slope, r_0;
    double b_unit, theta_e_unit, n_e_unit;
    double photon_n, bias_norm, d_tau_k, x1_min, x1_max;
    uint32_t seed_lo, seed_hi;
    /* derived logs (host std::log of the reference expressions) */
    double l_nu_min, n_l_n, d_l_nu;   /* consts.hpp:33-36 */
    double l_b_min, d_l_b;            /* consts.hpp:89-90 */
    double hc_l_min_w, hc_l_min_t, hc_d_l_w, hc_d_l_t; /* consts.hpp:108-112 */
    double jnu_l_min_k, jnu_d_l_k, jnu_l_min_t, jnu_d_l_t; /* consts.hpp:125-137 */
    double spec_l_e_0;                /* consts.hpp:156 */
    double nz_max;                    /* photon_n * ln(nu_max/nu_min), harm_model.cpp:1384 */
    /* reciprocals of run constants the hot loop divides by (x * (1/c) instead of x / c: <= 1 ulp apart) */
    double inv_dx1, inv_dx2, inv_b_unit, inv_hc_d_l_w, inv_hc_d_l_t, inv_jnu_d_l_t;
    /* device pointers */
    const double *grid;     /* interleaved primitives [n0][n1][8]: k_rho u u_1 u_2 u_3 b_1 b_2 b_3 */
    const double *geom_det; /* [n0][n1] */
    const double *hotcross; /* [221][81] */
    const double *f, *k2, *weight; /* [201] */
    const double *nint, *dndlnu_max; /* [20001] */
};

/* frozen scattering-bias statistics of one generation (reference harm_model.cpp:1391-1394 reads them live) */
struct GmBiasStats {
    double max_tau_scatt;
    double n_scatt;
    double n_recorded;
    /* bias_norm * max_tau_scatt * (n_scatt / (n_recorded + 1) + 2): the denominator of bias_func, the same for
     * every photon of the generation (computed on the host with the reference's operation order) */
    double bias_den;
};

inline GmBiasStats make_bias_stats(double bias_norm, double max_tau_scatt, double n_scatt, double n_recorded) {
    GmBiasStats s;
    s.max_tau_scatt = max_tau_scatt;
    s.n_scatt = n_scatt;
    s.n_recorded = n_recorded;
    const double avg_num_scatt = n_scatt / (1.0 * n_recorded + 1.0);
    s.bias_den = bias_norm * max_tau_scatt * (avg_num_scatt + 2.0);
    return s;
}

} /* namespace gm */
