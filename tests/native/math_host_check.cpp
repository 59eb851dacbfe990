// Host-side accuracy check of cuda-grmonty_b200/csrc/gm_math.cuh (compiled by tests/test_math_host.py with g++;
// the header is plain C++ when __CUDACC__ is not defined, with the MUFU seeds modelled as 21-bit truncations).
// Prints one line per function: name, samples, max error in ulp against the long-double libm value.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../cuda-grmonty_b200/csrc/gm_math.cuh"

static double ulp_err(double got, long double want) {
    if (want == 0.0L) return got == 0.0 ? 0.0 : 1e300;
    const double w = (double)want;
    const double u = std::ldexp(1.0, std::ilogb(w) - 52);
    return (double)(std::fabs((long double)got - want) / u);
}

int main() {
    std::mt19937_64 gen(12345);
    auto U = [&](double a, double b) { return a + (b - a) * std::generate_canonical<double, 53>(gen); };
    auto LU = [&](double la, double lb) { return std::pow(10.0, U(la, lb)); };
    const int N = 2000000;
    double m;
    using namespace gm::fm;
    m = 0; for (int i = 0; i < N; ++i) { double b = LU(-290, 290) * (i & 1 ? -1 : 1); m = std::fmax(m, ulp_err(rcp(b), 1.0L / b)); }
    std::printf("rcp %d %.3f\n", N, m);
    m = 0; for (int i = 0; i < N; ++i) { double a = LU(-140, 140), b = LU(-140, 140) * (i & 1 ? -1 : 1); m = std::fmax(m, ulp_err(div(a, b), (long double)a / b)); }
    std::printf("div %d %.3f\n", N, m);
    m = 0; for (int i = 0; i < N; ++i) { double x = LU(-290, 290); m = std::fmax(m, ulp_err(sqrt_(x), sqrtl(x))); }
    if (sqrt_(0.0) != 0.0 || !std::isnan(sqrt_(-1.0)) || !std::isnan(sqrt_(std::nan(""))) || std::signbit(sqrt_(-0.0)) != true)
        m = 1e300; /* 0 -> 0, -0 -> -0, negative and NaN -> NaN through the seed alone */
    std::printf("sqrt %d %.3f\n", N, m);
    m = 0; for (int i = 0; i < N; ++i) { double x = (i & 3) == 0 ? U(-708, 709) : U(-40, 40); m = std::fmax(m, ulp_err(exp_(x), expl(x))); }
    if (exp_(-800.0) != 0.0 || !std::isinf(exp_(720.0)) || exp_(0.0) != 1.0) m = 1e300;
    if (!std::isnan(exp_(std::nan(""))) || exp_(709.7) < 1.6e308 || !std::isinf(exp_(709.79)) || exp_(-708.0) <= 0.0) m = 1e300;
    std::printf("exp %d %.3f\n", N, m);
    /* the select-free variants agree bit for bit with the full ones on their domain */
    m = 0; for (int i = 0; i < N; ++i) { double x = U(-708, 709); if (exp_bounded(x) != exp_(x)) m = 1e300; }
    std::printf("exp_bounded %d %.3f\n", N, m);
    m = 0; for (int i = 0; i < N; ++i) { double x = U(-307, 308); if (exp10_bounded(x) != exp10_(x)) m = 1e300; }
    std::printf("exp10_bounded %d %.3f\n", N, m);
    /* min_/max_: fmin/fmax semantics for a NaN in the first argument and for ordinary values */
    m = 0;
    for (int i = 0; i < N; ++i) {
        double a = U(-10, 10), b = U(-10, 10);
        if (min_(a, b) != std::fmin(a, b) || max_(a, b) != std::fmax(a, b)) m = 1e300;
    }
    if (min_(std::nan(""), 1.0) != 1.0 || max_(std::nan(""), -1.0) != -1.0) m = 1e300;
    std::printf("minmax %d %.3f\n", N, m);
    m = 0; for (int i = 0; i < N; ++i) { double x = (i & 3) == 0 ? U(-307, 308) : U(-40, 20); m = std::fmax(m, ulp_err(exp10_(x), powl(10.0L, x))); }
    if (exp10_(-400.0) != 0.0 || !std::isinf(exp10_(310.0))) m = 1e300;
    std::printf("exp10 %d %.3f\n", N, m);
    m = 0; for (int i = 0; i < N; ++i) { double x = (i & 3) == 0 ? LU(-300, 300) : ((i & 3) == 1 ? U(0.5, 2.0) : LU(-20, 20)); m = std::fmax(m, ulp_err(log_(x), logl(x))); }
    if (log_(1.0) != 0.0) m = 1e300;
    std::printf("log %d %.3f\n", N, m);
    double ms = 0, mc = 0;
    for (int i = 0; i < N; ++i) {
        double t = U(-4.0, 6.0), s, c;
        sincospi_(t, &s, &c);
        const long double a = 3.14159265358979323846264338327950288L * (long double)t;
        // compare in absolute terms scaled by 2^-53 (results near zero are limited by the argument rounding)
        ms = std::fmax(ms, (double)(std::fabs((long double)s - sinl(a)) / 1.1102230246251565e-16L));
        mc = std::fmax(mc, (double)(std::fabs((long double)c - cosl(a)) / 1.1102230246251565e-16L));
    }
    { double s, c; sincospi_(0.5, &s, &c); if (s != 1.0 || c != 0.0) ms = 1e300; sincospi_(1.0, &s, &c); if (s != 0.0 || c != -1.0) ms = 1e300; }
    std::printf("sincospi %d %.3f %.3f\n", N, ms, mc);
    ms = mc = 0;
    for (int i = 0; i < N; ++i) {
        double x = (i & 1) ? U(-1.0, 4.0) : U(-1000.0, 1000.0), s, c;
        sincos_(x, &s, &c);
        ms = std::fmax(ms, (double)(std::fabs((long double)s - sinl((long double)x)) / 1.1102230246251565e-16L));
        mc = std::fmax(mc, (double)(std::fabs((long double)c - cosl((long double)x)) / 1.1102230246251565e-16L));
    }
    std::printf("sincos %d %.3f %.3f\n", N, ms, mc);
    m = 0; for (int i = 0; i < N; ++i) { double x = (i & 1) ? LU(-300, 300) : LU(-3, 13); m = std::fmax(m, ulp_err(cbrt_(x), cbrtl(x))); }
    std::printf("cbrt %d %.3f\n", N, m);
    /* The two sums of quotients that the kernels evaluate over one common denominator (gm_geometry.cuh: err_norm,
     * step_size) against the quotient-by-quotient forms in long double, including the eps-dominated corners. */
    {
        const double eps = 1.0e-40;
        double me = 0, msz = 0;
        for (int i = 0; i < N; ++i) {
            double kn[4], kp[4];
            for (int c = 0; c < 4; ++c) {
                const int mode = (int)(U(0, 1) * 8);
                kn[c] = mode == 0 ? 0.0 : (mode == 1 ? LU(-45, -35) : LU(-6, 4)) * (U(0, 1) < 0.5 ? -1 : 1);
                kp[c] = kn[c] * (1.0 + U(-1, 1) * LU(-8, -1)) + (mode == 0 ? LU(-60, -38) : 0.0);
            }
            const double d0 = std::fabs(kn[0] + eps), d1 = std::fabs(kn[1] + eps), d2 = std::fabs(kn[2] + eps), d3 = std::fabs(kn[3] + eps);
            const double n0 = std::fabs(kp[0] - kn[0]), n1 = std::fabs(kp[1] - kn[1]), n2 = std::fabs(kp[2] - kn[2]), n3 = std::fabs(kp[3] - kn[3]);
            const double d01 = d0 * d1, d23 = d2 * d3;
            const double num = std::fma(std::fma(n0, d1, n1 * d0), d23, std::fma(n2, d3, n3 * d2) * d01);
            const double got = div(num, d01 * d23);
            /* the inputs' own rounding (kp - kn, kn + eps) is common to both forms and excluded: compare in ulp */
            const long double ref =
                ((long double)n0 / d0 + (long double)n1 / d1) + ((long double)n2 / d2 + (long double)n3 / d3);
            if (ref > 0) me = std::fmax(me, ulp_err(got, ref));
            else if (got != 0.0) me = 1e300;
            if (!std::isfinite(got)) me = 1e300;
        }
        std::printf("errnorm_onediv %d %.3f\n", N, me);
        for (int i = 0; i < N; ++i) {
            double b[3], a[3];
            for (int c = 0; c < 3; ++c) {
                const int mode = (int)(U(0, 1) * 6);
                b[c] = (mode == 0 ? 0.0 : LU(-8, 4)) + eps;
                a[c] = (c == 2) ? 0.04 : (mode == 1 ? 0.0 : (mode == 2 ? LU(-30, -10) : LU(-4, 0)));
            }
            const double d1 = a[0] + eps * b[0], d2 = a[1] + eps * b[1], d3 = a[2] + eps * b[2];
            const long double ref = 1.0L / ((long double)b[0] / d1 + (long double)b[1] / d2 + (long double)b[2] / d3);
            const double d23 = d2 * d3;
            const double got = div(d1 * d23, std::fma(b[0], d23, d1 * std::fma(b[1], d3, b[2] * d2)));
            msz = std::fmax(msz, ulp_err(got, ref));
            if (!std::isfinite(got) || got <= 0.0) msz = 1e300;
        }
        std::printf("stepsize_onediv %d %.3f\n", N, msz);
    }
    return 0;
}
