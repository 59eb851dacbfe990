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
    if (sqrt_(0.0) != 0.0) m = 1e300;
    std::printf("sqrt %d %.3f\n", N, m);
    m = 0; for (int i = 0; i < N; ++i) { double x = (i & 3) == 0 ? U(-708, 709) : U(-40, 40); m = std::fmax(m, ulp_err(exp_(x), expl(x))); }
    if (exp_(-800.0) != 0.0 || !std::isinf(exp_(720.0)) || exp_(0.0) != 1.0) m = 1e300;
    std::printf("exp %d %.3f\n", N, m);
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
    return 0;
}
