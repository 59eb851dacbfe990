/*
 * gm_math.cuh -- branch-free FP64 elementary functions for the hot transport loop (sm_100a).
 *
 * Why not the CUDA math library here: every libm call the compiler inlines carries a convergence barrier
 * (BSSY/BSYNC) around its rare special-case path (huge arguments, denormals, NaN), which makes each call its
 * own scheduling region.  ncu on the transport kernel (profiles/r1_transport_*.txt) showed the Horner chains of
 * exp / sincos / log / division executing back to back, one dependent DFMA every ~7 cycles, with only two warps
 * per scheduler to hide them: 30% of all stall samples were fixed-latency waits.  The functions below
 *   - have no branches (special cases are folded into selects or excluded by the documented argument domain),
 *   - evaluate their polynomials by Estrin's scheme (dependency depth ~log2(degree) instead of degree), and
 *   - are plain inline code, so that ptxas interleaves independent evaluations (e.g. exp(x1) with sincospi(2 x2),
 *     or the three logarithms of the interaction step) in one basic block.
 * Accuracy: <= 1.5 ulp on the stated domains (checked on the host against glibc by tests/test_math_host.py,
 * which compiles this very header with g++); the parity bars of the path are 1e-10 .. 1e-12 relative.
 *
 * Reference call sites that use these (cuda_grmonty/harm_model.cpp): get_connection :1438-1445 (exp, sin, cos),
 * step_size :1620-1630 and push_photon :1257-1267 (divisions), get_fluid_params :638-668 (sqrt, division),
 * radiation.cpp:103-146 and jnu_mixed.cpp:75-168 (log, exp, pow, sqrt), hotcross.cpp:94-105 (log10, pow(10,.)).
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define GM_HD __host__ __device__ __forceinline__
#else
#define GM_HD inline
#endif

namespace gm {
namespace fm {

GM_HD double from_bits(uint64_t u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d;
    std::memcpy(&d, &u, 8);
    return d;
#endif
}
GM_HD uint64_t to_bits(double d) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    std::memcpy(&u, &d, 8);
    return u;
#endif
}
GM_HD int hi_word(double d) { return (int)(to_bits(d) >> 32); }
GM_HD double with_hi(double d, int hi) {
    return from_bits((to_bits(d) & 0xffffffffull) | ((uint64_t)(uint32_t)hi << 32));
}

/* ~2^-23 reciprocal / reciprocal square root seeds (MUFU.RCP64H / MUFU.RSQ64H) */
GM_HD double rcp_seed(double b) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    return y;
#else
    return from_bits(to_bits(1.0 / b) & 0xfffffffe00000000ull); /* host model of the seed: 21 bits */
#endif
}
GM_HD double rsqrt_seed(double b) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    return y;
#else
    return from_bits(to_bits(1.0 / std::sqrt(b)) & 0xfffffffe00000000ull);
#endif
}

/* 1/b for normal b (|b| in [2^-1000, 2^1000]); b = 0, inf, NaN, denormal are outside the domain */
GM_HD double rcp(double b) {
    const double y0 = rcp_seed(b);
    const double e0 = fma(-b, y0, 1.0);
    const double y1 = fma(y0, fma(e0, e0, e0), y0); /* (1 + e + e^2): error e^3 ~ 2^-69 */
    const double e1 = fma(-b, y1, 1.0);
    return fma(y1, e1, y1);
}

/* a/b for normal b and a quotient in the normal range; error <= 1 ulp */
GM_HD double div(double a, double b) {
    const double y0 = rcp_seed(b);
    const double e0 = fma(-b, y0, 1.0);
    const double y1 = fma(y0, fma(e0, e0, e0), y0);
    const double q = a * y1;
    const double r = fma(-b, q, a);
    return fma(r, y1, q);
}

/* sqrt(x): x normal or zero; 0 -> 0, negative -> NaN, NaN -> NaN (all by selects) */
GM_HD double sqrt_(double x) {
    const double y0 = rsqrt_seed(x);
    double g = x * y0, h = 0.5 * y0;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, x);
    g = fma(d, h, g);
    const double special = x < 0.0 ? from_bits(0x7ff8000000000000ull) : x; /* 0 -> 0, NaN -> NaN */
    return x > 0.0 ? g : special;                                           /* x == 0: the seed is inf, g is NaN */
}

/* exp(x), any finite x: 0 below -708 (denormal results are flushed), +inf above 709.78 */
GM_HD double exp_core(double r, double n) {
    /* exp(r) on |r| <= 0.3466, degree-11 minimax-like Taylor (|r|^12/12! < 7e-15 relative to 1: the last
     * coefficients are the classic fdlibm-style 1/k!), Estrin */
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p23 = fma(r, 1.6666666666666666e-01, 0.5);
    const double p45 = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    const double p67 = fma(r, 1.9841269841269841e-04, 1.3888888888888889e-03);
    const double p89 = fma(r, 2.7557319223985893e-06, 2.4801587301587302e-05);
    const double pab = fma(r, 2.5052108385441720e-08, 2.7557319223985888e-07);
    const double pcd = fma(r, 1.6059043836821613e-10, 2.0876756987868100e-09);
    const double q0 = fma(r2, p23, r);
    const double q1 = fma(r2, p67, p45);
    const double q2 = fma(r2, pab, p89);
    const double s0 = fma(r4, q1, q0);
    const double s1 = fma(r4, pcd, q2);
    const double p = 1.0 + fma(r8, s1, s0); /* the O(r) part is summed first: one rounding at the size of 1 */
    /* scale by 2^n, n integer in [-1021, 1023] */
    const int ni = (int)n;
    return with_hi(p, hi_word(p) + (ni << 20));
}

GM_HD double exp_(double x) {
    const double xc = fmin(fmax(x, -708.0), 709.0);
    const double t = fma(xc, 1.4426950408889634, 6755399441055744.0);
    const double n = t - 6755399441055744.0;
    double r = fma(n, -6.93147180369123816490e-01, xc);
    r = fma(n, -1.90821492927058770002e-10, r);
    double v = exp_core(r, n);
    v = x < -708.0 ? 0.0 : v;
    v = x > 709.78 ? from_bits(0x7ff0000000000000ull) : v;
    return x != x ? x : v; /* NaN in, NaN out (the clamps above would hide it) */
}

/* 10^x */
GM_HD double exp10_(double x) {
    const double xc = fmin(fmax(x, -307.0), 308.0);
    const double t = fma(xc, 3.3219280948873622, 6755399441055744.0);
    const double n = t - 6755399441055744.0;
    /* x - n log10(2) in two parts, then to the natural base */
    double r = fma(n, -3.01029995663611771306e-01, xc);
    r = fma(n, -3.69423907715893078616e-13, r);
    const double rh = r * 2.30258509299404590109e+00;
    const double rl = fma(r, 2.30258509299404590109e+00, -rh);
    const double rr = rh + fma(r, -2.17071551782250736e-16, rl);
    double v = exp_core(rr, n);
    v = x < -307.0 ? 0.0 : v;
    v = x > 308.25 ? from_bits(0x7ff0000000000000ull) : v;
    return x != x ? x : v;
}

/* ln(x) for normal x > 0 */
GM_HD double log_(double x) {
    int hi = hi_word(x);
    int e = (hi >> 20) - 1023;
    /* mantissa in [sqrt(1/2), sqrt(2)) */
    const int big = ((hi & 0x000fffff) >= 0x6a09f) ? 1 : 0;
    e += big;
    const double m = with_hi(x, (hi & 0x000fffff) | ((1023 - big) << 20));
    const double f = m - 1.0;
    const double d = m + 1.0;
    const double y = rcp(d);
    const double s = f * y;
    const double sl = fma(-s, d, f) * y; /* s + sl = f / d to ~2^-100 */
    const double v = s * s;
    /* 2 atanh(s) = 2 s + s^3 (2/3 + 2/5 v + ... + 2/21 v^9), Estrin in v; |s| <= 0.1716 */
    const double v2 = v * v, v4 = v2 * v2, v8 = v4 * v4;
    const double a01 = fma(v, 4.0000000000000000e-01, 6.6666666666666663e-01);
    const double a23 = fma(v, 2.2222222222222221e-01, 2.8571428571428570e-01);
    const double a45 = fma(v, 1.5384615384615385e-01, 1.8181818181818182e-01);
    const double a67 = fma(v, 1.1764705882352941e-01, 1.3333333333333333e-01);
    const double a89 = fma(v, 9.5238095238095233e-02, 1.0526315789473684e-01);
    const double b0 = fma(v2, a23, a01);
    const double b1 = fma(v2, a67, a45);
    const double c0 = fma(v4, b1, b0);
    const double poly = fma(v8, a89, c0);
    const double ed = (double)e;
    const double hi_part = fma(ed, 6.93147180369123816490e-01, 2.0 * s);
    const double lo_part = fma(ed, 1.90821492927058770002e-10, fma(s * v, poly, 2.0 * sl));
    return hi_part + lo_part;
}

/* sin and cos of pi*t for |t| < 2^30; exact at multiples of 1/2 */
GM_HD void sincospi_(double t, double *sp, double *cp) {
    const double q = rint(t + t);           /* nearest half-integer count */
    const double r = fma(q, -0.5, t);       /* |r| <= 1/4, exact */
    const int qi = (int)q;
    const double x = r * 3.14159265358979311600e+00;
    const double xl = fma(r, 3.14159265358979311600e+00, -x);
    const double xr = x + fma(r, 1.22464679914735317723e-16, xl); /* pi r, |.| <= pi/4 */
    const double z = xr * xr, z2 = z * z, z4 = z2 * z2;
    /* sin(x) = x + x z (S1 + S2 z + ... + S6 z^5), cos(x) = 1 - z/2 + z^2 (C1 + ... + C6 z^5): fdlibm kernels */
    const double s12 = fma(z, 8.33333333332248946124e-03, -1.66666666666666324348e-01);
    const double s34 = fma(z, 2.75573137070700676789e-06, -1.98412698298579493134e-04);
    const double s56 = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    const double sp_ = fma(z4, s56, fma(z2, s34, s12));
    const double sn = fma(xr * z, sp_, xr);
    const double c12 = fma(z, -1.38888888888741095749e-03, 4.16666666666666019037e-02);
    const double c34 = fma(z, -2.75573143513906633035e-07, 2.48015872894767294178e-05);
    const double c56 = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    const double cp_ = fma(z4, c56, fma(z2, c34, c12));
    const double cs = fma(z2, cp_, fma(z, -0.5, 1.0));
    /* quadrant */
    const bool swap = (qi & 1) != 0;
    double so = swap ? cs : sn, co = swap ? sn : cs;
    so = (qi & 2) ? -so : so;
    co = ((qi + 1) & 2) ? -co : co;
    *sp = so;
    *cp = co;
}

/* sin and cos of x for |x| < 1e5 (three-part Cody-Waite reduction by pi/2) */
GM_HD void sincos_(double x, double *sp, double *cp) {
    const double t = fma(x, 6.36619772367581382433e-01, 6755399441055744.0);
    const double q = t - 6755399441055744.0;
    const int qi = (int)q;
    double r = fma(q, -1.57079632673412561417e+00, x);
    r = fma(q, -6.07710050630396597660e-11, r);
    const double rl = q * 2.02226624879595063154e-21;
    const double xr = r - rl;
    const double z = xr * xr, z2 = z * z, z4 = z2 * z2;
    const double s12 = fma(z, 8.33333333332248946124e-03, -1.66666666666666324348e-01);
    const double s34 = fma(z, 2.75573137070700676789e-06, -1.98412698298579493134e-04);
    const double s56 = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    const double sp_ = fma(z4, s56, fma(z2, s34, s12));
    const double sn = fma(xr * z, sp_, xr);
    const double c12 = fma(z, -1.38888888888741095749e-03, 4.16666666666666019037e-02);
    const double c34 = fma(z, -2.75573143513906633035e-07, 2.48015872894767294178e-05);
    const double c56 = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    const double cp_ = fma(z4, c56, fma(z2, c34, c12));
    const double cs = fma(z2, cp_, fma(z, -0.5, 1.0));
    const bool swap = (qi & 1) != 0;
    double so = swap ? cs : sn, co = swap ? sn : cs;
    so = (qi & 2) ? -so : so;
    co = ((qi + 1) & 2) ? -co : co;
    *sp = so;
    *cp = co;
}

/* x^(1/3) for normal x > 0: exponent split + degree-4 seed on [1,8) + two Halley steps */
GM_HD double cbrt_(double x) {
    const int hi = hi_word(x);
    const int e = (hi >> 20) - 1023;
    /* e = 3 q + rem, rem in {0,1,2} (floor division for negative e) */
    const int q = (e + 3072) / 3 - 1024;
    const int rem = e - 3 * q;
    const double m = with_hi(x, (hi & 0x000fffff) | ((1023 + rem) << 20)); /* [1, 8) */
    /* seed: least-squares fit of m^(1/3) on [1,8) (relative error), degree 4: 5.4e-3 */
    double y = fma(m, fma(m, fma(m, fma(m, -3.47020774e-04, 7.96603547e-03), -7.33267142e-02), 4.22924391e-01),
                   6.48113736e-01);
    /* Halley: y <- y (y^3 + 2 m) / (2 y^3 + m), cubic convergence: 5e-3 -> 1e-7 -> 1e-21 */
    double y3 = y * y * y;
    y = y * div(y3 + 2.0 * m, fma(2.0, y3, m));
    y3 = y * y * y;
    y = y * div(y3 + 2.0 * m, fma(2.0, y3, m));
    /* final Newton correction in extended precision: y <- y - (y^3 - m) / (3 y^2) */
    const double y2 = y * y;
    const double y2l = fma(y, y, -y2);
    const double y3h = y2 * y;
    const double y3l = fma(y2, y, -y3h) + y2l * y;
    const double resid = (m - y3h) - y3l;
    y = fma(resid, rcp(3.0 * y2), y);
    return with_hi(y, hi_word(y) + (q << 20));
}

} /* namespace fm */
} /* namespace gm */
