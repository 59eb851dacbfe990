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
/* min/max by one compare and a select: fmin()/fmax() cost ~8 instructions each for their IEEE NaN rules.  A NaN in
 * the FIRST argument yields the second (as fmin/fmax do); callers put the constant bound second. */
GM_HD double min_(double a, double b) { return a < b ? a : b; }
GM_HD double max_(double a, double b) { return a > b ? a : b; }
GM_HD int hi_word(double d) { return (int)(to_bits(d) >> 32); }
GM_HD double with_hi(double d, int hi) {
    return from_bits((to_bits(d) & 0xffffffffull) | ((uint64_t)(uint32_t)hi << 32));
}

/* Polynomial coefficients and reduction constants.  On the device they live in constant memory: a DFMA takes a
 * constant-bank operand directly, while a 64-bit literal costs two extra instructions (IMAD.MOV / UMOV pairs were
 * 15 % of all executed instructions of the transport kernel with literals).  The host build uses the same table. */
#define GM_MATH_CONSTANTS(X) \
    X(E3, 1.6666666666666666e-01) \
    X(E5, 8.3333333333333332e-03) \
    X(E4, 4.1666666666666664e-02) \
    X(E7, 1.9841269841269841e-04) \
    X(E6, 1.3888888888888889e-03) \
    X(E9, 2.7557319223985893e-06) \
    X(E8, 2.4801587301587302e-05) \
    X(E11, 2.5052108385441720e-08) \
    X(E10, 2.7557319223985888e-07) \
    X(E13, 1.6059043836821613e-10) \
    X(E12, 2.0876756987868100e-09) \
    X(LOG2E, 1.4426950408889634) \
    X(MAGIC, 6755399441055744.0) \
    X(LN2_HI, 6.93147180369123816490e-01) \
    X(LN2_LO, 1.90821492927058770002e-10) \
    X(LOG2_10, 3.3219280948873622) \
    X(LG2_HI, 3.01029995663611771306e-01) \
    X(LG2_LO, 3.69423907715893078616e-13) \
    X(LN10_HI, 2.30258509299404590109e+00) \
    X(LN10_LO_NEG, 2.17071551782250736e-16) \
    X(L5, 4.0000000000000000e-01) \
    X(L3, 6.6666666666666663e-01) \
    X(L9, 2.2222222222222221e-01) \
    X(L7, 2.8571428571428570e-01) \
    X(L13, 1.5384615384615385e-01) \
    X(L11, 1.8181818181818182e-01) \
    X(L17, 1.1764705882352941e-01) \
    X(L15, 1.3333333333333333e-01) \
    X(L21, 9.5238095238095233e-02) \
    X(L19, 1.0526315789473684e-01) \
    X(PI_HI, 3.14159265358979311600e+00) \
    X(PI_LO, 1.22464679914735317723e-16) \
    X(S2, 8.33333333332248946124e-03) \
    X(S1_NEG, 1.66666666666666324348e-01) \
    X(S4, 2.75573137070700676789e-06) \
    X(S3_NEG, 1.98412698298579493134e-04) \
    X(S6, 1.58969099521155010221e-10) \
    X(S5_NEG, 2.50507602534068634195e-08) \
    X(C2_NEG, 1.38888888888741095749e-03) \
    X(C1, 4.16666666666666019037e-02) \
    X(C4_NEG, 2.75573143513906633035e-07) \
    X(C3, 2.48015872894767294178e-05) \
    X(C6_NEG, 1.13596475577881948265e-11) \
    X(C5, 2.08757232129817482790e-09) \
    X(TWO_OVER_PI, 6.36619772367581382433e-01) \
    X(PIO2_1, 1.57079632673412561417e+00) \
    X(PIO2_2, 6.07710050630396597660e-11) \
    X(PIO2_2T, 2.02226624879595063154e-21) \
    X(CB4_NEG, 3.47020774e-04) \
    X(CB3, 7.96603547e-03) \
    X(CB2_NEG, 7.33267142e-02) \
    X(CB1, 4.22924391e-01) \
    X(CB0, 6.48113736e-01)

enum MathConstant {
#define X(n, v) K_##n,
    GM_MATH_CONSTANTS(X)
#undef X
        K_COUNT
};
#ifdef __CUDACC__
__constant__ double gm_math_k[K_COUNT] = {
#define X(n, v) v,
    GM_MATH_CONSTANTS(X)
#undef X
};
#endif
static const double gm_math_k_host[K_COUNT] = {
#define X(n, v) v,
    GM_MATH_CONSTANTS(X)
#undef X
};
#ifdef __CUDA_ARCH__
#define C_(n) gm_math_k[K_##n]
#else
#define C_(n) gm_math_k_host[K_##n]
#endif

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
    /* negative or NaN x: the seed and therefore g are already NaN; only x == 0 (seed inf, g = 0 * inf) needs the
     * select */
    return x == 0.0 ? x : g;
}

/* exp(x), any finite x: 0 below -708 (denormal results are flushed), +inf above 709.78 */
GM_HD double exp_core(double r, double n) {
    /* exp(r) on |r| <= 0.3466, degree-11 minimax-like Taylor (|r|^12/12! < 7e-15 relative to 1: the last
     * coefficients are the classic fdlibm-style 1/k!), Estrin */
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p23 = fma(r, C_(E3), 0.5);
    const double p45 = fma(r, C_(E5), C_(E4));
    const double p67 = fma(r, C_(E7), C_(E6));
    const double p89 = fma(r, C_(E9), C_(E8));
    const double pab = fma(r, C_(E11), C_(E10));
    const double pcd = fma(r, C_(E13), C_(E12));
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

/* exp(x) for x known to lie in [-708, 709]: no range selects (table interpolants, ln r of the grid) */
GM_HD double exp_bounded(double x) {
    const double t = fma(x, C_(LOG2E), C_(MAGIC));
    const double n = t - C_(MAGIC);
    double r = fma(n, -C_(LN2_HI), x);
    r = fma(n, -C_(LN2_LO), r);
    return exp_core(r, n);
}

GM_HD double exp_(double x) {
    /* no clamp of x: inside [-708, 709.78] n stays in [-1021, 1024] (n = 1024 only with r < 0, p < 1, so the
     * exponent field still fits); outside, whatever the core produced is replaced by the selects below.  A NaN x
     * needs no select on the device: n, r and p are NaN, (int)NaN is 0, the scaling leaves the NaN alone. */
    double v = exp_bounded(x);
    v = x < -708.0 ? 0.0 : v;
    v = x > 709.78 ? from_bits(0x7ff0000000000000ull) : v;
#ifdef __CUDA_ARCH__
    return v;
#else
    return x != x ? x : v; /* the host model's (int)NaN is not 0 */
#endif
}

/* 10^x for x known to lie in [-307, 308] */
GM_HD double exp10_bounded(double x) {
    const double t = fma(x, C_(LOG2_10), C_(MAGIC));
    const double n = t - C_(MAGIC);
    /* x - n log10(2) in two parts, then to the natural base */
    double r = fma(n, -C_(LG2_HI), x);
    r = fma(n, -C_(LG2_LO), r);
    const double rh = r * C_(LN10_HI);
    const double rl = fma(r, C_(LN10_HI), -rh);
    const double rr = rh + fma(r, -C_(LN10_LO_NEG), rl);
    return exp_core(rr, n);
}

/* 10^x */
GM_HD double exp10_(double x) {
    double v = exp10_bounded(x);
    v = x < -307.0 ? 0.0 : v;
    v = x > 308.25 ? from_bits(0x7ff0000000000000ull) : v;
#ifdef __CUDA_ARCH__
    return v;
#else
    return x != x ? x : v;
#endif
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
    const double a01 = fma(v, C_(L5), C_(L3));
    const double a23 = fma(v, C_(L9), C_(L7));
    const double a45 = fma(v, C_(L13), C_(L11));
    const double a67 = fma(v, C_(L17), C_(L15));
    const double a89 = fma(v, C_(L21), C_(L19));
    const double b0 = fma(v2, a23, a01);
    const double b1 = fma(v2, a67, a45);
    const double c0 = fma(v4, b1, b0);
    const double poly = fma(v8, a89, c0);
    const double ed = (double)e;
    const double hi_part = fma(ed, C_(LN2_HI), 2.0 * s);
    const double lo_part = fma(ed, C_(LN2_LO), fma(s * v, poly, 2.0 * sl));
    return hi_part + lo_part;
}

/* sin and cos of pi*t for |t| < 2^30; exact at multiples of 1/2 */
GM_HD void sincospi_(double t, double *sp, double *cp) {
    const double q = rint(t + t);           /* nearest half-integer count */
    const double r = fma(q, -0.5, t);       /* |r| <= 1/4, exact */
    const int qi = (int)q;
    const double x = r * C_(PI_HI);
    const double xl = fma(r, C_(PI_HI), -x);
    const double xr = x + fma(r, C_(PI_LO), xl); /* pi r, |.| <= pi/4 */
    const double z = xr * xr, z2 = z * z, z4 = z2 * z2;
    /* sin(x) = x + x z (S1 + S2 z + ... + S6 z^5), cos(x) = 1 - z/2 + z^2 (C1 + ... + C6 z^5): fdlibm kernels */
    const double s12 = fma(z, C_(S2), -C_(S1_NEG));
    const double s34 = fma(z, C_(S4), -C_(S3_NEG));
    const double s56 = fma(z, C_(S6), -C_(S5_NEG));
    const double sp_ = fma(z4, s56, fma(z2, s34, s12));
    const double sn = fma(xr * z, sp_, xr);
    const double c12 = fma(z, -C_(C2_NEG), C_(C1));
    const double c34 = fma(z, -C_(C4_NEG), C_(C3));
    const double c56 = fma(z, -C_(C6_NEG), C_(C5));
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
    const double t = fma(x, C_(TWO_OVER_PI), C_(MAGIC));
    const double q = t - C_(MAGIC);
    const int qi = (int)q;
    double r = fma(q, -C_(PIO2_1), x);
    r = fma(q, -C_(PIO2_2), r);
    const double rl = q * C_(PIO2_2T);
    const double xr = r - rl;
    const double z = xr * xr, z2 = z * z, z4 = z2 * z2;
    const double s12 = fma(z, C_(S2), -C_(S1_NEG));
    const double s34 = fma(z, C_(S4), -C_(S3_NEG));
    const double s56 = fma(z, C_(S6), -C_(S5_NEG));
    const double sp_ = fma(z4, s56, fma(z2, s34, s12));
    const double sn = fma(xr * z, sp_, xr);
    const double c12 = fma(z, -C_(C2_NEG), C_(C1));
    const double c34 = fma(z, -C_(C4_NEG), C_(C3));
    const double c56 = fma(z, -C_(C6_NEG), C_(C5));
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
    double y = fma(m, fma(m, fma(m, fma(m, -C_(CB4_NEG), C_(CB3)), -C_(CB2_NEG)), C_(CB1)),
                   C_(CB0));
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
