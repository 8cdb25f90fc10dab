// pm_math.cuh -- parity-mode definitions of the four transcendentals on the hot path.
//
// The reference calls MSVC SVML (_mm256_sin_ps/_cos_ps/_sincos_ps/_atan2_ps/_asin_ps,
// mathlib.h:449-499; call sites demofox_path_tracing_v2.cpp:85-86, texture.cpp:112,148-149,
// 172-173,194-195), which is closed source and absent from the reference tree.  Parity mode
// therefore evaluates them in IEEE binary64 using only + - * / sqrt fma rint -- operations that
// round identically on sm_100a and on the host -- and rounds once to binary32.  The result is
// the correctly rounded value for every input we could test (tests/test_portable_math.py) and
// makes whole images bit-comparable between CPU and GPU.  B200 runs FP64 FMA at half the FP32
// rate, so this costs parity mode little; fast mode uses MUFU-based intrinsics instead.
#pragma once

namespace b200pt {
namespace pm {

// Polynomial coefficients live in constant memory: a DFMA takes a constant-bank operand directly,
// while a binary64 literal costs two extra (uniform-datapath) move instructions per use.
// fdlibm __kernel_sin / __kernel_cos minimax coefficients, |r| <= pi/4
static __constant__ double c_ksin[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                                        2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10};
static __constant__ double c_kcos[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                        -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};

__device__ __forceinline__ double ksin(double r)
{
    double z = __dmul_rn(r, r);
    double p = __fma_rn(z, c_ksin[5], c_ksin[4]);
    p = __fma_rn(z, p, c_ksin[3]);
    p = __fma_rn(z, p, c_ksin[2]);
    p = __fma_rn(z, p, c_ksin[1]);
    p = __fma_rn(z, p, c_ksin[0]);
    return __fma_rn(__dmul_rn(r, z), p, r);
}

__device__ __forceinline__ double kcos(double r)
{
    double z = __dmul_rn(r, r);
    double p = __fma_rn(z, c_kcos[5], c_kcos[4]);
    p = __fma_rn(z, p, c_kcos[3]);
    p = __fma_rn(z, p, c_kcos[2]);
    p = __fma_rn(z, p, c_kcos[1]);
    p = __fma_rn(z, p, c_kcos[0]);
    return __fma_rn(__dmul_rn(z, z), p, __fma_rn(z, -0.5, 1.0));
}

static __constant__ double c_reduce[4] = {6.36619772367581382433e-01 /* 2/pi */, 1.57079632673412561417e+00 /* pi/2 hi */,
                                          6.07710050650619224932e-11 /* pi/2 lo */, 0.0};

__device__ __forceinline__ void sincosf_portable(float a, float* s_out, float* c_out)
{
    double x = (double)a;
    double kd = rint(__dmul_rn(x, c_reduce[0]));
    double r = __fma_rn(-kd, c_reduce[1], x);
    r = __fma_rn(-kd, c_reduce[2], r);
    int k = (int)kd;
    // quadrant bookkeeping on the rounded binary32 values: rounding commutes with swapping and with negation
    const float s = __double2float_rn(ksin(r)), c = __double2float_rn(kcos(r));
    const float ss = (k & 1) ? c : s;
    const float cc = (k & 1) ? s : c;
    *s_out = __uint_as_float(__float_as_uint(ss) ^ (((unsigned)k << 30) & 0x80000000u));        // k & 2: negate
    *c_out = __uint_as_float(__float_as_uint(cc) ^ (((unsigned)(k + 1) << 30) & 0x80000000u));  // (k + 1) & 2
}

__device__ __forceinline__ double atan01(double t)
{
    double kd = rint(__dmul_rn(t, 4.0));
    double c = __dmul_rn(kd, 0.25);
    double z = __ddiv_rn(__dsub_rn(t, c), __fma_rn(t, c, 1.0));
    double w = __dmul_rn(z, z);
    double p = -1.0 / 19.0;
    p = __fma_rn(w, p, 1.0 / 17.0);
    p = __fma_rn(w, p, -1.0 / 15.0);
    p = __fma_rn(w, p, 1.0 / 13.0);
    p = __fma_rn(w, p, -1.0 / 11.0);
    p = __fma_rn(w, p, 1.0 / 9.0);
    p = __fma_rn(w, p, -1.0 / 7.0);
    p = __fma_rn(w, p, 1.0 / 5.0);
    p = __fma_rn(w, p, -1.0 / 3.0);
    int ki = (int)kd;
    double base = ki == 0 ? 0.0
                : ki == 1 ? 2.44978663126864143e-01
                : ki == 2 ? 4.63647609000806094e-01
                : ki == 3 ? 6.43501108793284371e-01
                          : 7.85398163397448279e-01;
    return __dadd_rn(base, __fma_rn(__dmul_rn(z, w), p, z));
}

__device__ __forceinline__ double atan2d_portable(double y, double x)
{
    const double PI = 3.14159265358979311600e+00, PIO2 = 1.57079632679489655800e+00;
    double ax = fabs(x), ay = fabs(y);
    if (ax != ax || ay != ay) return __longlong_as_double(0x7ff8000000000000LL);
    double mx = ax > ay ? ax : ay;
    double mn = ax > ay ? ay : ax;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double r;
    if (mx == 0.0) r = 0.0;
    else if (mx == inf) r = (mn == inf) ? 0.78539816339744827900 : 0.0;
    else r = atan01(__ddiv_rn(mn, mx));
    if (ay > ax) r = __dsub_rn(PIO2, r);
    if (signbit(x)) r = __dsub_rn(PI, r);
    return copysign(r, y);
}

// ---- the literal algorithms of oracle/portable_math.h (second tier; also the test reference) ----
static __device__ __noinline__ float atan2f_literal(float y, float x)
{
    return (float)atan2d_portable((double)y, (double)x);
}

static __device__ __noinline__ float asinf_literal(float v)
{
    double x = (double)v;
    if (!(fabs(x) <= 1.0)) return __int_as_float(0x7fc00000);
    double c = __dsqrt_rn(__fma_rn(-x, x, 1.0));
    return (float)atan2d_portable(x, c);
}

// ---- first tier ---------------------------------------------------------------------------
// The literal algorithms cost ~130-150 instructions each (two IEEE binary64 divisions, a square
// root, selects on doubles).  Their binary64 result differs from the real atan2 / asin by a few
// units of 2^-52, and only its rounding to binary32 is used.  The first tier computes the same real
// number to better than 2^-48 with ~45 instructions (Newton-refined MUFU seeds, one polynomial;
// coefficients: scripts/gen_minimax.py) and uses its rounding whenever the value keeps a distance of
// 2^-42 (relative) from every binary32 rounding boundary -- then both tiers round alike.  Otherwise
// (probability 2^-18 per call), and for zeros, infinities, NaNs and results below 2^-120, the literal
// algorithm runs.  b200pt_check_portable_tiers() compares the tiers on the device (exhaustively for
// asin), tests/test_gpu_parity.py compares the result with the oracle.
// atan(t)/t on [0, 1], degree 18 in u = t^2 (error < 2^-51), lowest power first
static __constant__ double c_atan[19] = {
    9.99999999999999778e-01, -3.33333333333187987e-01, 1.99999999982444054e-01, -1.42857142011534849e-01,
    1.11111089472156838e-01, -9.09087510009998212e-02, 7.69195041879679464e-02, -6.66400760765404609e-02,
    5.86777600170671904e-02, -5.20258292466982061e-02, 4.56687247194499449e-02, -3.85260683461067371e-02,
    2.99240529380271104e-02, -2.03308700330274601e-02, 1.14300236750545080e-02, -5.00036645373264955e-03,
    1.57394932417557619e-03, -3.14200773668073339e-04, 2.96963566016577910e-05,
};
// (asin(t) - t)/t^3 on [0, 1/2], degree 12 in u = t^2 (error < 2^-52), lowest power first
static __constant__ double c_asin[13] = {
    1.66666666666666685e-01, 7.49999999999843292e-02, 4.46428571463554288e-02, 3.03819441385312465e-02,
    2.23721729421498886e-02, 1.73523927208699726e-02, 1.39712129735529329e-02, 1.14791774151849057e-02,
    1.03228143501857793e-02, 5.45750671864035815e-03, 1.74008794426940214e-02, -1.48518870712472037e-02,
    2.87578513674215663e-02,
};

__device__ __forceinline__ double rcp_newton(double d)  // d finite, normal, nonzero
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = __fma_rn(-d, y, 1.0);
    y = __fma_rn(y, e, y);
    e = __fma_rn(-d, y, 1.0);
    return __fma_rn(y, e, y);
}

__device__ __forceinline__ double sqrt_newton(double a)  // a in [2^-60, 4]
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double g = __dmul_rn(a, y), h = __dmul_rn(y, 0.5);
    double r = __fma_rn(-h, g, 0.5);
    g = __fma_rn(g, r, g);
    h = __fma_rn(h, r, h);
    r = __fma_rn(-h, g, 0.5);
    return __fma_rn(g, r, g);
}

// true when rounding the binary64 value d (0 < d < 2^100) to binary32 cannot be changed by an error of
// 1024 units in its last place: the 29 dropped bits stay away from the half-way pattern, and the
// result is a normal binary32 number
__device__ __forceinline__ bool rounds_safely(double d)
{
    const unsigned lo = (unsigned)__double2loint(d), hi = (unsigned)__double2hiint(d);
    const unsigned dropped = lo & 0x1FFFFFFFu;
    return (dropped - (0x10000000u - 1024u)) > 2048u && hi >= 0x38700000u;  // |d| >= 2^-120
}

// first tier: true and the result when it is decisive
__device__ __forceinline__ bool atan2f_first_tier(float yf, float xf, float& out)
{
    const double PI = 3.14159265358979311600e+00, PIO2 = 1.57079632679489655800e+00;
    const double x = (double)xf, y = (double)yf;
    const double ax = fabs(x), ay = fabs(y);
    const unsigned ux = __float_as_uint(xf) & 0x7fffffffu, uy = __float_as_uint(yf) & 0x7fffffffu;
    if (max(ux, uy) - 1u < 0x7f7fffffu) {  // both finite (no NaN), not both zero
        const bool steep = ay > ax;
        const double mx = steep ? ay : ax, mn = steep ? ax : ay;
        const double t = __dmul_rn(mn, rcp_newton(mx));  // [0, 1]
        const double u = __dmul_rn(t, t);
        double p = c_atan[18];
        p = __fma_rn(u, p, c_atan[17]);
        p = __fma_rn(u, p, c_atan[16]);
        p = __fma_rn(u, p, c_atan[15]);
        p = __fma_rn(u, p, c_atan[14]);
        p = __fma_rn(u, p, c_atan[13]);
        p = __fma_rn(u, p, c_atan[12]);
        p = __fma_rn(u, p, c_atan[11]);
        p = __fma_rn(u, p, c_atan[10]);
        p = __fma_rn(u, p, c_atan[9]);
        p = __fma_rn(u, p, c_atan[8]);
        p = __fma_rn(u, p, c_atan[7]);
        p = __fma_rn(u, p, c_atan[6]);
        p = __fma_rn(u, p, c_atan[5]);
        p = __fma_rn(u, p, c_atan[4]);
        p = __fma_rn(u, p, c_atan[3]);
        p = __fma_rn(u, p, c_atan[2]);
        p = __fma_rn(u, p, c_atan[1]);
        p = __fma_rn(u, p, c_atan[0]);
        double r = __dmul_rn(t, p);
        if (steep) r = __dsub_rn(PIO2, r);
        if (__float_as_int(xf) < 0) r = __dsub_rn(PI, r);  // signbit(x), -0 included
        if (rounds_safely(r)) {
            out = copysignf(__double2float_rn(r), yf);
            return true;
        }
    }
    return false;
}
__device__ __forceinline__ float atan2f_portable(float yf, float xf)
{
    float r;
    if (atan2f_first_tier(yf, xf, r)) return r;
    // NaN directions are not rare (v3_redo / v4: total internal reflection in the roughness-0 sphere yields a zero
    // vector, normalised to NaN, exactly as in the reference): the literal algorithm's answer is the quiet NaN
    if (yf != yf || xf != xf) return __int_as_float(0x7fc00000);
    return atan2f_literal(yf, xf);
}

__device__ __forceinline__ bool asinf_first_tier(float v, float& out)
{
    const double PIO2 = 1.57079632679489655800e+00;
    const double ax = fabs((double)v);
    if (ax < 1.0) {  // |v| == 1, |v| > 1 and NaN: literal algorithm
        const bool big = ax > 0.5;  // asin(x) = pi/2 - 2 asin(sqrt((1 - x) / 2))
        double u = __dmul_rn(ax, ax), t = ax;
        if (big) {
            u = __fma_rn(ax, -0.5, 0.5);  // exact
            t = sqrt_newton(u);
        }
        double p = c_asin[12];
        p = __fma_rn(u, p, c_asin[11]);
        p = __fma_rn(u, p, c_asin[10]);
        p = __fma_rn(u, p, c_asin[9]);
        p = __fma_rn(u, p, c_asin[8]);
        p = __fma_rn(u, p, c_asin[7]);
        p = __fma_rn(u, p, c_asin[6]);
        p = __fma_rn(u, p, c_asin[5]);
        p = __fma_rn(u, p, c_asin[4]);
        p = __fma_rn(u, p, c_asin[3]);
        p = __fma_rn(u, p, c_asin[2]);
        p = __fma_rn(u, p, c_asin[1]);
        p = __fma_rn(u, p, c_asin[0]);
        double r = __fma_rn(__dmul_rn(t, u), p, t);  // asin(t) = t + t^3 P(t^2)
        if (big) r = __fma_rn(r, -2.0, PIO2);
        if (rounds_safely(r)) {
            out = copysignf(__double2float_rn(r), v);
            return true;
        }
    }
    return false;
}
__device__ __forceinline__ float asinf_portable(float v)
{
    float r;
    if (asinf_first_tier(v, r)) return r;
    if (!(fabsf(v) <= 1.0f)) return __int_as_float(0x7fc00000);  // |v| > 1 or NaN: what the literal algorithm returns
    return asinf_literal(v);
}

// exp of a binary32 argument (v3_redo absorption): k = rint(x / ln 2), two-part reduction, degree-13
// Taylor polynomial, scaling through the exponent field, one rounding to binary32
__device__ __forceinline__ double exp_core(double x);
__device__ __forceinline__ float expf_portable(float a)
{
    if (a == 0.f) return 1.0f;  // what the evaluation below gives; the built-in refraction colour has a zero channel
    double x = (double)a;
    if (x != x) return __int_as_float(0x7fc00000);
    if (x > 89.0) return __int_as_float(0x7f800000);
    if (x < -104.0) return 0.0f;
    return __double2float_rn(exp_core(x));
}
// exp of a binary64 argument in [-104, 89], not yet rounded to binary32
__device__ __forceinline__ double exp_core(double x)
{
    const double LOG2E = 1.44269504088896338700e+00;
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    double kd = rint(__dmul_rn(x, LOG2E));
    double r = __fma_rn(-kd, LN2_HI, x);
    r = __fma_rn(-kd, LN2_LO, r);
    double p = 1.0 / 6227020800.0;
    p = __fma_rn(r, p, 1.0 / 479001600.0);
    p = __fma_rn(r, p, 1.0 / 39916800.0);
    p = __fma_rn(r, p, 1.0 / 3628800.0);
    p = __fma_rn(r, p, 1.0 / 362880.0);
    p = __fma_rn(r, p, 1.0 / 40320.0);
    p = __fma_rn(r, p, 1.0 / 5040.0);
    p = __fma_rn(r, p, 1.0 / 720.0);
    p = __fma_rn(r, p, 1.0 / 120.0);
    p = __fma_rn(r, p, 1.0 / 24.0);
    p = __fma_rn(r, p, 1.0 / 6.0);
    p = __fma_rn(r, p, 0.5);
    p = __fma_rn(r, p, 1.0);
    p = __fma_rn(r, p, 1.0);
    const double scale = __longlong_as_double((long long)((long long)kd + 1023) << 52);
    return __dmul_rn(p, scale);
}

// pow(x, y) for x >= 0 (the non-fast gamma: pow_ps(rgb, 1/2.4), v4.cpp:185) = oracle/portable_math.h's pm_powf, operation for
// operation: exp(y log x) in binary64 -- x = 2^e m, m in [sqrt(1/2), sqrt(2)], log m = 2 s (1 + z/3 + ... + z^10/21) with
// s = (m - 1) / (m + 1), z = s^2 -- rounded once to binary32
__device__ __forceinline__ float powf_portable(float xf, float yf)
{
    const double x = (double)xf, y = (double)yf;
    const float nan = __int_as_float(0x7fc00000), inf = __int_as_float(0x7f800000);
    if (y == 0.0 || x == 1.0) return 1.0f;
    if (x != x || y != y || x < 0.0) return nan;
    if (x == 0.0) return y > 0.0 ? 0.0f : inf;
    if (x == (double)inf) return y > 0.0 ? inf : 0.0f;
    if (y == (double)inf) return x < 1.0 ? 0.0f : inf;
    if (y == -(double)inf) return x < 1.0 ? inf : 0.0f;
    long long u = __double_as_longlong(x);
    long long e = ((u >> 52) & 0x7ff) - 1023;
    double m = __longlong_as_double((u & 0x000fffffffffffffll) | 0x3ff0000000000000ll);
    if (m > 1.41421356237309514547) {
        m = __dmul_rn(m, 0.5);
        e += 1;
    }
    const double s = __ddiv_rn(__dadd_rn(m, -1.0), __dadd_rn(m, 1.0));
    const double z = __dmul_rn(s, s);
    double p = 1.0 / 21.0;
    p = __fma_rn(z, p, 1.0 / 19.0);
    p = __fma_rn(z, p, 1.0 / 17.0);
    p = __fma_rn(z, p, 1.0 / 15.0);
    p = __fma_rn(z, p, 1.0 / 13.0);
    p = __fma_rn(z, p, 1.0 / 11.0);
    p = __fma_rn(z, p, 1.0 / 9.0);
    p = __fma_rn(z, p, 1.0 / 7.0);
    p = __fma_rn(z, p, 1.0 / 5.0);
    p = __fma_rn(z, p, 1.0 / 3.0);
    p = __fma_rn(z, p, 1.0);
    const double logm = __dmul_rn(__dadd_rn(s, s), p);
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double ed = (double)e;
    const double lx = __fma_rn(ed, LN2_HI, __fma_rn(ed, LN2_LO, logm));
    const double t = __dmul_rn(y, lx);
    if (t > 89.0) return inf;
    if (t < -104.0) return 0.0f;
    return __double2float_rn(exp_core(t));
}

}  // namespace pm
}  // namespace b200pt
