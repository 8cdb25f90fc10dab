/*
 * portable_math.h -- TEST INFRASTRUCTURE (oracle side). Not linked by the product.
 *
 * Bit-reproducible stand-ins for the closed-source MSVC SVML calls on the hot path
 * (reference mathlib.h:449-499: _mm256_sin_ps/_cos_ps/_sincos_ps/_atan2_ps/_asin_ps,
 * call sites demofox_path_tracing_v2.cpp:85-86, demofox_path_tracing_simt_textured.cpp:85-86,
 * texture.cpp:91,112,148-149,172-173,194-195; _mm256_exp_ps: demofox_path_tracing_v3_redo.cpp:649-651 and v4's
 * USE_FAST_APPROXIMATE_EXP 0; _mm256_pow_ps: v4's USE_FAST_APPROXIMATE_GAMMA 0, ..._optimization_v4.cpp:185).
 *
 * SVML is absent from /root/reference (it ships inside the MSVC v142 runtime, no version pin,
 * no source), so no golden vector pins this boundary: "parity unpinned" for these
 * functions (sin, cos, atan2, asin, exp, pow).  The oracle therefore DEFINES them: evaluated in IEEE binary64 with only
 * + - * / sqrt fma rint (every one of which is correctly rounded on x86-64 and on sm_100a),
 * then rounded once to binary32.  The CUDA parity kernel carries an independent copy of the
 * same algorithm (csrc/pm_math.cuh), so CPU and GPU agree bit for bit; tests/test_portable_math.py
 * checks the definitions against glibc libm (max 1 ulp apart, > 99.99 % identical).
 */
#ifndef ORACLE_PORTABLE_MATH_H
#define ORACLE_PORTABLE_MATH_H

#include <math.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_FMA(a, b, c) __builtin_fma((a), (b), (c))

/* pi/2 split: HI has 33 significant bits so k*HI is exact for |k| < 2^20 */
#define PM_PIO2_HI 1.57079632673412561417e+00 /* 0x3FF921FB54400000 */
#define PM_PIO2_LO 6.07710050650619224932e-11 /* 0x3DD0B4611A626331 */
#define PM_TWO_OVER_PI 6.36619772367581382433e-01
#define PM_PI 3.14159265358979311600e+00
#define PM_PIO2 1.57079632679489655800e+00

/* fdlibm __kernel_sin / __kernel_cos minimax coefficients on [-pi/4, pi/4] */
static inline double pm_ksin(double r)
{
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = r * r;
    double p = PM_FMA(z, S6, S5);
    p = PM_FMA(z, p, S4);
    p = PM_FMA(z, p, S3);
    p = PM_FMA(z, p, S2);
    p = PM_FMA(z, p, S1);
    return PM_FMA(r * z, p, r);
}

static inline double pm_kcos(double r)
{
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = r * r;
    double p = PM_FMA(z, C6, C5);
    p = PM_FMA(z, p, C4);
    p = PM_FMA(z, p, C3);
    p = PM_FMA(z, p, C2);
    p = PM_FMA(z, p, C1);
    /* 1 - z/2 + z^2 * p */
    return PM_FMA(z * z, p, PM_FMA(z, -0.5, 1.0));
}

/* sin and cos of a binary32 angle, |a| < 2^20 (the hot path only passes [0, 2*pi]) */
static inline void pm_sincosf(float a, float* s_out, float* c_out)
{
    double x = (double)a;
    double kd = __builtin_rint(x * PM_TWO_OVER_PI);
    double r = PM_FMA(-kd, PM_PIO2_HI, x);
    r = PM_FMA(-kd, PM_PIO2_LO, r);
    int k = (int)kd;
    double s = pm_ksin(r), c = pm_kcos(r);
    double ss, cc;
    switch (k & 3) {
    case 0: ss = s; cc = c; break;
    case 1: ss = c; cc = -s; break;
    case 2: ss = -s; cc = -c; break;
    default: ss = -c; cc = s; break;
    }
    *s_out = (float)ss;
    *c_out = (float)cc;
}

static inline float pm_sinf(float a) { float s, c; pm_sincosf(a, &s, &c); return s; }
static inline float pm_cosf(float a) { float s, c; pm_sincosf(a, &s, &c); return c; }

/* atan(t) for t in [0, 1]: table of atan(k/4) + 9-term Taylor series of the reduced argument */
static inline double pm_atan01(double t)
{
    static const double ATAN_K4[5] = {
        0.0,
        2.44978663126864143e-01, /* atan(0.25) */
        4.63647609000806094e-01, /* atan(0.50) */
        6.43501108793284371e-01, /* atan(0.75) */
        7.85398163397448279e-01  /* atan(1.00) */
    };
    double kd = __builtin_rint(t * 4.0);
    double c = kd * 0.25;
    double z = (t - c) / PM_FMA(t, c, 1.0); /* |z| <= 0.1251 */
    double w = z * z;
    double p = -1.0 / 19.0;
    p = PM_FMA(w, p, 1.0 / 17.0);
    p = PM_FMA(w, p, -1.0 / 15.0);
    p = PM_FMA(w, p, 1.0 / 13.0);
    p = PM_FMA(w, p, -1.0 / 11.0);
    p = PM_FMA(w, p, 1.0 / 9.0);
    p = PM_FMA(w, p, -1.0 / 7.0);
    p = PM_FMA(w, p, 1.0 / 5.0);
    p = PM_FMA(w, p, -1.0 / 3.0);
    return ATAN_K4[(int)kd] + PM_FMA(z * w, p, z);
}

static inline double pm_atan2d(double y, double x)
{
    double ax = fabs(x), ay = fabs(y);
    if (ax != ax || ay != ay) return NAN;
    double mx = ax > ay ? ax : ay;
    double mn = ax > ay ? ay : ax;
    double r;
    if (mx == 0.0) r = 0.0;
    else if (mx == INFINITY) r = (mn == INFINITY) ? 0.78539816339744827900 : 0.0;
    else r = pm_atan01(mn / mx);
    if (ay > ax) r = PM_PIO2 - r;
    if (signbit(x)) r = PM_PI - r;
    return copysign(r, y);
}

static inline float pm_atan2f(float y, float x) { return (float)pm_atan2d((double)y, (double)x); }

static inline float pm_asinf(float v)
{
    double x = (double)v;
    if (!(fabs(x) <= 1.0)) return NAN;
    double c = sqrt(PM_FMA(-x, x, 1.0));
    return (float)pm_atan2d(x, c);
}

/* exp of a binary32 argument (v3_redo absorption, demofox_path_tracing_v3_redo.cpp:649-651):
 * k = rint(x / ln 2), r = x - k ln 2 (two-part ln 2), degree-13 Taylor polynomial of exp(r),
 * |r| <= 0.347, scaled by 2^k through the exponent field, rounded once to binary32. */
static inline double pm_exp_core(double x);
static inline float pm_expf(float a)
{
    double x = (double)a;
    if (x != x) return NAN;
    if (x > 89.0) return INFINITY;
    if (x < -104.0) return 0.0f;
    return (float)pm_exp_core(x);
}
/* exp of a binary64 argument in [-104, 89], result in binary64 (not yet rounded to binary32) */
static inline double pm_exp_core(double x)
{
    const double LOG2E = 1.44269504088896338700e+00;
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    double kd = __builtin_rint(x * LOG2E);
    double r = PM_FMA(-kd, LN2_HI, x);
    r = PM_FMA(-kd, LN2_LO, r);
    double p = 1.0 / 6227020800.0; /* 1/13! */
    p = PM_FMA(r, p, 1.0 / 479001600.0);
    p = PM_FMA(r, p, 1.0 / 39916800.0);
    p = PM_FMA(r, p, 1.0 / 3628800.0);
    p = PM_FMA(r, p, 1.0 / 362880.0);
    p = PM_FMA(r, p, 1.0 / 40320.0);
    p = PM_FMA(r, p, 1.0 / 5040.0);
    p = PM_FMA(r, p, 1.0 / 720.0);
    p = PM_FMA(r, p, 1.0 / 120.0);
    p = PM_FMA(r, p, 1.0 / 24.0);
    p = PM_FMA(r, p, 1.0 / 6.0);
    p = PM_FMA(r, p, 0.5);
    p = PM_FMA(r, p, 1.0);
    p = PM_FMA(r, p, 1.0);
    union { uint64_t u; double d; } scale;
    scale.u = (uint64_t)((int64_t)kd + 1023) << 52;
    return p * scale.d;
}

/* pow(x, y) for x >= 0 (the non-fast gamma, pow_ps(rgb, 1/2.4), demofox_path_tracing_optimization_v4.cpp:185; SVML
 * _mm256_pow_ps in the reference): exp(y * log x) in binary64, rounded once to binary32.
 * log x: x = 2^e * m with m in [sqrt(1/2), sqrt(2)], s = (m - 1) / (m + 1), log m = 2 s (1 + z/3 + ... + z^10/21), z = s^2
 * (|s| <= 0.1716: the first omitted term is below 2^-60 relative), log x = e * LN2_HI + (log m + e * LN2_LO) with e * LN2_HI
 * exact.  The product y * log x carries <= 2^-52 relative error, so the binary32 result is the correctly rounded power except
 * within ~|y log x| * 2^-28 ulp of a rounding boundary.  x < 0 -> NaN (the caller saturates to [0, 1] first). */
static inline float pm_powf(float xf, float yf)
{
    double x = (double)xf, y = (double)yf;
    if (y == 0.0 || x == 1.0) return 1.0f;
    if (x != x || y != y || x < 0.0) return NAN;
    if (x == 0.0) return y > 0.0 ? 0.0f : INFINITY;
    if (x == INFINITY) return y > 0.0 ? INFINITY : 0.0f;
    if (y == INFINITY) return x < 1.0 ? 0.0f : INFINITY;
    if (y == -INFINITY) return x < 1.0 ? INFINITY : 0.0f;
    union { double d; uint64_t u; } bits;
    bits.d = x; /* a binary32 value is a normal binary64 number */
    int64_t e = (int64_t)((bits.u >> 52) & 0x7ff) - 1023;
    bits.u = (bits.u & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    double m = bits.d; /* [1, 2) */
    if (m > 1.41421356237309514547) {
        m = m * 0.5;
        e += 1;
    }
    const double s = (m - 1.0) / (m + 1.0);
    const double z = s * s;
    double p = 1.0 / 21.0;
    p = PM_FMA(z, p, 1.0 / 19.0);
    p = PM_FMA(z, p, 1.0 / 17.0);
    p = PM_FMA(z, p, 1.0 / 15.0);
    p = PM_FMA(z, p, 1.0 / 13.0);
    p = PM_FMA(z, p, 1.0 / 11.0);
    p = PM_FMA(z, p, 1.0 / 9.0);
    p = PM_FMA(z, p, 1.0 / 7.0);
    p = PM_FMA(z, p, 1.0 / 5.0);
    p = PM_FMA(z, p, 1.0 / 3.0);
    p = PM_FMA(z, p, 1.0);
    const double logm = (s + s) * p;
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double ed = (double)e;
    const double lx = PM_FMA(ed, LN2_HI, PM_FMA(ed, LN2_LO, logm));
    const double t = y * lx;
    if (t > 89.0) return INFINITY;
    if (t < -104.0) return 0.0f;
    return (float)pm_exp_core(t);
}

#ifdef __cplusplus
}
#endif
#endif
