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

__device__ __forceinline__ double ksin(double r)
{
    // fdlibm __kernel_sin minimax coefficients, |r| <= pi/4
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = __dmul_rn(r, r);
    double p = __fma_rn(z, S6, S5);
    p = __fma_rn(z, p, S4);
    p = __fma_rn(z, p, S3);
    p = __fma_rn(z, p, S2);
    p = __fma_rn(z, p, S1);
    return __fma_rn(__dmul_rn(r, z), p, r);
}

__device__ __forceinline__ double kcos(double r)
{
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = __dmul_rn(r, r);
    double p = __fma_rn(z, C6, C5);
    p = __fma_rn(z, p, C4);
    p = __fma_rn(z, p, C3);
    p = __fma_rn(z, p, C2);
    p = __fma_rn(z, p, C1);
    return __fma_rn(__dmul_rn(z, z), p, __fma_rn(z, -0.5, 1.0));
}

__device__ __forceinline__ void sincosf_portable(float a, float* s_out, float* c_out)
{
    const double PIO2_HI = 1.57079632673412561417e+00, PIO2_LO = 6.07710050650619224932e-11;
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    double x = (double)a;
    double kd = rint(__dmul_rn(x, TWO_OVER_PI));
    double r = __fma_rn(-kd, PIO2_HI, x);
    r = __fma_rn(-kd, PIO2_LO, r);
    int k = (int)kd;
    double s = ksin(r), c = kcos(r);
    double ss = (k & 1) ? c : s;
    double cc = (k & 1) ? s : c;
    if (k & 2) ss = -ss;
    if ((k + 1) & 2) cc = -cc;
    *s_out = (float)ss;
    *c_out = (float)cc;
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

__device__ __forceinline__ float atan2f_portable(float y, float x)
{
    return (float)atan2d_portable((double)y, (double)x);
}

__device__ __forceinline__ float asinf_portable(float v)
{
    double x = (double)v;
    if (!(fabs(x) <= 1.0)) return __int_as_float(0x7fc00000);
    double c = __dsqrt_rn(__fma_rn(-x, x, 1.0));
    return (float)atan2d_portable(x, c);
}

// exp of a binary32 argument (v3_redo absorption): k = rint(x / ln 2), two-part reduction, degree-13
// Taylor polynomial, scaling through the exponent field, one rounding to binary32
__device__ __forceinline__ float expf_portable(float a)
{
    double x = (double)a;
    if (x != x) return __int_as_float(0x7fc00000);
    if (x > 89.0) return __int_as_float(0x7f800000);
    if (x < -104.0) return 0.0f;
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
    return __double2float_rn(__dmul_rn(p, scale));
}

}  // namespace pm
}  // namespace b200pt
