/*
 * shim.h -- TEST INFRASTRUCTURE.  Force-included in front of the UNMODIFIED reference sources
 * (read in place from /root/reference; never copied) so that the MSVC/Win32 dialect they are
 * written in compiles under g++ 13.  It adds no algorithm: only
 *   (1) class wrappers for __m256 / __m256i (GCC forbids operator overloads on raw vector
 *       types; the reference declares them in mathlib.h:571-726 and mathlib.h:810-857, and
 *       indexes lanes through MSVC's .m256_f32[] union member, texture.cpp:107-134),
 *   (2) stand-ins for the MSVC SVML intrinsics (mathlib.h:449-499, :823),
 *   (3) in ORACLE_EXACT mode, _mm256_rcp_ps -> 1/x and _mm256_rsqrt_ps -> 1/sqrt(x)
 *       (the 12-bit hardware approximations are CPU-vendor specific).
 *
 * Modes (compile-time):
 *   -DORACLE_EXACT=1  rcp/rsqrt exact, transcendentals = oracle/portable_math.h  (parity anchor)
 *   -DORACLE_EXACT=0  hardware rcpps/rsqrtps, transcendentals = glibc libm per lane ("asis":
 *                     CPU timing + statistical comparison only)
 */
#ifndef ORACLE_REF_SHIM_H
#define ORACLE_REF_SHIM_H

#include <immintrin.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdint>

#include "portable_math.h"

#ifndef ORACLE_EXACT
#define ORACLE_EXACT 1
#endif

/* runtime knobs the build recipe substitutes for compile-time constants (see build_ref.sh) */
extern int oracle_num_threads; /* stands in for NUM_THREADS, global_preprocessor_flags.h:69 */

union M256W {
    __m256 v;
    float m256_f32[8];
    M256W() = default;
    M256W(__m256 x) : v(x) {}
    M256W(float x) : v(_mm256_set1_ps(x)) {}
    M256W(int x) : v(_mm256_set1_ps((float)x)) {}
    operator __m256() const { return v; }
};

union M256IW {
    __m256i v;
    int m256i_i32[8];
    M256IW() = default;
    M256IW(__m256i x) : v(x) {}
    M256IW(int x) : v(_mm256_set1_epi32(x)) {}
    operator __m256i() const { return v; }
};

/* ---- SVML stand-ins (take/return the raw vector types; wrappers convert implicitly) ---- */
#define ORACLE_LANEWISE1(NAME, EXPR)                                   \
    static inline __m256 NAME(__m256 a_)                               \
    {                                                                  \
        float a[8], r[8];                                              \
        _mm256_storeu_ps(a, a_);                                       \
        for (int i = 0; i < 8; i++) { float x = a[i]; r[i] = (EXPR); } \
        return _mm256_loadu_ps(r);                                     \
    }
#define ORACLE_LANEWISE2(NAME, EXPR)                                                 \
    static inline __m256 NAME(__m256 a_, __m256 b_)                                  \
    {                                                                                \
        float a[8], b[8], r[8];                                                      \
        _mm256_storeu_ps(a, a_);                                                     \
        _mm256_storeu_ps(b, b_);                                                     \
        for (int i = 0; i < 8; i++) { float x = a[i], y = b[i]; r[i] = (EXPR); }     \
        return _mm256_loadu_ps(r);                                                   \
    }

#if ORACLE_EXACT
ORACLE_LANEWISE1(_mm256_sin_ps, pm_sinf(x))
ORACLE_LANEWISE1(_mm256_cos_ps, pm_cosf(x))
ORACLE_LANEWISE1(_mm256_asin_ps, pm_asinf(x))
ORACLE_LANEWISE2(_mm256_atan2_ps, pm_atan2f(x, y))
ORACLE_LANEWISE1(_mm256_exp_ps, pm_expf(x))
ORACLE_LANEWISE2(_mm256_pow_ps, pm_powf(x, y))
#else
ORACLE_LANEWISE1(_mm256_sin_ps, sinf(x))
ORACLE_LANEWISE1(_mm256_cos_ps, cosf(x))
ORACLE_LANEWISE1(_mm256_asin_ps, asinf(x))
ORACLE_LANEWISE2(_mm256_atan2_ps, atan2f(x, y))
ORACLE_LANEWISE1(_mm256_exp_ps, expf(x))
ORACLE_LANEWISE2(_mm256_pow_ps, powf(x, y))
#endif
ORACLE_LANEWISE1(_mm256_tan_ps, tanf(x))
ORACLE_LANEWISE1(_mm256_acos_ps, acosf(x))
ORACLE_LANEWISE1(_mm256_atan_ps, atanf(x))
ORACLE_LANEWISE1(_mm256_log_ps, logf(x))

static inline __m256i _mm256_div_epi32(__m256i a_, __m256i b_)
{
    int a[8], b[8], r[8];
    _mm256_storeu_si256((__m256i*)a, a_);
    _mm256_storeu_si256((__m256i*)b, b_);
    for (int i = 0; i < 8; i++) r[i] = b[i] ? a[i] / b[i] : 0;
    return _mm256_loadu_si256((__m256i*)r);
}

#if ORACLE_EXACT
static inline __m256 oracle_exact_rcp(__m256 a) { return _mm256_div_ps(_mm256_set1_ps(1.0f), a); }
static inline __m256 oracle_exact_rsqrt(__m256 a)
{
    return _mm256_div_ps(_mm256_set1_ps(1.0f), _mm256_sqrt_ps(a));
}
#define _mm256_rcp_ps(x) oracle_exact_rcp(x)
#define _mm256_rsqrt_ps(x) oracle_exact_rsqrt(x)
#endif

/* from here on the reference's spelling of the vector types means the wrappers */
#define __m256 M256W
#define __m256i M256IW

static inline __m256 _mm256_sincos_ps(__m256* p_cos, __m256 a)
{
    float x[8], s[8], c[8];
    _mm256_storeu_ps(x, a.v);
    for (int i = 0; i < 8; i++) {
#if ORACLE_EXACT
        pm_sincosf(x[i], &s[i], &c[i]);
#else
        s[i] = sinf(x[i]);
        c[i] = cosf(x[i]);
#endif
    }
    *p_cos = M256W(_mm256_loadu_ps(c));
    return M256W(_mm256_loadu_ps(s));
}

/* The reference calls unqualified tan(f32) / atan2(f32,f32) / asin(f32)
 * (demofox_path_tracing_v2.cpp:546, texture.cpp:91,112).  With MSVC's <cmath> and with
 * libstdc++'s <math.h> wrapper these resolve to the single-precision overloads.  In exact mode
 * the two that sit on the hot path are routed to the portable definitions instead of glibc. */
#if ORACLE_EXACT
static inline float oracle_atan2(float y, float x) { return pm_atan2f(y, x); }
static inline float oracle_asin(float x) { return pm_asinf(x); }
static inline double oracle_atan2(double y, double x) { return ::atan2(y, x); }
static inline double oracle_asin(double x) { return ::asin(x); }
#define atan2 oracle_atan2
#define asin oracle_asin
#endif

#define __declspec(x)

#endif
