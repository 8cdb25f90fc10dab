// pt_tonemap.cuh -- LinearToSRGB(ACESFilm(c)) -> 8 bit, the reference's OutputToScreen / OutputToFile
// pixel math (demofox_path_tracing_optimization_v4.cpp:144-187,1279-1290,1316-1325) with the fast
// ACES / fast gamma variants (global_preprocessor_flags.h:62-63).  rcp and sqrt are the correctly
// rounded ones in every arithmetic policy (the oracle's "exact" definition), and every fused
// operation is spelled out, so the 8-bit result is the same in all translation units.
#pragma once
#include <stdint.h>

#include "pm_math.cuh"

namespace b200pt {
namespace tonemap {

__device__ __forceinline__ float max_ps(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float min_ps(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float saturate(float x) { return min_ps(max_ps(x, 0.f), 1.f); }

__device__ __forceinline__ float fast_pow_gamma(float x)  // v4.cpp:144-155
{
    const float sqrtx = __fsqrt_rn(x);
    const float onethird = 1.f / 3.f, twothirds = 2.f / 3.f;
    const float nit1 = fmaf(sqrtx, twothirds, onethird);
    const float nit2 = fmaf(nit1, twothirds, __fmul_rn(__fmul_rn(x, __frcp_rn(__fmul_rn(nit1, nit1))), onethird));
    const float nit3 = fmaf(nit2, twothirds, __fmul_rn(__fmul_rn(x, __frcp_rn(__fmul_rn(nit2, nit2))), onethird));
    return __fsqrt_rn(__fmul_rn(sqrtx, nit3));
}
__device__ __forceinline__ float aces(float X)  // v4.cpp:166-176
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    const float rcpDenom = __frcp_rn(fmaf(X, fmaf(c, X, d), e));
    return saturate(__fmul_rn(__fmul_rn(X, fmaf(a, X, b)), rcpDenom));
}
__device__ __forceinline__ float aces_exact(float X)  // USE_FAST_APPROXIMATE_ACES_TONEMAP 0, v4.cpp:172-175: no fused operation, a true division
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    const float num = __fmul_rn(X, __fadd_rn(__fmul_rn(a, X), b));
    const float den = __fadd_rn(__fmul_rn(X, __fadd_rn(__fmul_rn(c, X), d)), e);
    return saturate(__fdiv_rn(num, den));
}
__device__ __forceinline__ float srgb(float v)  // v4.cpp:178-187
{
    v = saturate(v);
    return (v < 0.0031308f) ? __fmul_rn(v, 12.92f) : fmaf(1.055f, fast_pow_gamma(v), -0.055f);
}
// USE_FAST_APPROXIMATE_GAMMA 0, v4.cpp:185: 1.055f * pow_ps(rgb, 1.f / 2.4f) - 0.055f with pow_ps as the oracle defines it
// (pm::powf_portable).  Only the resolve kernel instantiates it (pack<true>): a binary64 routine behind a non-default switch
// must not shape the register allocation of the render kernels, whose tail calls pack<false>.
__device__ __forceinline__ float srgb_exact(float v)
{
    v = saturate(v);
    if (v < 0.0031308f) return __fmul_rn(v, 12.92f);
    return __fadd_rn(__fmul_rn(1.055f, pm::powf_portable(v, __fdiv_rn(1.0f, 2.4f))), -0.055f);
}
template <bool WITH_EXACT_GAMMA> __device__ __forceinline__ uint32_t quantize(float c, bool exact_aces, bool exact_gamma)
{
    const float x = __fmul_rn(c, 1.0f);  // c_exposure = 1
    const float t = exact_aces ? aces_exact(x) : aces(x);
    float v;
    if constexpr (WITH_EXACT_GAMMA) v = exact_gamma ? srgb_exact(t) : srgb(t);
    else v = srgb(t);
    v = __fmul_rn(saturate(v), 255.f);
    return (uint32_t)__float2int_rn(v) & 0xFFu;
}
// mode bit 0: 0 = file packing A=FF | B<<16 | G<<8 | R (v4.cpp:1321-1325), 1 = screen R<<16 | G<<8 | B (:1285-1289)
// mode bit 1 (B200PT_LDR_EXACT_ACES): the exact ACES curve instead of the fast one; bit 2 (B200PT_LDR_EXACT_GAMMA): pow() gamma,
// honoured by pack<true> only (the render kernels' fused tone map is pack<false>: the host runs the resolve kernel after them
// when that bit is wanted, b200pt_capi.cu)
template <bool WITH_EXACT_GAMMA> __device__ __forceinline__ uint32_t pack(float r, float g, float b, int mode)
{
    const bool ex = (mode & 2) != 0, eg = (mode & 4) != 0;
    const uint32_t R = quantize<WITH_EXACT_GAMMA>(r, ex, eg), G = quantize<WITH_EXACT_GAMMA>(g, ex, eg), B = quantize<WITH_EXACT_GAMMA>(b, ex, eg);
    return ((mode & 1) == 0) ? (0xFF000000u | (B << 16) | (G << 8) | R) : ((R << 16) | (G << 8) | B);
}

}  // namespace tonemap
}  // namespace b200pt
