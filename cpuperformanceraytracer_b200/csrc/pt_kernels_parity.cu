// pt_kernels_parity.cu -- instantiates the megakernel for the ParityMath policy.
// This translation unit is compiled with --fmad=false -prec-div=true -prec-sqrt=true -ftz=false: the only fused operations are the explicit fmaf() calls.
#include "pt_device.cuh"

namespace b200pt {

cudaError_t launch_render_parity(const LaunchConfig& lc, const RenderParams& rp, const SceneSet& scenes, cudaStream_t stream)
{
    return dispatch_config<ParityMath>(lc, [&](auto kernel) -> cudaError_t {
        using KernelT = decltype(kernel);
        if constexpr (std::is_same<KernelT, void (*)(RenderParams, V4Scene)>::value) {
            kernel<<<lc.grid, lc.block, 0, stream>>>(rp, scenes.v4);
        } else if constexpr (std::is_same<KernelT, void (*)(RenderParams, V3RedoScene)>::value) {
            kernel<<<lc.grid, lc.block, 0, stream>>>(rp, scenes.v3redo);
        } else if constexpr (std::is_same<KernelT, void (*)(RenderParams, V3RedoScene0)>::value) {
            kernel<<<lc.grid, lc.block, 0, stream>>>(rp, scenes.v3redo0);
        } else {
            kernel<<<lc.grid, lc.block, 0, stream>>>(rp, scenes.cornell);
        }
        return cudaGetLastError();
    });
}

cudaError_t occupancy_parity(const LaunchConfig& lc, int* blocks_per_sm)
{
    return dispatch_config<ParityMath>(lc, [&](auto kernel) -> cudaError_t {
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, lc.block, 0);
    });
}

// ---- parity hooks for the portable transcendentals (pm_math.cuh) ----------------------------
// op: 0 sin, 1 cos, 2 atan2(a, b), 3 asin, 4 exp -- exactly what ParityMath hands the megakernel; 9 pow(a, b) of the tone map
__global__ void eval_portable_kernel(int op, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float r, t;
        switch (op) {
        case 0: ParityMath::sincos(a[i], &r, &t); break;
        case 1: ParityMath::sincos(a[i], &t, &r); break;
        case 2: r = ParityMath::atan2(a[i], b[i]); break;
        case 3: r = ParityMath::asin(a[i]); break;
        case 9: r = pm::powf_portable(a[i], b[i]); break;
        default: r = ParityMath::exp(a[i]); break;
        }
        out[i] = r;
    }
}

__device__ __forceinline__ float random01_bits(uint32_t h) { return __int2float_rn((int)(h & 0x7FFFFFFFu)) * 4.656612873077392578125e-10f; }

__device__ __forceinline__ uint32_t mix32(uint64_t v)
{
    v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
    return (uint32_t)v;
}

// Compares the two tiers of atan2f_portable / asinf_portable on generated inputs.  asin (op 3): input
// number i is the binary32 value with bit pattern (uint32)i -- first = 0, count = 2^32 is exhaustive.
// sqrt / rcp (ops 5, 6): same enumeration, the unchecked mid-range forms against __fsqrt_rn / __frcp_rn.
// atan2 (op 2): pairs hashed from i; every fourth pair is two arbitrary bit patterns (NaN, infinities,
// denormals included), the others are components of direction-like vectors in [-1, 1], some exactly 0.
// counts[0] = inputs whose tiers disagree (bit patterns compared, NaNs as NaNs), counts[1] = inputs
// the first tier handed to the literal algorithm
__global__ void check_portable_tiers_kernel(int op, unsigned long long first, unsigned long long count, unsigned long long* counts)
{
    unsigned long long bad = 0, second = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = first + i;
        float got, want;
        if (op == 5 || op == 6) {  // sqrt_mid / rcp_mid against the checked forms over their whole valid range
            const uint32_t bits = (uint32_t)k, mag = bits & 0x7fffffffu;
            const float v = __uint_as_float(bits);
            const bool in_range = op == 5 ? (bits >= 0x0d000000u && bits <= 0x7f7fffffu)   // [2^-101, FLT_MAX]
                                          : (mag >= 0x00800000u && mag < 0x7e800000u);    // 2^-126 <= |v| < 2^126
            got = in_range ? (op == 5 ? ParityMath::sqrt_mid(v) : ParityMath::rcp_mid(v)) : 0.f;
            want = in_range ? (op == 5 ? ParityMath::sqrt(v) : ParityMath::rcp(v)) : 0.f;
            second += !in_range;
        } else if (op == 8) {  // texel index of the random-jitter equirect lookup: bracketed fast form vs exact angles
            const uint32_t h0 = mix32(3 * k), h1 = mix32(3 * k + 1), h2 = mix32(3 * k + 2);
            // a direction: normalised like the kernels do, sometimes axis-aligned / near a pole / a seam
            float dx = (float)(int)(h0 >> 8) * (1.f / 8388608.f) - 1.f, dy = (float)(int)(h1 >> 8) * (1.f / 8388608.f) - 1.f;
            float dz = (float)(int)(h2 >> 8) * (1.f / 8388608.f) - 1.f;
            if ((h0 & 0xff) == 0) dx = 0.f;
            if ((h1 & 0xff) == 0) dy = 0.f;
            if ((h2 & 0xff) == 0) dz = 0.f;
            if ((h0 & 0xff) == 1) { dx *= 1e-4f; dz *= 1e-4f; }   // near a pole
            if ((h2 & 0xff) == 1) dz *= 1e-6f;                     // near the seam / the centre column
            const v3 d = normalize3<ParityMath>(mk(dx, dy, dz));
            const float r1 = random01_bits(mix32(k ^ 0x1234567ULL)), r2 = random01_bits(mix32(k ^ 0x89abcdeULL));
            RenderParams rp{};
            const int sizes[4][2] = {{512, 256}, {2048, 1024}, {4096, 2048}, {1000, 333}};
            rp.env_w = sizes[k & 3][0];
            rp.env_h = sizes[k & 3][1];
            const float ux = fract1(fmaf(0.1591f, ParityMath::atan2(d.z, d.x), 0.5f));
            const float uy = fract1(fmaf(0.3183f, ParityMath::asin(d.y), 0.5f));
            const float u = saturate1(ux), v = saturate1(uy);
            const float Row = fmaf(v, (float)rp.env_h, -v), Col = fmaf(u, (float)rp.env_w, -u);
            const int exact = __float2int_rn(fmaf(floorf(Row + r1), (float)rp.env_w, floorf(Col + r2)));
            int fast = 0;
            const bool certain = equirect_random_texel_certain(rp, d, r1, r2, fast);
            // the brackets rest on |approximate angle - exact angle| < kAngleEps: demand a factor 3 of slack
            const float ea = fabsf(atan2_approx(d.z, d.x) - ParityMath::atan2(d.z, d.x));
            float c;
            const float om = fmaf(-d.y, d.y, 1.f);
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(om));
            const float eb = fabsf(atan2_approx(d.y, c) - ParityMath::asin(d.y));
            const bool finite = d.x == d.x && fabsf(d.y) < 1.f && (d.x != 0.f || d.z != 0.f);
            const bool angles_ok = !finite || (ea < kAngleEps / 3.f && eb < kAngleEps / 3.f);
            // the point sampler of texture.cpp:101-139 (simt_textured), same direction and map
            bool point_bad = false;
            {
                float px = ParityMath::atan2(d.z, d.x) * 0.1591f, py = ParityMath::asin(d.y) * 0.3183f;
                px = px + 0.5f;
                py = py + 0.5f;
                int point_exact = -1;
                if (px == px && py == py) {
                    px -= (float)(int)px;
                    py -= (float)(int)py;
                    if (px >= 0.f && px < 1.f && py >= 0.f && py < 1.f)
                        point_exact = (int)(py * (float)(rp.env_h - 1)) * rp.env_w + (int)(px * (float)(rp.env_w - 1));
                }
                int point_fast = 0;
                if (equirect_point_texel_certain(rp, d, point_fast)) point_bad = point_fast != point_exact;
            }
            got = (certain && fast != exact) || !angles_ok || point_bad ? 1.f : 0.f;
            want = 0.f;
            second += !certain;
        } else if (op == 7) {  // div_mid against __fdiv_rn: hashed pairs, |a| in [2^-60, 2^60] or 0, |b| in [2^-60, 2^60]
            const uint32_t h0 = mix32(2 * k), h1 = mix32(2 * k + 1);
            const uint32_t ea = 67u + (h0 >> 8) % 121u, eb = 67u + (h1 >> 8) % 121u;  // exponent fields 67..187
            float a = __uint_as_float((h0 & 0x80000000u) | (ea << 23) | (mix32(k ^ 0x9e3779b97f4a7c15ULL) & 0x7fffffu));
            float b = __uint_as_float((h1 & 0x80000000u) | (eb << 23) | (mix32(k + 0x51ed270b7f4a7c15ULL) & 0x7fffffu));
            if ((k & 7) == 1) b = (k & 8) ? 1.77777779f : 1.33333337f;  // 16:9, 4:3
            if ((h0 & 0xff) == 0) {  // +0 / positive
                a = 0.f;
                b = fabsf(b);
            }
            got = ParityMath::div_mid(a, b, ParityMath::div_mid_reciprocal(b));
            want = ParityMath::div(a, b);
        } else if (op == 3) {
            const float v = __uint_as_float((uint32_t)k);
            got = pm::asinf_portable(v);
            want = pm::asinf_literal(v);
            float unused;
            second += !pm::asinf_first_tier(v, unused);
        } else {
            const uint32_t h0 = mix32(2 * k), h1 = mix32(2 * k + 1);
            float y, x;
            if ((k & 3) == 3) {
                y = __uint_as_float(h0);
                x = __uint_as_float(h1);
            } else {
                y = (float)(int)(h0 >> 8) * (1.f / 8388608.f) - 1.f;
                x = (float)(int)(h1 >> 8) * (1.f / 8388608.f) - 1.f;
                if ((h0 & 0xff) == 0) y = 0.f;
                if ((h1 & 0xff) == 0) x = -0.f;
                if ((h1 & 0xff) == 1) x = y;
            }
            got = pm::atan2f_portable(y, x);
            want = pm::atan2f_literal(y, x);
            float unused;
            second += !pm::atan2f_first_tier(y, x, unused);
        }
        const bool same = (__float_as_uint(got) == __float_as_uint(want)) || (got != got && want != want);
        bad += !same;
    }
    for (int o = 16; o > 0; o >>= 1) {
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        second += __shfl_xor_sync(0xffffffffu, second, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(&counts[0], bad);
        if (second) atomicAdd(&counts[1], second);
    }
}

cudaError_t launch_eval_portable(int op, const float* a, const float* b, float* out, size_t n, cudaStream_t stream)
{
    eval_portable_kernel<<<148 * 8, 256, 0, stream>>>(op, a, b, out, n);
    return cudaGetLastError();
}

cudaError_t launch_check_portable_tiers(int op, unsigned long long first, unsigned long long count, unsigned long long* counts,
                                        cudaStream_t stream)
{
    check_portable_tiers_kernel<<<148 * 8, 256, 0, stream>>>(op, first, count, counts);
    return cudaGetLastError();
}

}  // namespace b200pt
