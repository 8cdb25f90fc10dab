/*
 * pt_oracle.c -- TEST INFRASTRUCTURE (see pt_oracle.h).  Scalar restatement of the reference's
 * AVX2 path tracer: every function below cites the reference lines it follows and performs the
 * same IEEE binary32 operations in the same order (fused only where the reference writes
 * fmadd/fmsub/fnmadd), so that it reproduces the "exact"-mode reference binaries bit for bit.
 * Compile with -ffp-contract=off -mfma (oracle/Makefile).
 *
 * One lane of the reference's 8-wide code = one call here.  A lane whose path has missed keeps
 * executing masked segments in the reference; nothing it computes after the miss can change its
 * result (all write-backs are blended on shouldBreak), so the scalar code returns at the miss.
 */
#include "pt_oracle.h"
#include "portable_math.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef struct { float x, y, z; } v3;

#define FMA(a, b, c) __builtin_fmaf((a), (b), (c))

static const float c_minimumRayHitTime = 0.01f; /* v2.cpp:9, v4.cpp:10 */
static const float c_rayPosNormalNudge = 0.01f; /* v2.cpp:13, v4.cpp:14 */
static const float c_superFar = 10000.0f;       /* v2.cpp:16, v4.cpp:17 */
static const float c_FOVDegrees = 90.0f;        /* v2.cpp:19, v4.cpp:20 */
static const float c_pi = 3.14159265359f;       /* v2.cpp:27, mathutils.h:5 */

/* ---- mathlib.h wrappers ---------------------------------------------------------------- */
static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add3(v3 u, v3 v) { return V3(u.x + v.x, u.y + v.y, u.z + v.z); }  /* :94 */
static inline v3 sub3(v3 u, v3 v) { return V3(u.x - v.x, u.y - v.y, u.z - v.z); }  /* :99 */
static inline v3 mul3(v3 u, v3 v) { return V3(u.x * v.x, u.y * v.y, u.z * v.z); }  /* :104 */
static inline v3 muls(v3 u, float c) { return V3(u.x * c, u.y * c, u.z * c); }     /* :129 */
static inline v3 neg3(v3 u) { return V3(-u.x, -u.y, -u.z); }                       /* :382,716 */
static inline float dot3(v3 u, v3 v) { return FMA(u.x, v.x, FMA(u.y, v.y, u.z * v.z)); } /* :145 */
static inline v3 cross3(v3 u, v3 v)                                                /* :770-778 */
{
    return V3(FMA(u.y, v.z, -(u.z * v.y)), FMA(u.z, v.x, -(u.x * v.z)), FMA(u.x, v.y, -(u.y * v.x)));
}
static inline v3 normalize3(v3 v) { return muls(v, 1.0f / sqrtf(dot3(v, v))); }    /* :759 */
static inline float rcp_exact(float a) { return 1.0f / a; }            /* :417, exact mode */
static inline float rsroot_exact(float a) { return 1.0f / sqrtf(a); }  /* :444, exact mode */
static inline v3 fast_approx_normalize3(v3 v) { return muls(v, rsroot_exact(dot3(v, v))); } /* :755 */
static inline v3 lerp3(v3 u, v3 v, float x) { return add3(u, muls(sub3(v, u), x)); } /* :763 */
static inline float max_ps(float a, float b) { return a > b ? a : b; } /* :360, x86 maxps */
static inline float min_ps(float a, float b) { return a < b ? a : b; } /* :365, x86 minps */
static inline float saturate1(float x) { return min_ps(max_ps(x, 0.f), 1.f); }     /* :405,410 */
static inline v3 saturate3(v3 v) { return V3(saturate1(v.x), saturate1(v.y), saturate1(v.z)); }
static inline float fract1(float a) { return a - floorf(a); }                       /* :400 */
static inline int to_epi32(float a) { return (int)lrintf(a); } /* :863 cvtps2dq, RN-even */
static inline float approx_exp1(float a)                                            /* :501-516 */
{
    float b = FMA(a, 0.05995203836930455f, 1.f);
    float b2 = b * b, b4 = b2 * b2, b8 = b4 * b4;
    return b8 * b8;
}

/* ---- RNG: mathutils.h:8-26 --------------------------------------------------------------- */
uint32_t oracle_wang_hash(uint32_t* s)
{
    uint32_t v = *s;
    v = (v ^ 61u) ^ (v >> 16);
    v = v * 9u;
    v = v ^ (v >> 4);
    v = v * 0x27d4eb2du;
    v = v ^ (v >> 15);
    *s = v;
    return v;
}
float oracle_random01(uint32_t* s)
{
    return (float)(int32_t)(oracle_wang_hash(s) & 0x7FFFFFFFu) / 2147483648.0f;
}
#define RAND01(s) oracle_random01(s)

/* v4.cpp:1096-1101, v2.cpp:530-538 */
uint32_t oracle_seed(int x, int y_flipped, int frame)
{
    return ((uint32_t)x * 1973u + (uint32_t)y_flipped * 9277u + (uint32_t)frame * 26699u) | 1u;
}

float oracle_camera_distance(void)
{
    return 1.0f / tanf(c_FOVDegrees * 0.5f * c_pi / 180.0f); /* v2.cpp:546, v4.cpp:1500 */
}

/* mathutils.h:33-47 == v2.cpp:76-88 (cos_ps/sin_ps vs sincos_ps: same values) */
static v3 RandomUnitVector(uint32_t* state)
{
    const float c_twopi = 2.0f * c_pi;
    float wide_z = RAND01(state);
    float wide_a = RAND01(state);
    float z = wide_z * 2.f - 1.f;
    float a = wide_a * c_twopi;
    float r = sqrtf(1.f - z * z);
    float s, c;
    pm_sincosf(a, &s, &c);
    return V3(r * c, r * s, z);
}

/* ---- hit record --------------------------------------------------------------------------- */
typedef struct {
    float dist;
    v3 normal;
    int fromInside;
    int matIndex; /* v4: object index; legacy: index into the Cornell material table */
} hit_t;

/* ---- legacy quad / sphere tests: v2.cpp:159-317 == simt_textured.cpp:117-275 ----------- */
static int TestQuadTrace_legacy(v3 rayPos, v3 rayDir, hit_t* info, v3 a, v3 b, v3 c, v3 d)
{
    int early_return = 0;
    v3 normal = normalize3(cross3(sub3(c, a), sub3(c, b)));
    if (dot3(normal, rayDir) > 0.f) {
        normal = muls(normal, -1.0f);
        v3 t = d; d = a; a = t;
        t = b; b = c; c = t;
    }
    v3 p = rayPos;
    v3 q = add3(rayPos, rayDir);
    v3 pq = sub3(q, p);
    v3 pa = sub3(a, p), pb = sub3(b, p), pc = sub3(c, p);
    v3 m = cross3(pc, pq);
    float v = dot3(pa, m);
    v3 intersectPos;
    if (v >= 0.f) {
        float u = -dot3(pb, m);
        if (u < 0.f) early_return = 1;
        float w = dot3(cross3(pq, pb), pa);
        if (w < 0.f) early_return = 1;
        float denom = 1.0f / (u + v + w);
        u = u * denom;
        v = v * denom;
        w = w * denom;
        intersectPos = add3(add3(muls(a, u), muls(b, v)), muls(c, w));
    } else {
        v3 pd = sub3(d, p);
        float u = dot3(pd, m);
        if (u < 0.f) early_return = 1;
        float w = dot3(cross3(pq, pa), pd);
        if (w < 0.f) early_return = 1;
        v = -v;
        float denom = 1.0f / (u + v + w);
        u = u * denom;
        v = v * denom;
        w = w * denom;
        intersectPos = add3(add3(muls(a, u), muls(d, v)), muls(c, w));
    }
    float dist = (intersectPos.z - rayPos.z) / rayDir.z;
    if (fabsf(rayDir.y) > 0.f) dist = (intersectPos.y - rayPos.y) / rayDir.y;
    if (fabsf(rayDir.x) > 0.f) dist = (intersectPos.x - rayPos.x) / rayDir.x;
    if (!early_return && dist > c_minimumRayHitTime && dist < info->dist) {
        info->fromInside = 0; /* v3_redo.cpp:305 (the v2 struct has no such field) */
        info->dist = dist;
        info->normal = normal;
        return 1;
    }
    return 0;
}

static int TestSphereTrace_legacy(v3 rayPos, v3 rayDir, hit_t* info, v3 center, float radius)
{
    v3 m = sub3(rayPos, center);
    float b = dot3(m, rayDir);
    float c = dot3(m, m) - radius * radius;
    int early_return = (c > 0.f && b > 0.f);
    float discr = b * b - c;
    if (discr < 0.f) early_return = 1;
    float sq = sqrtf(discr);
    float dist = -b - sq;
    int fromInside = dist < 0.f;
    if (fromInside) dist = -b + sq;
    if (!early_return && dist > c_minimumRayHitTime && dist < info->dist) {
        info->fromInside = fromInside; /* v3_redo.cpp:366 */
        info->dist = dist;
        v3 n = normalize3(sub3(add3(rayPos, muls(rayDir, dist)), center));
        info->normal = muls(n, fromInside ? -1.0f : 1.0f);
        return 1;
    }
    return 0;
}

/* Cornell scene: v2.cpp:320-454 (materials) / simt_textured.cpp:278-385 (albedo+emissive only) */
typedef struct { v3 albedo, emissive, specularColor; float percentSpecular, roughness; } legacy_mat_t;
typedef struct {
    v3 quad[6][4];
    v3 sphereCenter[3];
    float sphereRadius[3];
    legacy_mat_t mat[9];
} cornell_t;

static void cornell_init(cornell_t* s, int profile)
{
    const v3 T = V3(0.0f, 0.0f, 10.0f);
    static const float Q[6][4][3] = {
        {{-12.6f, -12.6f, 25.0f}, {12.6f, -12.6f, 25.0f}, {12.6f, 12.6f, 25.0f}, {-12.6f, 12.6f, 25.0f}},     /* back wall */
        {{-12.6f, -12.45f, 25.0f}, {12.6f, -12.45f, 25.0f}, {12.6f, -12.45f, 15.0f}, {-12.6f, -12.45f, 15.0f}}, /* floor */
        {{-12.6f, 12.5f, 25.0f}, {12.6f, 12.5f, 25.0f}, {12.6f, 12.5f, 15.0f}, {-12.6f, 12.5f, 15.0f}},       /* ceiling */
        {{-12.5f, -12.6f, 25.0f}, {-12.5f, -12.6f, 15.0f}, {-12.5f, 12.6f, 15.0f}, {-12.5f, 12.6f, 25.0f}},   /* left wall */
        {{12.5f, -12.6f, 25.0f}, {12.5f, -12.6f, 15.0f}, {12.5f, 12.6f, 15.0f}, {12.5f, 12.6f, 25.0f}},       /* right wall */
        {{-5.0f, 12.4f, 22.5f}, {5.0f, 12.4f, 22.5f}, {5.0f, 12.4f, 17.5f}, {-5.0f, 12.4f, 17.5f}}};          /* light */
    for (int i = 0; i < 6; i++)
        for (int k = 0; k < 4; k++) s->quad[i][k] = add3(V3(Q[i][k][0], Q[i][k][1], Q[i][k][2]), T);
    static const float SX[3] = {-9.0f, 0.0f, 9.0f};
    for (int i = 0; i < 3; i++) {
        s->sphereCenter[i] = add3(V3(SX[i], -9.5f, 20.0f), T);
        s->sphereRadius[i] = 3.0f + 0.0f;
    }
    memset(s->mat, 0, sizeof(s->mat));
    s->mat[0].albedo = V3(0.7f, 0.7f, 0.7f);
    s->mat[1].albedo = V3(0.7f, 0.7f, 0.7f);
    s->mat[2].albedo = V3(0.7f, 0.7f, 0.7f);
    s->mat[3].albedo = V3(0.7f, 0.1f, 0.1f);
    s->mat[4].albedo = V3(0.1f, 0.7f, 0.1f);
    s->mat[5].emissive = muls(V3(1.0f, 0.9f, 0.7f), 20.0f);
    if (profile == ORACLE_PROFILE_V2) {
        s->mat[6].albedo = V3(0.9f, 0.9f, 0.5f); s->mat[6].percentSpecular = 0.1f; s->mat[6].roughness = 0.2f; s->mat[6].specularColor = V3(0.9f, 0.9f, 0.9f);
        s->mat[7].albedo = V3(0.9f, 0.5f, 0.9f); s->mat[7].percentSpecular = 0.3f; s->mat[7].roughness = 0.2f; s->mat[7].specularColor = V3(0.9f, 0.9f, 0.9f);
        s->mat[8].albedo = V3(0.f, 0.f, 1.f);    s->mat[8].percentSpecular = 0.5f; s->mat[8].roughness = 0.4f; s->mat[8].specularColor = V3(1.f, 0.f, 0.f);
    } else {
        s->mat[6].albedo = V3(0.9f, 0.9f, 0.75f);
        s->mat[7].albedo = V3(0.9f, 0.75f, 0.9f);
        s->mat[8].albedo = V3(0.9f, 0.75f, 0.9f);
    }
}

static void cornell_from(cornell_t* s, const oracle_scene_cornell* src)
{
    for (int i = 0; i < 6; i++)
        for (int k = 0; k < 4; k++) s->quad[i][k] = V3(src->quad_vertices[12 * i + 3 * k], src->quad_vertices[12 * i + 3 * k + 1], src->quad_vertices[12 * i + 3 * k + 2]);
    for (int i = 0; i < 3; i++) {
        s->sphereCenter[i] = V3(src->spheres[4 * i], src->spheres[4 * i + 1], src->spheres[4 * i + 2]);
        s->sphereRadius[i] = src->spheres[4 * i + 3];
    }
    memset(s->mat, 0, sizeof(s->mat));
    for (int i = 0; i < 9; i++) {
        const float* m = src->materials + 11 * i;
        s->mat[i].albedo = V3(m[0], m[1], m[2]);
        s->mat[i].emissive = V3(m[3], m[4], m[5]);
        s->mat[i].specularColor = V3(m[6], m[7], m[8]);
        s->mat[i].percentSpecular = m[9];
        s->mat[i].roughness = m[10];
    }
}

static void TestSceneTrace_cornell(const cornell_t* s, v3 rayPos, v3 rayDir, hit_t* h)
{
    for (int i = 0; i < 6; i++)
        if (TestQuadTrace_legacy(rayPos, rayDir, h, s->quad[i][0], s->quad[i][1], s->quad[i][2], s->quad[i][3]))
            h->matIndex = i;
    for (int i = 0; i < 3; i++)
        if (TestSphereTrace_legacy(rayPos, rayDir, h, s->sphereCenter[i], s->sphereRadius[i]))
            h->matIndex = 6 + i;
}

/* ---- env samplers: texture.cpp ----------------------------------------------------------- */
typedef struct { const float* data; int W, H; } tex_t;

static v3 fetch_flat(const tex_t* t, int idx) /* GatherRGB, texture.cpp:16-27 */
{
    int64_t last = (int64_t)t->W * t->H * 3 - 3;
    if (idx < 0) idx = 0;          /* the reference would read out of bounds here; keep the */
    if (idx > last) idx = (int)last; /* restatement memory-safe (unreachable with valid uv) */
    return V3(t->data[idx], t->data[idx + 1], t->data[idx + 2]);
}

/* texture.cpp:101-139, one lane */
static v3 EquirectSamplePoint(const tex_t* t, v3 d)
{
    float ux = pm_atan2f(d.z, d.x), uy = pm_asinf(d.y);
    ux = ux * 0.1591f;
    uy = uy * 0.3183f;
    ux = ux + 0.5f;
    uy = uy + 0.5f;
    if (ux != ux || uy != uy) return V3(0.f, 0.f, 0.f);
    ux -= (float)(int)ux;
    uy -= (float)(int)uy;
    if (ux >= 0.f && ux < 1.f && uy >= 0.f && uy < 1.f) {
        int Row = (int)(uy * (float)(t->H - 1));
        int Col = (int)(ux * (float)(t->W - 1));
        const float* px = t->data + 3 * ((int64_t)Row * t->W + Col);
        return V3(px[0], px[1], px[2]);
    }
    return V3(0.f, 0.f, 0.f);
}

/* texture.cpp:39-76 */
static v3 TexelSampleBilinear(const tex_t* t, float u, float v)
{
    float Row = v * (float)(t->H - 1);
    float Col = u * (float)(t->W - 1);
    float Row0 = floorf(Row), Row1 = ceilf(Row), Col0 = floorf(Col), Col1 = ceilf(Col);
    float dV = Row - Row0, dU = Col - Col0;
    float texWidth = 3.0f * (float)t->W;
    Row0 = Row0 * texWidth;
    Row1 = Row1 * texWidth;
    Col0 = Col0 * 3.0f;
    Col1 = Col1 * 3.0f;
    v3 C00 = fetch_flat(t, to_epi32(Col0 + Row0));
    v3 C10 = fetch_flat(t, to_epi32(Col1 + Row0));
    v3 C01 = fetch_flat(t, to_epi32(Col0 + Row1));
    v3 C11 = fetch_flat(t, to_epi32(Col1 + Row1));
    v3 C0 = lerp3(C00, C10, dU);
    v3 C1 = lerp3(C01, C11, dU);
    return lerp3(C0, C1, dV);
}

/* texture.cpp:78-86 */
static v3 TexelSampleRandom(const tex_t* t, float u, float v, uint32_t* state)
{
    float Row = FMA(v, (float)t->H, -v);
    float Col = FMA(u, (float)t->W, -u);
    float RandRow = floorf(Row + RAND01(state));
    float RandCol = floorf(Col + RAND01(state));
    int idx = 3 * to_epi32(FMA(RandRow, (float)t->W, RandCol));
    return fetch_flat(t, idx);
}

/* texture.cpp:164-184 */
static v3 EquirectSampleBilinear(const tex_t* t, v3 d)
{
    float ux = pm_atan2f(d.z, d.x), uy = pm_asinf(d.y);
    ux = ux * 0.1591f;
    uy = uy * 0.3183f;
    ux = ux + 0.5f;
    uy = uy + 0.5f;
    ux -= floorf(ux);
    uy -= floorf(uy);
    return TexelSampleBilinear(t, saturate1(ux), saturate1(uy));
}

/* texture.cpp:186-203 */
static v3 EquirectSampleRandom(const tex_t* t, v3 d, uint32_t* state)
{
    float ux = fract1(FMA(0.1591f, pm_atan2f(d.z, d.x), 0.5f));
    float uy = fract1(FMA(0.3183f, pm_asinf(d.y), 0.5f));
    return TexelSampleRandom(t, saturate1(ux), saturate1(uy), state);
}

/* face selection shared by texture.cpp:275-339 and :341-404 */
static void cubemap_face(v3 D, float sixth, float* fu, float* fv, float* vOffset, float* maxAbs)
{
    v3 a = V3(fabsf(D.x), fabsf(D.y), fabsf(D.z));
    int xpos = D.x >= 0.f;
    float u = xpos ? -D.z : D.z, v = D.y;
    float off = xpos ? 0.f : sixth;
    {
        int ypos = D.y >= 0.f;
        float yo = ypos ? 2.f * sixth : 3.f * sixth;
        if (a.y >= a.x) { off = yo; u = D.x; v = ypos ? -D.z : D.z; }
    }
    {
        int zpos = D.z >= 0.f;
        float zo = zpos ? 4.f * sixth : 5.f * sixth;
        if (a.z >= a.x && a.z >= a.y) { off = zo; u = zpos ? D.x : -D.x; v = D.y; }
    }
    *fu = u; *fv = v; *vOffset = off;
    *maxAbs = max_ps(a.x, max_ps(a.y, a.z));
}

/* texture.cpp:275-339 */
static v3 CubemapSampleBilinear(const tex_t* t, v3 D)
{
    float fu, fv, off, mx;
    cubemap_face(D, 1.f / 6.f, &fu, &fv, &off, &mx);
    /* offsets are 0, 1.f/6.f, 2.f/6.f ... in the bilinear variant (:287-315) */
    {
        v3 a = V3(fabsf(D.x), fabsf(D.y), fabsf(D.z));
        off = (D.x >= 0.f) ? 0.f : 1.f / 6.f;
        if (a.y >= a.x) off = (D.y >= 0.f) ? 2.f / 6.f : 3.f / 6.f;
        if (a.z >= a.x && a.z >= a.y) off = (D.z >= 0.f) ? 4.f / 6.f : 5.f / 6.f;
    }
    float su = fu / mx, sv = fv / mx;
    float pu = saturate1(su * 0.5f + 0.5f), pv = saturate1(sv * 0.5f + 0.5f);
    float v = saturate1(FMA(pv, 1.f / 6.f, off));
    return TexelSampleBilinear(t, pu, v);
}

/* texture.cpp:341-404 */
static v3 CubemapSampleRandom(const tex_t* t, v3 D, uint32_t* state)
{
    const float sixth = 0.166666666666667f;
    float fu, fv, off, mx;
    cubemap_face(D, sixth, &fu, &fv, &off, &mx);
    float r = rcp_exact(mx);
    float pu = saturate1(FMA(fu * r, 0.5f, 0.5f)), pv = saturate1(FMA(fv * r, 0.5f, 0.5f));
    float v = saturate1(FMA(pv, sixth, off));
    return TexelSampleRandom(t, pu, v, state);
}

/* ---- v4 scene: v4.cpp:248-319 (quads), :330-349 (materials), :1403-1496 (data) ------------ */
typedef struct { v3 V0, NxV01, NxV20, NxV02, NxV30, normal; } quad4_t;
typedef struct {
    v3 albedo, emissive, specularColor, refractionColor;
    float specularChance, specularRoughness, IOR, refractionChance, refractionRoughness;
} mat4_t;
#define MAX_OBJECTS 12 /* v4.cpp:327-328 */
typedef struct {
    int numQuads, numSpheres;
    quad4_t quad[MAX_OBJECTS];
    v3 sphereCenter[MAX_OBJECTS];
    float sphereRadius[MAX_OBJECTS];
    mat4_t mat[MAX_OBJECTS];
    v3 cameraPosition;
    float cameraDistance;
} scene4_t;

static v3 div3s(v3 u, float c) { return V3(u.x / c, u.y / c, u.z / c); }

/* PrecomputeQuadData, v4.cpp:269-319 */
static void quad4_init(quad4_t* q, v3 V0, v3 V1, v3 V2, v3 V3_)
{
    v3 V01 = sub3(V1, V0), V02 = sub3(V2, V0), V30 = sub3(V0, V3_);
    v3 V20 = neg3(V02);
    v3 V01xV02 = cross3(V01, V02);
    v3 V02xV03 = cross3(V30, V01);
    v3 N = normalize3(V01xV02);
    float DetTop = dot3(V02xV03, N);
    float DetBot = dot3(V01xV02, N);
    q->V0 = V0;
    q->normal = N;
    q->NxV01 = div3s(cross3(N, V01), DetBot);
    q->NxV20 = div3s(cross3(N, V20), DetBot);
    q->NxV02 = div3s(cross3(N, V02), DetTop);
    q->NxV30 = div3s(cross3(N, V30), DetTop);
}

static void scene4_init(scene4_t* s)
{
    const v3 T = V3(0.0f, 0.0f, 10.0f);
    memset(s, 0, sizeof(*s));
    s->numQuads = 4;
    s->numSpheres = 7;
    quad4_init(&s->quad[0], add3(V3(-25.0f, -12.5f, 5.0f), T), add3(V3(25.0f, -12.5f, 5.0f), T),
               add3(V3(25.0f, -12.5f, -5.0f), T), add3(V3(-25.0f, -12.5f, -5.0f), T));
    quad4_init(&s->quad[1], V3(-25.0f, -1.5f, 5.0f), V3(25.0f, -1.5f, 5.0f), V3(25.0f, -10.5f, 5.0f),
               V3(-25.0f, -10.5f, 5.0f)); /* no translation, v4.cpp:1430-1433 */
    quad4_init(&s->quad[2], add3(V3(-7.5f, 12.5f, 5.0f), T), add3(V3(7.5f, 12.5f, 5.0f), T),
               add3(V3(7.5f, 12.5f, -5.0f), T), add3(V3(-7.5f, 12.5f, -5.0f), T));
    quad4_init(&s->quad[3], add3(V3(-5.0f, 12.4f, 2.5f), T), add3(V3(5.0f, 12.4f, 2.5f), T),
               add3(V3(5.0f, 12.4f, -2.5f), T), add3(V3(-5.0f, 12.4f, -2.5f), T));
    /* AddMaterialToScene stores albedo.x in all three channels, v4.cpp:1370-1372 */
    s->mat[0].albedo = V3(0.7f, 0.7f, 0.7f);
    s->mat[1].albedo = V3(.35f, .35f, .35f);
    s->mat[2].albedo = V3(0.7f, 0.7f, 0.7f);
    s->mat[3].emissive = muls(V3(1.0f, 0.9f, 0.7f), 20.0f);
    for (int i = 0; i < 7; i++) {
        s->sphereCenter[i] = add3(V3(-18.0f + 6.0f * (float)i, -8.0f, 0.0f), T);
        s->sphereRadius[i] = 2.8f + 0.0f;
        mat4_t* m = &s->mat[4 + i];
        float r = (((float)i) / (float)(7 - 1)) * 0.5f;
        m->specularChance = 0.02f;
        m->IOR = 1.1f;
        m->refractionChance = 1.0f;
        m->albedo = V3(0.9f, 0.9f, 0.9f);
        m->refractionColor = V3(0.0f, 0.5f, 1.0f);
        m->specularColor = muls(V3(1.0f, 1.0f, 1.0f), 0.8f);
        m->specularRoughness = r;
        m->refractionRoughness = r;
    }
    s->cameraDistance = oracle_camera_distance();
    s->cameraPosition = V3(0.f, 0.f, 1.f * 40.f); /* v4.cpp:1501 */
}

/* AddQuadObjectToScene / AddSphereObjectToScene / AddMaterialToScene (v4.cpp:1368-1401) on caller data */
static int scene4_from(scene4_t* s, const oracle_scene_v4* src)
{
    int nq = src->num_quads, ns = src->num_spheres;
    if (nq < 0 || ns < 0 || nq + ns < 1 || nq + ns > MAX_OBJECTS || !src->materials) return -1;
    memset(s, 0, sizeof(*s));
    s->numQuads = nq;
    s->numSpheres = ns;
    for (int i = 0; i < nq; i++) {
        const float* v = src->quad_vertices + 12 * i;
        quad4_init(&s->quad[i], V3(v[0], v[1], v[2]), V3(v[3], v[4], v[5]), V3(v[6], v[7], v[8]), V3(v[9], v[10], v[11]));
    }
    for (int i = 0; i < ns; i++) {
        s->sphereCenter[i] = V3(src->spheres[4 * i], src->spheres[4 * i + 1], src->spheres[4 * i + 2]);
        s->sphereRadius[i] = src->spheres[4 * i + 3];
    }
    for (int i = 0; i < nq + ns; i++) {
        const float* m = src->materials + 17 * i;
        mat4_t* d = &s->mat[i];
        d->albedo = V3(m[0], m[0], m[0]); /* albedo.x three times, v4.cpp:1370-1372 */
        d->emissive = V3(m[3], m[4], m[5]);
        d->specularChance = m[6];
        d->specularRoughness = m[7];
        d->specularColor = V3(m[8], m[9], m[10]);
        d->IOR = m[11];
        d->refractionChance = m[12];
        d->refractionRoughness = m[13];
        d->refractionColor = V3(m[14], m[15], m[16]);
    }
    s->cameraPosition = V3(src->camera_position[0], src->camera_position[1], src->camera_position[2]);
    s->cameraDistance = src->camera_distance;
    return 0;
}

/* v4.cpp:575-645 */
static int TestQuadTrace_v4(v3 rayPos, v3 rayDir, hit_t* info, const quad4_t* q)
{
    v3 normal = q->normal;
    v3 rayOffset = sub3(q->V0, rayPos);
    float rayDirDotN = dot3(rayDir, normal);
    float rayOffsetDotN = dot3(rayOffset, normal);
    float dist = rayOffsetDotN * rcp_exact(rayDirDotN);
    v3 hit = V3(FMA(dist, rayDir.x, -rayOffset.x), FMA(dist, rayDir.y, -rayOffset.y), FMA(dist, rayDir.z, -rayOffset.z));
    float A0 = dot3(hit, q->NxV01), A1 = dot3(hit, q->NxV20), A2 = 1.0f - A0 - A1;
    float B0 = dot3(hit, q->NxV30), B1 = dot3(hit, q->NxV02), B2 = 1.0f - B0 - B1;
    int tri1 = (A0 >= 0.f) && (A1 >= 0.f) && (A2 >= 0.f);
    int tri2 = (B0 >= 0.f) && (B1 >= 0.f) && (B2 >= 0.f);
    if ((tri1 || tri2) && dist > c_minimumRayHitTime && dist < info->dist) {
        info->fromInside = 0;
        info->dist = dist;
        if (dot3(normal, rayDir) > 0.f) info->normal = neg3(normal); /* front-face hits keep the old normal, :639 */
        return 1;
    }
    return 0;
}

/* v4.cpp:649-695 */
static int TestSphereTrace_v4(v3 rayPos, v3 rayDir, hit_t* info, v3 center, float radius)
{
    v3 m = sub3(rayPos, center);
    float b = dot3(m, rayDir);
    float c = FMA(-radius, radius, dot3(m, m));
    int cond = (c > 0.f && b > 0.f);
    float discr = FMA(b, b, -c);
    int early_return = (discr < 0.f) || cond;
    if (early_return) return 0;
    float sroot_discr = sqrtf(discr);
    int fromInside = (-b < sroot_discr);
    float dist = (fromInside ? sroot_discr : -sroot_discr) - b;
    if (dist > c_minimumRayHitTime && dist < info->dist) {
        info->fromInside = fromInside;
        info->dist = dist;
        v3 n = normalize3(V3(FMA(rayDir.x, dist, m.x), FMA(rayDir.y, dist, m.y), FMA(rayDir.z, dist, m.z)));
        info->normal = muls(n, fromInside ? -1.0f : 1.0f);
        return 1;
    }
    return 0;
}

/* v4.cpp:429-453 */
static float FresnelReflectAmount(float n1, float n2, v3 normal, v3 incident, float f0, float f90)
{
    float r0 = (n1 - n2) * rcp_exact(n1 + n2);
    r0 = r0 * r0;
    float cosX = -dot3(normal, incident);
    int cond = n1 > n2;
    float n = n1 * rcp_exact(n2);
    float sinT2Compl = FMA(-(n * n), FMA(-cosX, cosX, 1.f), 1.f);
    float newCosX = sqrtf(sinT2Compl);
    int tir = 0.f > sinT2Compl;
    if (cond && !tir) cosX = newCosX;
    float x = 1.f - cosX;
    float x2 = x * x;
    float ret = FMA((1.f - r0) * x2 * x2, x, r0);
    if (cond && tir) ret = 1.f;
    return FMA(ret, f90 - f0, f0);
}

/* mathlib.h:781-789 */
static v3 rfrct(v3 v, v3 n, float ior)
{
    float vdotn = dot3(v, n);
    float k = FMA(-ior, ior * FMA(-vdotn, vdotn, 1.f), 1.f);
    if (k < 0.f) return V3(0.f, 0.f, 0.f);
    float t = FMA(ior, vdotn, sqrtf(k));
    return V3(FMA(ior, v.x, -(t * n.x)), FMA(ior, v.y, -(t * n.y)), FMA(ior, v.z, -(t * n.z)));
}

/* v4.cpp:109-129 */
static v3 RandomUnitVectorRejectionSample(uint32_t* state)
{
    float u = FMA(2.0f, RAND01(state), -1.f);
    float v = FMA(2.0f, RAND01(state), -1.f);
    float w = FMA(2.0f, RAND01(state), -1.f);
    float uv_d2 = FMA(u, u, v * v);
    float uvw_d2 = FMA(w, w, uv_d2);
    return muls(V3(u, v, w), rsroot_exact(uvw_d2));
}

/* ---- v3_redo: v3_redo.cpp:195-219 (Fresnel), :379-602 (scenes: SCENE 1 = the v4 geometry, the checked-in choice;
 * ---- SCENE 0 = a Cornell box with Fresnel-specular spheres and a light outside the box), :607-754 (shading) ---- */
typedef struct {
    int nquads, nspheres;
    int backdrop;          /* index of the striped backdrop quad (albedo computed at the hit), -1: none */
    v3 quad[6][4];
    v3 sphereCenter[7];
    float sphereRadius[7];
    mat4_t mat[13];        /* quads first, then spheres */
} scene3_t;

static void scene3_init(scene3_t* s, int scene)
{
    memset(s, 0, sizeof(*s));
    for (int i = 0; i < 13; i++) s->mat[i].IOR = 1.f; /* GetZeroedMaterial, :155-168 */
    if (scene == 0) { /* :392-479, :530-580; sceneTranslation = 0 (:387) */
        const v3 T = V3(0.0f, 0.0f, 0.0f);
        static const float Q[6][4][3] = {
            {{-12.6f, -12.6f, 25.0f}, {12.6f, -12.6f, 25.0f}, {12.6f, 12.6f, 25.0f}, {-12.6f, 12.6f, 25.0f}},        /* back wall */
            {{-12.6f, -12.45f, 25.0f}, {12.6f, -12.45f, 25.0f}, {12.6f, -12.45f, 15.0f}, {-12.6f, -12.45f, 15.0f}},  /* floor */
            {{-12.6f, 12.5f, 25.0f}, {12.6f, 12.5f, 25.0f}, {12.6f, 12.5f, 15.0f}, {-12.6f, 12.5f, 15.0f}},          /* ceiling */
            {{-12.5f, -12.6f, 25.0f}, {-12.5f, -12.6f, 15.0f}, {-12.5f, 12.6f, 15.0f}, {-12.5f, 12.6f, 25.0f}},      /* left wall */
            {{12.5f, -12.6f, 25.0f}, {12.5f, -12.6f, 15.0f}, {12.5f, 12.6f, 15.0f}, {12.5f, 12.6f, 25.0f}},          /* right wall */
            {{-5.0f, 12.4f, -22.5f}, {5.0f, 12.4f, -22.5f}, {5.0f, 12.4f, -17.5f}, {-5.0f, 12.4f, -17.5f}}};         /* light */
        s->nquads = 6; s->nspheres = 3; s->backdrop = -1;
        for (int i = 0; i < 6; i++)
            for (int k = 0; k < 4; k++) s->quad[i][k] = add3(V3(Q[i][k][0], Q[i][k][1], Q[i][k][2]), T);
        s->mat[0].albedo = V3(0.7f, 0.7f, 0.7f);
        s->mat[1].albedo = V3(0.7f, 0.7f, 0.7f);
        s->mat[2].albedo = V3(0.7f, 0.7f, 0.7f);
        s->mat[3].albedo = V3(0.7f, 0.1f, 0.1f);
        s->mat[4].albedo = V3(0.1f, 0.7f, 0.1f);
        s->mat[5].emissive = muls(V3(1.0f, 0.9f, 0.7f), 20.0f);
        static const float C[3] = {-9.0f, 0.0f, 9.0f};
        for (int i = 0; i < 3; i++) {
            s->sphereCenter[i] = add3(V3(C[i], -9.5f, 20.0f), T);
            s->sphereRadius[i] = 3.0f + 0.0f;
        }
        s->mat[6].albedo = V3(0.9f, 0.9f, 0.5f); s->mat[6].specularChance = 0.1f; s->mat[6].specularRoughness = 0.2f; s->mat[6].specularColor = V3(0.9f, 0.9f, 0.9f);
        s->mat[7].albedo = V3(0.9f, 0.5f, 0.9f); s->mat[7].specularChance = 0.3f; s->mat[7].specularRoughness = 0.2f; s->mat[7].specularColor = V3(0.9f, 0.9f, 0.9f);
        s->mat[8].albedo = V3(0.f, 0.f, 1.f);    s->mat[8].specularChance = 0.5f; s->mat[8].specularRoughness = 0.4f; s->mat[8].specularColor = V3(1.f, 0.f, 0.f);
        return;
    }
    const v3 T = V3(0.0f, 0.0f, 10.0f);
    static const float Q[4][4][3] = {
        {{-25.0f, -12.5f, 5.0f}, {25.0f, -12.5f, 5.0f}, {25.0f, -12.5f, -5.0f}, {-25.0f, -12.5f, -5.0f}},
        {{-25.0f, -1.5f, 5.0f}, {25.0f, -1.5f, 5.0f}, {25.0f, -10.5f, 5.0f}, {-25.0f, -10.5f, 5.0f}},
        {{-7.5f, 12.5f, 5.0f}, {7.5f, 12.5f, 5.0f}, {7.5f, 12.5f, -5.0f}, {-7.5f, 12.5f, -5.0f}},
        {{-5.0f, 12.4f, 2.5f}, {5.0f, 12.4f, 2.5f}, {5.0f, 12.4f, -2.5f}, {-5.0f, 12.4f, -2.5f}}};
    s->nquads = 4; s->nspheres = 7; s->backdrop = 1;
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 4; k++) {
            v3 p = V3(Q[i][k][0], Q[i][k][1], Q[i][k][2]);
            s->quad[i][k] = (i == 1) ? p : add3(p, T); /* the backdrop is not translated, :505-508 */
        }
    s->mat[0].albedo = V3(0.7f, 0.7f, 0.7f);
    s->mat[2].albedo = V3(0.7f, 0.7f, 0.7f);
    s->mat[3].emissive = muls(V3(1.0f, 0.9f, 0.7f), 20.0f);
    for (int i = 0; i < 7; i++) {
        s->sphereCenter[i] = add3(V3(-18.0f + 6.0f * (float)i, -8.0f, 0.0f), T);
        s->sphereRadius[i] = 2.8f + 0.0f;
        mat4_t* m = &s->mat[4 + i];
        float r = ((float)i / (float)(7 - 1)) * 0.5f;
        m->albedo = V3(0.9f, 0.25f, 0.25f);
        m->specularChance = 0.02f;
        m->specularRoughness = r;
        m->specularColor = muls(V3(1.0f, 1.0f, 1.0f), 0.8f);
        m->IOR = 1.1f;
        m->refractionChance = 1.0f;
        m->refractionRoughness = r;
        m->refractionColor = V3(0.0f, 0.5f, 1.0f);
    }
}

/* ---- per-path radiance --------------------------------------------------------------------- */
typedef struct {
    const oracle_params* p;
    cornell_t cornell;
    scene4_t scene4;
    scene3_t scene3;
    tex_t tex;
} ctx_t;

typedef struct { uint64_t segments, escapes; } path_stats_t;

/* v2.cpp:456-524 */
static v3 GetColorForRay_v2(const ctx_t* c, v3 rayPos, v3 rayDir, uint32_t* rng, path_stats_t* st)
{
    v3 ret = V3(0.f, 0.f, 0.f), throughput = V3(1.f, 1.f, 1.f);
    for (int bounceIndex = 0; bounceIndex <= c->p->num_bounces; ++bounceIndex) {
        hit_t h;
        h.dist = c_superFar; h.normal = V3(0.f, 0.f, 0.f); h.fromInside = 0; h.matIndex = -1;
        st->segments++;
        TestSceneTrace_cornell(&c->cornell, rayPos, rayDir, &h);
        if (h.dist == c_superFar) {
            v3 ambient = mul3(V3(.11f, .1f, .15f), throughput);
            st->escapes++;
            return add3(ret, ambient);
        }
        const legacy_mat_t* m = &c->cornell.mat[h.matIndex];
        rayPos = add3(add3(rayPos, muls(rayDir, h.dist)), muls(h.normal, c_rayPosNormalNudge));
        float doSpecular = (RAND01(rng) < m->percentSpecular) ? 1.f : 0.f;
        v3 diffuseRayDir = normalize3(add3(h.normal, RandomUnitVector(rng)));
        v3 specularRayDir = sub3(rayDir, muls(muls(h.normal, 2.f), dot3(rayDir, h.normal)));
        float roughnessSqrd = m->roughness * m->roughness;
        specularRayDir = normalize3(lerp3(specularRayDir, diffuseRayDir, roughnessSqrd));
        rayDir = lerp3(diffuseRayDir, specularRayDir, doSpecular);
        ret = add3(ret, mul3(m->emissive, throughput));
        throughput = mul3(throughput, lerp3(m->albedo, m->specularColor, doSpecular));
    }
    return ret;
}

/* simt_textured.cpp:387-431 */
static v3 GetColorForRay_simt_textured(const ctx_t* c, v3 rayPos, v3 rayDir, uint32_t* rng, path_stats_t* st)
{
    v3 ret = V3(0.f, 0.f, 0.f), throughput = V3(1.f, 1.f, 1.f);
    for (int bounceIndex = 0; bounceIndex <= c->p->num_bounces; ++bounceIndex) {
        hit_t h;
        h.dist = c_superFar; h.normal = V3(0.f, 0.f, 0.f); h.fromInside = 0; h.matIndex = -1;
        st->segments++;
        TestSceneTrace_cornell(&c->cornell, rayPos, rayDir, &h);
        if (h.dist == c_superFar) {
            st->escapes++;
            return add3(ret, EquirectSamplePoint(&c->tex, rayDir)); /* no throughput factor, :408-411 */
        }
        const legacy_mat_t* m = &c->cornell.mat[h.matIndex];
        rayPos = add3(add3(rayPos, muls(rayDir, h.dist)), muls(h.normal, c_rayPosNormalNudge));
        rayDir = normalize3(add3(h.normal, RandomUnitVector(rng)));
        ret = add3(ret, mul3(m->emissive, throughput));
        throughput = mul3(throughput, m->albedo);
    }
    return ret;
}

static v3 fma3(v3 a, v3 b, v3 c) { return V3(FMA(a.x, b.x, c.x), FMA(a.y, b.y, c.y), FMA(a.z, b.z, c.z)); }
static v3 fma3s(float a, v3 b, v3 c) { return V3(FMA(a, b.x, c.x), FMA(a, b.y, c.y), FMA(a, b.z, c.z)); }

/* v4.cpp:721-910 */
static v3 GetColorForRay_v4(const ctx_t* c, v3 rayPos, v3 rayDir, uint32_t* rng, path_stats_t* st)
{
    const scene4_t* s = &c->scene4;
    const oracle_params* p = c->p;
    v3 ret = V3(0.f, 0.f, 0.f), throughput = V3(1.f, 1.f, 1.f);
    for (int bounceIndex = 0; bounceIndex <= p->num_bounces; ++bounceIndex) {
        hit_t h;
        h.fromInside = 0; h.dist = c_superFar; h.normal = V3(0.f, 0.f, 0.f); h.matIndex = 0;
        st->segments++;
        for (int i = 0; i < s->numQuads; i++)
            if (TestQuadTrace_v4(rayPos, rayDir, &h, &s->quad[i])) h.matIndex = i;
        for (int i = 0; i < s->numSpheres; i++)
            if (TestSphereTrace_v4(rayPos, rayDir, &h, s->sphereCenter[i], s->sphereRadius[i])) h.matIndex = s->numQuads + i;
        int miss = (h.dist == c_superFar);
        /* the env lookup runs for every lane on every segment (and draws 2 numbers in
         * random-jitter mode); only a missing lane uses the value, :753-778 */
        v3 ambient = V3(.11f, .1f, .15f);
        if (p->env_kind == ORACLE_ENV_CUBEMAP) {
            if (p->env_sampler == ORACLE_SAMPLER_RANDOM) {
                if (miss) ambient = CubemapSampleRandom(&c->tex, rayDir, rng);
                else { RAND01(rng); RAND01(rng); }
            } else if (miss) ambient = CubemapSampleBilinear(&c->tex, rayDir);
        } else if (p->env_kind == ORACLE_ENV_EQUIRECT) {
            v3 SampleDir = V3(-rayDir.x, rayDir.y, -rayDir.z);
            if (p->env_sampler == ORACLE_SAMPLER_RANDOM) {
                if (miss) ambient = EquirectSampleRandom(&c->tex, SampleDir, rng);
                else { RAND01(rng); RAND01(rng); }
            } else if (miss) ambient = EquirectSampleBilinear(&c->tex, SampleDir);
        }
        if (miss) {
            st->escapes++;
            return fma3(ambient, throughput, ret);
        }
        const mat4_t* m = &s->mat[h.matIndex];
        if (h.fromInside) {
            v3 a = muls(neg3(m->refractionColor), h.dist);
            if (c->p->v4_flags & 1) /* USE_FAST_APPROXIMATE_EXP 0: exp_ps, v4.cpp:786 */
                throughput = mul3(throughput, V3(pm_expf(a.x), pm_expf(a.y), pm_expf(a.z)));
            else
                throughput = mul3(throughput, V3(approx_exp1(a.x), approx_exp1(a.y), approx_exp1(a.z)));
        }
        float specularChance = m->specularChance;
        float refractionChance = m->refractionChance;
        {
            int hasSpecularChance = specularChance > 0.f;
            float n1 = h.fromInside ? m->IOR : 1.f;
            float n2 = h.fromInside ? 1.f : m->IOR;
            float newSpecularChance = FresnelReflectAmount(n1, n2, h.normal, rayDir, m->specularChance, 1.f);
            float rcpC = rcp_exact(1.f - m->specularChance);
            float chanceMultiplier = FMA(-newSpecularChance, rcpC, rcpC);
            if (hasSpecularChance) {
                specularChance = newSpecularChance;
                refractionChance = refractionChance * chanceMultiplier;
            }
        }
        float raySelectRoll = RAND01(rng);
        int doSpecular = (specularChance > 0.f) && (raySelectRoll < specularChance);
        int doRefraction = (!doSpecular) && (refractionChance > 0.f) && (raySelectRoll < (specularChance + refractionChance));
        int doDiffuse = (!doSpecular) && (!doRefraction);
        float diffuseChance = max_ps(1.f - (specularChance + refractionChance), 0.f);
        float rayProbability = 1.f;
        if (doSpecular) rayProbability = specularChance;
        if (doRefraction) rayProbability = refractionChance;
        if (doDiffuse) rayProbability = diffuseChance;
        rayProbability = max_ps(rayProbability, 0.001f);

        float doRefractionSign = doRefraction ? -1.f : 1.f;
        v3 newRayPos = fma3s(c_rayPosNormalNudge * doRefractionSign, h.normal, fma3s(h.dist, rayDir, rayPos));

        const int sincos_uv = (c->p->v4_flags & 2) != 0; /* USE_UNIT_VECTOR_REJECTION_SAMPLING 0, v4.cpp:838-861 */
        v3 diffuseRayDir = sincos_uv ? normalize3(add3(h.normal, RandomUnitVector(rng)))
                                     : fast_approx_normalize3(add3(h.normal, RandomUnitVectorRejectionSample(rng)));
        v3 specularRayDir = fma3s(-(2.f * dot3(rayDir, h.normal)), h.normal, rayDir);
        float specularRoughnessSqrd = m->specularRoughness * m->specularRoughness;
        specularRayDir = fma3s(specularRoughnessSqrd, sub3(diffuseRayDir, specularRayDir), specularRayDir);
        float IOR = h.fromInside ? m->IOR : rcp_exact(m->IOR);
        float refractionRoughnessSquared = m->refractionRoughness * m->refractionRoughness;
        v3 refractionRayDir = rfrct(rayDir, h.normal, IOR);
        if (sincos_uv) {
            refractionRayDir = normalize3(lerp3(refractionRayDir, normalize3(sub3(RandomUnitVector(rng), h.normal)), refractionRoughnessSquared));
        } else {
            v3 newRefractionDir = fast_approx_normalize3(sub3(RandomUnitVectorRejectionSample(rng), h.normal));
            refractionRayDir = fma3s(refractionRoughnessSquared, sub3(newRefractionDir, refractionRayDir), refractionRayDir);
        }
        v3 newRayDir = doSpecular ? specularRayDir : diffuseRayDir;
        if (doRefraction) newRayDir = refractionRayDir;
        newRayDir = normalize3(newRayDir);

        ret = fma3(m->emissive, throughput, ret);
        v3 colorFactor = doSpecular ? m->specularColor : m->albedo;
        if (!doRefraction) throughput = mul3(throughput, colorFactor);
        throughput = muls(throughput, rcp_exact(rayProbability));
        {
            float pmax = max_ps(throughput.x, max_ps(throughput.y, throughput.z));
            int rouletteTermination = RAND01(rng) > pmax;
            if (!rouletteTermination) throughput = muls(throughput, rcp_exact(pmax));
        }
        rayPos = newRayPos;
        rayDir = newRayDir;
    }
    return ret;
}

/* v3_redo.cpp:195-219: exact divisions, no fused operations */
static float FresnelReflectAmount_v3(float n1, float n2, v3 normal, v3 incident, float f0, float f90)
{
    float r0 = (n1 - n2) / (n1 + n2);
    r0 = r0 * r0;
    float cosX = -dot3(normal, incident);
    int cond = n1 > n2;
    float n = n1 / n2;
    float sinT2 = n * n * (1.f - cosX * cosX);
    float newCosX = sqrtf(1.f - sinT2);
    int tir = sinT2 > 1.f;
    if (cond && !tir) cosX = newCosX;
    float x = 1.f - cosX;
    float x2 = x * x;
    float ret = r0 + (1.f - r0) * x2 * x2 * x;
    if (cond && tir) ret = 1.f;
    return f0 + ret * (f90 - f0);
}

#define IS_V3REDO(profile) ((profile) == ORACLE_PROFILE_V3REDO || (profile) == ORACLE_PROFILE_V3REDO_SCENE0)

static v3 GetColorForRay_v3redo(const scene3_t* s, const oracle_params* p, const tex_t* tex, v3 rayPos, v3 rayDir,
                                uint32_t* rng, path_stats_t* st)
{
    v3 ret = V3(0.f, 0.f, 0.f), throughput = V3(1.f, 1.f, 1.f);
    for (int bounceIndex = 0; bounceIndex <= p->num_bounces; ++bounceIndex) {
        hit_t h;
        h.fromInside = 0; h.dist = c_superFar; h.normal = V3(0.f, 0.f, 0.f); h.matIndex = -1;
        v3 backdropAlbedo = V3(0.f, 0.f, 0.f);
        st->segments++;
        for (int i = 0; i < s->nquads; i++)
            if (TestQuadTrace_legacy(rayPos, rayDir, &h, s->quad[i][0], s->quad[i][1], s->quad[i][2], s->quad[i][3])) {
                h.matIndex = i;
                if (i == s->backdrop) { /* striped backdrop, :511-515 */
                    v3 hitPos = add3(rayPos, muls(rayDir, h.dist));
                    float shade = floorf(fract1(hitPos.x) * 2.0f);
                    backdropAlbedo = V3(shade, shade, shade);
                }
            }
        for (int i = 0; i < s->nspheres; i++)
            if (TestSphereTrace_legacy(rayPos, rayDir, &h, s->sphereCenter[i], s->sphereRadius[i])) h.matIndex = s->nquads + i;
        if (h.dist == c_superFar) {
            v3 SampleDir = V3(-rayDir.x, rayDir.y, -rayDir.z);
            v3 ambient = mul3(EquirectSampleBilinear(tex, SampleDir), throughput);
            st->escapes++;
            return add3(ret, ambient);
        }
        mat4_t m = s->mat[h.matIndex];
        if (h.matIndex == s->backdrop) m.albedo = backdropAlbedo;
        if (h.fromInside) {
            throughput.x = throughput.x * pm_expf(-m.refractionColor.x * h.dist);
            throughput.y = throughput.y * pm_expf(-m.refractionColor.y * h.dist);
            throughput.z = throughput.z * pm_expf(-m.refractionColor.z * h.dist);
        }
        float specularChance = m.specularChance, refractionChance = m.refractionChance;
        {
            int hasSpecularChance = specularChance > 0.f;
            float n1 = h.fromInside ? m.IOR : 1.f, n2 = h.fromInside ? 1.f : m.IOR;
            float newSpecularChance = FresnelReflectAmount_v3(n1, n2, h.normal, rayDir, m.specularChance, 1.f);
            float chanceMultiplier = (1.f - newSpecularChance) / (1.f - m.specularChance);
            if (hasSpecularChance) {
                specularChance = newSpecularChance;
                refractionChance = refractionChance * chanceMultiplier;
            }
        }
        float raySelectRoll = RAND01(rng);
        int doSpecular = (specularChance > 0.f) && (raySelectRoll < specularChance);
        int doRefraction = (!doSpecular) && (refractionChance > 0.f) && (raySelectRoll < (specularChance + refractionChance));
        int doDiffuse = (!doSpecular) && (!doRefraction);
        float diffuseChance = max_ps(1.f - (specularChance + refractionChance), 0.f);
        float rayProbability = 1.f;
        if (doSpecular) rayProbability = specularChance;
        if (doRefraction) rayProbability = refractionChance;
        if (doDiffuse) rayProbability = diffuseChance;
        rayProbability = max_ps(rayProbability, 1.f * 0.001f);
        float doRefractionSign = doRefraction ? -1.f : 1.f;
        v3 newRayPos = add3(rayPos, add3(muls(rayDir, h.dist), muls(muls(h.normal, doRefractionSign), c_rayPosNormalNudge)));
        v3 diffuseRayDir = normalize3(add3(h.normal, RandomUnitVector(rng)));
        v3 specularRayDir = sub3(rayDir, muls(muls(h.normal, 2.f), dot3(rayDir, h.normal)));
        float specularRoughnessSqrd = m.specularRoughness * m.specularRoughness;
        specularRayDir = normalize3(lerp3(specularRayDir, diffuseRayDir, specularRoughnessSqrd));
        float IOR = h.fromInside ? m.IOR : 1.0f / m.IOR;
        float refractionRoughnessSquared = m.refractionRoughness * m.refractionRoughness;
        v3 refractionRayDir = rfrct(rayDir, h.normal, IOR);
        refractionRayDir = normalize3(lerp3(refractionRayDir, normalize3(sub3(RandomUnitVector(rng), h.normal)), refractionRoughnessSquared));
        v3 newRayDir = doSpecular ? specularRayDir : diffuseRayDir;
        if (doRefraction) newRayDir = refractionRayDir;
        ret = add3(ret, mul3(m.emissive, throughput));
        v3 colorFactor = doSpecular ? m.specularColor : m.albedo;
        if (!doRefraction) throughput = mul3(throughput, colorFactor);
        throughput = V3(throughput.x / rayProbability, throughput.y / rayProbability, throughput.z / rayProbability);
        {
            float pmax = max_ps(throughput.x, max_ps(throughput.y, throughput.z));
            int rouletteTermination = RAND01(rng) > pmax;
            if (!rouletteTermination) throughput = muls(throughput, 1.0f / pmax);
        }
        rayPos = newRayPos;
        rayDir = newRayDir;
    }
    return ret;
}

/* mainImage: v2.cpp:526-568, simt_textured.cpp:433-474, v4.cpp:1092-1131 */
static v3 mainImage(const ctx_t* c, int x, int yflip, int frame, uint32_t* rng_out, path_stats_t* st)
{
    const oracle_params* p = c->p;
    uint32_t rng = oracle_seed(x, yflip, frame);
    float fx = (float)x, fy = (float)yflip;
    float resx = (float)p->width, resy = (float)p->height;
    v3 color;
    if (p->profile == ORACLE_PROFILE_V4) {
        float rcpx = rcp_exact(resx), rcpy = rcp_exact(resy);
        float jx = RAND01(&rng) - .5f;
        float jy = RAND01(&rng) - .5f;
        float tx = FMA((fx + jx) * rcpx, 2.f, -1.f);
        float ty = FMA((fy + jy) * rcpy, 2.f, -1.f);
        v3 rayTarget = V3(tx, ty, -c->scene4.cameraDistance);
        rayTarget.y = rayTarget.y * (rcpx * resy);
        v3 rayDir = normalize3(sub3(rayTarget, V3(0.f, 0.f, 0.f)));
        v3 col = GetColorForRay_v4(c, c->scene4.cameraPosition, rayDir, &rng, st);
        color = fma3s(1.f / 1, col, V3(0.f, 0.f, 0.f));
    } else {
        float cameraDistance = oracle_camera_distance();
        float tx, ty;
        if (p->profile == ORACLE_PROFILE_V2 || IS_V3REDO(p->profile)) {
            float jx = RAND01(&rng) - .5f;
            float jy = RAND01(&rng) - .5f;
            tx = ((fx + jx) / resx) * 2.0f - 1.f;
            ty = ((fy + jy) / resy) * 2.0f - 1.f;
        } else {
            tx = (fx / resx) * 2.0f - 1.f;
            ty = (fy / resy) * 2.0f - 1.f;
        }
        v3 rayTarget = V3(tx, ty, cameraDistance);
        float aspectRatio = resx / resy;
        rayTarget.y = rayTarget.y / aspectRatio;
        v3 rayPosition = V3(0.f, 0.f, 0.f);
        v3 rayDir = normalize3(sub3(rayTarget, rayPosition));
        v3 col;
        if (IS_V3REDO(p->profile)) { /* v3_redo.cpp:791-798: camera at (0,0,40) looking down -z */
            rayDir.z = rayDir.z * -1.f;
            col = GetColorForRay_v3redo(&c->scene3, p, &c->tex, V3(0.f, 0.f, 1.f * 40.f), rayDir, &rng, st);
        } else
            col = (p->profile == ORACLE_PROFILE_V2) ? GetColorForRay_v2(c, rayPosition, rayDir, &rng, st)
                                                    : GetColorForRay_simt_textured(c, rayPosition, rayDir, &rng, st);
        color = add3(V3(0.f, 0.f, 0.f), muls(col, 1.f / 1));
    }
    if (rng_out) *rng_out = rng;
    return color;
}

static int ctx_init(ctx_t* c, const oracle_params* p)
{
    if (!p || p->width <= 0 || p->height <= 0 || p->num_tiles_x <= 0 || p->num_tiles_y <= 0) return -1;
    if (p->width % p->num_tiles_x || p->height % p->num_tiles_y || (p->width / p->num_tiles_x) % 8) return -1;
    if (p->profile < 0 || p->profile > 4 || p->num_bounces < 0) return -1;
    int needs_env = (p->profile == ORACLE_PROFILE_SIMT_TEXTURED) || IS_V3REDO(p->profile) ||
                    (p->profile == ORACLE_PROFILE_V4 && p->env_kind != ORACLE_ENV_NONE);
    if (needs_env && (!p->env || p->env_width <= 0 || p->env_height <= 0)) return -1;
    c->p = p;
    cornell_init(&c->cornell, p->profile);
    if (p->scene_cornell && (p->profile == ORACLE_PROFILE_V2 || p->profile == ORACLE_PROFILE_SIMT_TEXTURED)) {
        if (!p->scene_cornell->quad_vertices || !p->scene_cornell->spheres || !p->scene_cornell->materials) return -1;
        cornell_from(&c->cornell, p->scene_cornell);
    }
    scene4_init(&c->scene4);
    if (p->profile == ORACLE_PROFILE_V4 && p->scene_v4 && scene4_from(&c->scene4, p->scene_v4)) return -1;
    scene3_init(&c->scene3, p->profile == ORACLE_PROFILE_V3REDO_SCENE0 ? 0 : 1);
    c->tex.data = p->env; c->tex.W = p->env_width; c->tex.H = p->env_height;
    return 0;
}

int64_t oracle_buffer_index(int W, int H, int ntx, int nty, int x, int y, int ch)
{
    int TW = W / ntx, TH = H / nty;
    int tx = x / TW, ty = y / TH, lx = x % TW, ly = y % TH;
    return (int64_t)ty * TH * W * 3 + (int64_t)tx * TW * TH * 3 + ((int64_t)ly * TW + (lx & ~7)) * 3 + ch * 8 + (lx & 7);
}

typedef struct {
    const ctx_t* c;
    float* target;
    int first_frame, nframes;
    int next_row; /* atomic row counter: rows are independent (per-pixel RNG, no shared state) */
    uint64_t segments, escapes;
    pthread_mutex_t lock;
} render_job_t;

static void* render_worker(void* arg)
{
    render_job_t* job = (render_job_t*)arg;
    const ctx_t* c = job->c;
    const oracle_params* p = c->p;
    const int W = p->width, H = p->height;
    const int legacy_blend = (p->profile != ORACLE_PROFILE_V4);
    path_stats_t st = {0, 0};
    for (;;) {
        int y = __sync_fetch_and_add(&job->next_row, 1);
        if (y >= H) break;
        for (int x = 0; x < W; x++) {
            int64_t i0 = oracle_buffer_index(W, H, p->num_tiles_x, p->num_tiles_y, x, y, 0);
            float* px = job->target + i0;
            v3 avg = V3(px[0], px[8], px[16]);
            for (int f = 0; f < job->nframes; f++) {
                int frame = job->first_frame + f;
                v3 color = mainImage(c, x, H - 1 - y, frame, 0, &st);
                float blend = 1.0f / (float)((float)frame + 1.f);
                if (legacy_blend) avg = lerp3(avg, color, blend);   /* v2.cpp:623 */
                else avg = fma3s(blend, sub3(color, avg), avg);      /* v4.cpp:1239 */
            }
            px[0] = avg.x; px[8] = avg.y; px[16] = avg.z;
        }
    }
    pthread_mutex_lock(&job->lock);
    job->segments += st.segments;
    job->escapes += st.escapes;
    pthread_mutex_unlock(&job->lock);
    return 0;
}

int oracle_render(const oracle_params* p, float* target, int first_frame, int nframes, int nthreads, oracle_counters* counters)
{
    ctx_t c;
    if (ctx_init(&c, p) || !target || nframes < 0) return -1;
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    render_job_t job;
    memset(&job, 0, sizeof(job));
    job.c = &c; job.target = target; job.first_frame = first_frame; job.nframes = nframes;
    pthread_mutex_init(&job.lock, 0);
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nthreads - 1; i++)
        if (pthread_create(&th[started], 0, render_worker, &job) == 0) started++;
    render_worker(&job);
    for (int i = 0; i < started; i++) pthread_join(th[i], 0);
    pthread_mutex_destroy(&job.lock);
    if (counters) {
        counters->paths = (uint64_t)p->width * p->height * (uint64_t)nframes;
        counters->segments = job.segments;
        counters->escapes = job.escapes;
    }
    return 0;
}

int oracle_path_radiance(const oracle_params* p, int x, int y, int frame, float rgb[3])
{
    ctx_t c;
    if (ctx_init(&c, p)) return -1;
    path_stats_t st = {0, 0};
    v3 col = mainImage(&c, x, p->height - 1 - y, frame, 0, &st);
    rgb[0] = col.x; rgb[1] = col.y; rgb[2] = col.z;
    return 0;
}

uint32_t oracle_final_rng_state(const oracle_params* p, int x, int y, int frame)
{
    ctx_t c;
    if (ctx_init(&c, p)) return 0;
    path_stats_t st = {0, 0};
    uint32_t rng = 0;
    mainImage(&c, x, p->height - 1 - y, frame, &rng, &st);
    return rng;
}

/* ---- LDR resolve: v4.cpp:144-187 (fast gamma, fast ACES), :1260-1331 ------------------------ */
static float fast_pow_gamma(float x)
{
    float sqrtx = sqrtf(x);
    float onethird = 1.f / 3.f, twothirds = 2.f / 3.f;
    float nit1 = FMA(sqrtx, twothirds, onethird);
    float nit2 = FMA(nit1, twothirds, (x * rcp_exact(nit1 * nit1)) * onethird);
    float nit3 = FMA(nit2, twothirds, (x * rcp_exact(nit2 * nit2)) * onethird);
    return sqrtf(sqrtx * nit3);
}
static float aces1(float X)
{
    const float a = 2.51f, b = 0.03f, cc = 2.43f, d = 0.59f, e = 0.14f;
    float rcpDenom = rcp_exact(FMA(X, FMA(cc, X, d), e));
    return saturate1((X * FMA(a, X, b)) * rcpDenom);
}
static float aces1_exact(float X) /* USE_FAST_APPROXIMATE_ACES_TONEMAP 0, v4.cpp:172-175 */
{
    const float a = 2.51f, b = 0.03f, cc = 2.43f, d = 0.59f, e = 0.14f;
    return saturate1((X * (a * X + b)) / (X * (cc * X + d) + e));
}
static float srgb1(float v)
{
    v = saturate1(v);
    return (v < 0.0031308f) ? v * 12.92f : FMA(1.055f, fast_pow_gamma(v), -0.055f);
}
static float srgb1_exact(float v) /* USE_FAST_APPROXIMATE_GAMMA 0, v4.cpp:185: 1.055f * pow_ps(rgb, 1.f / 2.4f) - 0.055f */
{
    v = saturate1(v);
    return (v < 0.0031308f) ? v * 12.92f : 1.055f * pm_powf(v, 1.0f / 2.4f) - 0.055f;
}

int oracle_resolve_ldr(const float* target, int W, int H, int ntx, int nty, uint32_t* out, int mode)
{
    if (!target || !out || W % ntx || H % nty || (W / ntx) % 8) return -1;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const float* px = target + oracle_buffer_index(W, H, ntx, nty, x, y, 0);
            float c[3] = {px[0], px[8], px[16]};
            uint32_t q[3];
            for (int k = 0; k < 3; k++) {
                float v = (mode & 2) ? aces1_exact(c[k] * 1.0f) : aces1(c[k] * 1.0f);
                v = (mode & 4) ? srgb1_exact(v) : srgb1(v);
                v = saturate1(v) * 255.f;
                q[k] = (uint32_t)to_epi32(v) & 0xFFu;
            }
            out[(size_t)y * W + x] = ((mode & 1) == 0) ? (0xFF000000u | (q[2] << 16) | (q[1] << 8) | q[0])
                                                 : ((q[0] << 16) | (q[1] << 8) | q[2]);
        }
    return 0;
}

/* ---- test hooks for oracle/portable_math.h -------------------------------------------------- */
void oracle_pm_sincosf(float a, float* s, float* c) { pm_sincosf(a, s, c); }
float oracle_pm_atan2f(float y, float x) { return pm_atan2f(y, x); }
float oracle_pm_asinf(float x) { return pm_asinf(x); }
void oracle_pm_sincosf_array(const float* a, float* s, float* c, int64_t n)
{
    for (int64_t i = 0; i < n; i++) pm_sincosf(a[i], &s[i], &c[i]);
}
void oracle_pm_atan2f_array(const float* y, const float* x, float* out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) out[i] = pm_atan2f(y[i], x[i]);
}
void oracle_pm_asinf_array(const float* x, float* out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) out[i] = pm_asinf(x[i]);
}

/* max number of traced segments over frames [first_frame, first_frame+nframes) for every pixel
 * (row-major, row 0 = top): 1 means every path of the pixel escaped on its camera ray */
int oracle_max_segments(const oracle_params* p, int first_frame, int nframes, uint32_t* out)
{
    ctx_t c;
    if (ctx_init(&c, p) || !out) return -1;
    for (int y = 0; y < p->height; y++)
        for (int x = 0; x < p->width; x++) {
            uint32_t mx = 0;
            for (int f = 0; f < nframes; f++) {
                path_stats_t st = {0, 0};
                mainImage(&c, x, p->height - 1 - y, first_frame + f, 0, &st);
                if (st.segments > mx) mx = (uint32_t)st.segments;
            }
            out[(size_t)y * p->width + x] = mx;
        }
    return 0;
}
void oracle_pm_powf_array(const float* x, const float* y, float* out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) out[i] = pm_powf(x[i], y[i]);
}
void oracle_pm_expf_array(const float* x, float* out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) out[i] = pm_expf(x[i]);
}
