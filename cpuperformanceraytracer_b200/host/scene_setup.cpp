// scene_setup.cpp -- host-side construction of the scene tables the kernel reads.
//
// The reference builds its scenes at start-up with AVX2 float arithmetic
// (InitializeScene / PrecomputeQuadData / InitializeCamera,
// demofox_path_tracing_optimization_v4.cpp:269-319,1403-1502) or re-evaluates constant vertex
// expressions per ray (demofox_path_tracing_v2.cpp:320-454).  The same binary32 expressions are
// evaluated here once, on the host, in the same order, so the tables hold the same bits.
// Compile with -ffp-contract=off; fused operations are spelled std::fmaf where the reference's
// dot()/cross() fuse (mathlib.h:144-146, 770-778).
#include "scene_setup.h"

#include <cmath>
#include <cstring>

namespace b200pt {
namespace {

inline v3 mk(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
inline v3 add(v3 u, v3 v) { return mk(u.x + v.x, u.y + v.y, u.z + v.z); }
inline v3 sub(v3 u, v3 v) { return mk(u.x - v.x, u.y - v.y, u.z - v.z); }
inline v3 muls(v3 u, float c) { return mk(u.x * c, u.y * c, u.z * c); }
inline v3 divs(v3 u, float c) { return mk(u.x / c, u.y / c, u.z / c); }
inline v3 neg(v3 u) { return mk(-u.x, -u.y, -u.z); }
inline float dot(v3 u, v3 v) { return std::fmaf(u.x, v.x, std::fmaf(u.y, v.y, u.z * v.z)); }
inline v3 cross(v3 u, v3 v)
{
    return mk(std::fmaf(u.y, v.z, -(u.z * v.y)), std::fmaf(u.z, v.x, -(u.x * v.z)), std::fmaf(u.x, v.y, -(u.y * v.x)));
}
inline v3 normalize(v3 v) { return muls(v, 1.0f / std::sqrt(dot(v, v))); }

}  // namespace

float camera_distance()
{
    const float c_FOVDegrees = 90.0f;
    const float c_pi = 3.14159265359f;
    return 1.0f / std::tan(c_FOVDegrees * 0.5f * c_pi / 180.0f);  // v2.cpp:546, v4.cpp:1500
}

void build_cornell_scene(CornellScene* s, bool simt_textured_materials)
{
    std::memset(s, 0, sizeof(*s));
    const v3 T = mk(kCornellTranslation[0], kCornellTranslation[1], kCornellTranslation[2]);  // sceneTranslation, v2.cpp:323
    const auto& Q = kCornellQuadVerts;  // v2.cpp:328-331,345-348,362-365,379-382,396-399,413-416
    for (int i = 0; i < kCornellQuads; i++) {
        LegacyQuad& q = s->quad[i];
        q.a = add(mk(Q[i][0][0], Q[i][0][1], Q[i][0][2]), T);
        q.b = add(mk(Q[i][1][0], Q[i][1][1], Q[i][1][2]), T);
        q.c = add(mk(Q[i][2][0], Q[i][2][1], Q[i][2][2]), T);
        q.d = add(mk(Q[i][3][0], Q[i][3][1], Q[i][3][2]), T);
        q.n = normalize(cross(sub(q.c, q.a), sub(q.c, q.b)));  // v2.cpp:166
    }
    static const float SX[kCornellSpheres] = {-9.0f, 0.0f, 9.0f};  // v2.cpp:429,438,447
    for (int i = 0; i < kCornellSpheres; i++) {
        const v3 c = add(mk(SX[i], -9.5f, 20.0f), T);
        s->sphere[i] = make_float4(c.x, c.y, c.z, 3.0f + 0.0f);
    }
    LegacyMaterial* m = s->mat;
    m[0].albedo = mk(0.7f, 0.7f, 0.7f);
    m[1].albedo = mk(0.7f, 0.7f, 0.7f);
    m[2].albedo = mk(0.7f, 0.7f, 0.7f);
    m[3].albedo = mk(0.7f, 0.1f, 0.1f);
    m[4].albedo = mk(0.1f, 0.7f, 0.1f);
    m[5].emissive = muls(mk(1.0f, 0.9f, 0.7f), 20.0f);
    if (!simt_textured_materials) {  // v2.cpp:430-452
        m[6].albedo = mk(0.9f, 0.9f, 0.5f); m[6].percentSpecular = 0.1f; m[6].roughness = 0.2f; m[6].specularColor = mk(0.9f, 0.9f, 0.9f);
        m[7].albedo = mk(0.9f, 0.5f, 0.9f); m[7].percentSpecular = 0.3f; m[7].roughness = 0.2f; m[7].specularColor = mk(0.9f, 0.9f, 0.9f);
        m[8].albedo = mk(0.f, 0.f, 1.f);    m[8].percentSpecular = 0.5f; m[8].roughness = 0.4f; m[8].specularColor = mk(1.f, 0.f, 0.f);
    } else {  // simt_textured.cpp:370-383
        m[6].albedo = mk(0.9f, 0.9f, 0.75f);
        m[7].albedo = mk(0.9f, 0.75f, 0.9f);
        m[8].albedo = mk(0.9f, 0.75f, 0.9f);
    }
}

bool build_cornell_scene_from(CornellScene* s, const float* qv, const float* sp, const float* mats)
{
    std::memset(s, 0, sizeof(*s));
    for (int i = 0; i < kCornellQuads * 12; i++)
        if (!(std::fabs(qv[i]) <= 1e6f)) return false;
    for (int i = 0; i < kCornellSpheres; i++) {
        for (int k = 0; k < 3; k++)
            if (!(std::fabs(sp[4 * i + k]) <= 1e6f)) return false;
        if (!(sp[4 * i + 3] >= 1e-3f && sp[4 * i + 3] <= 1e6f)) return false;
    }
    for (int i = 0; i < kCornellQuads; i++) {
        LegacyQuad& q = s->quad[i];
        const float* v = qv + 12 * i;
        q.a = mk(v[0], v[1], v[2]);
        q.b = mk(v[3], v[4], v[5]);
        q.c = mk(v[6], v[7], v[8]);
        q.d = mk(v[9], v[10], v[11]);
        q.n = normalize(cross(sub(q.c, q.a), sub(q.c, q.b)));  // v2.cpp:166
    }
    for (int i = 0; i < kCornellSpheres; i++) s->sphere[i] = make_float4(sp[4 * i], sp[4 * i + 1], sp[4 * i + 2], sp[4 * i + 3]);
    for (int i = 0; i < kCornellObjects; i++) {
        const float* m = mats + 11 * i;
        LegacyMaterial& d = s->mat[i];
        d.albedo = mk(m[0], m[1], m[2]);
        d.emissive = mk(m[3], m[4], m[5]);
        d.specularColor = mk(m[6], m[7], m[8]);
        d.percentSpecular = m[9];
        d.roughness = m[10];
    }
    return true;
}

// PrecomputeQuadData, v4.cpp:269-319
static void precompute_quad(V4Quad* q, v3 V0, v3 V1, v3 V2, v3 V3)
{
    const v3 V01 = sub(V1, V0), V02 = sub(V2, V0), V30 = sub(V0, V3);
    const v3 V20 = neg(V02);
    const v3 V01xV02 = cross(V01, V02);
    const v3 V02xV03 = cross(V30, V01);
    const v3 N = normalize(V01xV02);
    const float DetTop = dot(V02xV03, N);
    const float DetBot = dot(V01xV02, N);
    q->V0 = V0;
    q->n = N;
    q->NxV01 = divs(cross(N, V01), DetBot);
    q->NxV20 = divs(cross(N, V20), DetBot);
    q->NxV02 = divs(cross(N, V02), DetTop);
    q->NxV30 = divs(cross(N, V30), DetTop);
}

void build_v4_scene(V4Scene* s)
{
    std::memset(s, 0, sizeof(*s));
    s->numQuads = kV4Quads;
    s->numSpheres = kV4Spheres;
    const v3 T = mk(0.0f, 0.0f, 10.0f);  // v4.cpp:1407
    precompute_quad(&s->quad[0], add(mk(-25.0f, -12.5f, 5.0f), T), add(mk(25.0f, -12.5f, 5.0f), T),
                    add(mk(25.0f, -12.5f, -5.0f), T), add(mk(-25.0f, -12.5f, -5.0f), T));           // floor :1416-1419
    precompute_quad(&s->quad[1], mk(-25.0f, -1.5f, 5.0f), mk(25.0f, -1.5f, 5.0f), mk(25.0f, -10.5f, 5.0f),
                    mk(-25.0f, -10.5f, 5.0f));                                                       // backdrop, untranslated :1430-1433
    precompute_quad(&s->quad[2], add(mk(-7.5f, 12.5f, 5.0f), T), add(mk(7.5f, 12.5f, 5.0f), T),
                    add(mk(7.5f, 12.5f, -5.0f), T), add(mk(-7.5f, 12.5f, -5.0f), T));               // ceiling :1447-1450
    precompute_quad(&s->quad[3], add(mk(-5.0f, 12.4f, 2.5f), T), add(mk(5.0f, 12.4f, 2.5f), T),
                    add(mk(5.0f, 12.4f, -2.5f), T), add(mk(-5.0f, 12.4f, -2.5f), T));               // light :1461-1464
    // AddMaterialToScene copies albedo.x into all three channels (v4.cpp:1370-1372): preserved
    s->mat[0].albedo = mk(0.7f, 0.7f, 0.7f);
    s->mat[1].albedo = mk(.35f, .35f, .35f);
    s->mat[2].albedo = mk(0.7f, 0.7f, 0.7f);
    s->mat[3].emissive = muls(mk(1.0f, 0.9f, 0.7f), 20.0f);
    const int c_numSpheres = kV4Spheres;
    for (int i = 0; i < c_numSpheres; i++) {  // v4.cpp:1474-1495
        const v3 c = add(mk(-18.0f + 6.0f * (float)i, -8.0f, 0.0f), T);  // == (v4_sphere_x(i), kV4SphereY, kV4SphereZ)
        s->sphere[i] = make_float4(c.x, c.y, c.z, 2.8f + 0.0f);
        V4Material& m = s->mat[kV4Quads + i];
        const float r = (((float)i) / (float)(c_numSpheres - 1)) * 0.5f;
        m.specularChance = 0.02f;
        m.IOR = 1.1f;
        m.refractionChance = 1.0f;
        m.albedo = mk(0.9f, 0.9f, 0.9f);
        m.refractionColor = mk(0.0f, 0.5f, 1.0f);
        m.specularColor = muls(mk(1.0f, 1.0f, 1.0f), 0.8f);
        m.specularRoughness = r;
        m.refractionRoughness = r;
    }
    s->cameraDistance = camera_distance();
    s->cameraPosition = mk(0.f, 0.f, 1.f * 40.f);  // v4.cpp:1501
}

bool build_v4_scene_from(V4Scene* s, const float* qv, int nq, const float* sp, int ns, const float* mats,
                         const float cam_pos[3], float cam_dist)
{
    if (nq < 0 || ns < 0 || nq + ns < 1 || nq + ns > kV4MaxObjects || (nq && !qv) || (ns && !sp) || !mats || !cam_pos) return false;
    std::memset(s, 0, sizeof(*s));
    s->numQuads = nq;
    s->numSpheres = ns;
    for (int i = 0; i < nq; i++) {
        const float* v = qv + 12 * i;
        precompute_quad(&s->quad[i], mk(v[0], v[1], v[2]), mk(v[3], v[4], v[5]), mk(v[6], v[7], v[8]), mk(v[9], v[10], v[11]));
    }
    for (int i = 0; i < ns; i++) s->sphere[i] = make_float4(sp[4 * i], sp[4 * i + 1], sp[4 * i + 2], sp[4 * i + 3]);
    for (int i = 0; i < nq + ns; i++) {
        const float* m = mats + 17 * i;
        V4Material& d = s->mat[i];
        d.albedo = mk(m[0], m[0], m[0]);  // AddMaterialToScene stores albedo.x three times, v4.cpp:1370-1372
        d.emissive = mk(m[3], m[4], m[5]);
        d.specularChance = m[6];
        d.specularRoughness = m[7];
        d.specularColor = mk(m[8], m[9], m[10]);
        d.IOR = m[11];
        d.refractionChance = m[12];
        d.refractionRoughness = m[13];
        d.refractionColor = mk(m[14], m[15], m[16]);
    }
    s->cameraPosition = mk(cam_pos[0], cam_pos[1], cam_pos[2]);
    s->cameraDistance = cam_dist;
    return true;
}

// Scene of demofox_path_tracing_v3_redo.cpp, SCENE 1 (:485-600)
void build_v3redo_scene(V3RedoScene* s)
{
    std::memset(s, 0, sizeof(*s));
    for (int i = 0; i < kV3Quads; i++) {
        LegacyQuad& q = s->quad[i];
        const v3 T = mk(kV3Translation[i][0], kV3Translation[i][1], kV3Translation[i][2]);
        v3* dst[4] = {&q.a, &q.b, &q.c, &q.d};
        for (int k = 0; k < 4; k++) {
            const v3 p = mk(kV3QuadVerts[i][k][0], kV3QuadVerts[i][k][1], kV3QuadVerts[i][k][2]);
            *dst[k] = (i == kV3BackdropQuad) ? p : add(p, T);  // the backdrop literals are used untranslated (:505-508)
        }
        q.n = normalize(cross(sub(q.c, q.a), sub(q.c, q.b)));  // :225
    }
    for (int i = 0; i < kV3Objects; i++) s->mat[i].IOR = 1.f;  // GetZeroedMaterial, :155-168
    s->mat[0].albedo = mk(0.7f, 0.7f, 0.7f);
    s->mat[2].albedo = mk(0.7f, 0.7f, 0.7f);
    s->mat[3].emissive = muls(mk(1.0f, 0.9f, 0.7f), 20.0f);
    for (int i = 0; i < kV3Spheres; i++) {  // :575-598
        const v3 c = add(mk(-18.0f + 6.0f * (float)i, -8.0f, 0.0f), mk(0.0f, 0.0f, 10.0f));
        s->sphere[i] = make_float4(c.x, c.y, c.z, 2.8f + 0.0f);
        V4Material& m = s->mat[kV3Quads + i];
        const float r = ((float)i / (float)(kV3Spheres - 1)) * 0.5f;
        m.albedo = mk(0.9f, 0.25f, 0.25f);
        m.specularChance = 0.02f;
        m.specularRoughness = r;
        m.specularColor = muls(mk(1.0f, 1.0f, 1.0f), 0.8f);
        m.IOR = 1.1f;
        m.refractionChance = 1.0f;
        m.refractionRoughness = r;
        m.refractionColor = mk(0.0f, 0.5f, 1.0f);
    }
    s->cameraPosition = mk(0.f, 0.f, 1.f * 40.f);  // :796
}

// Scene of demofox_path_tracing_v3_redo.cpp compiled with SCENE 0 (:392-479 walls, :530-543 light, :545-580 spheres)
void build_v3redo_scene0(V3RedoScene0* s)
{
    std::memset(s, 0, sizeof(*s));
    const v3 T = mk(0.0f, 0.0f, 0.0f);  // sceneTranslation, :387
    for (int i = 0; i < kV3S0Quads; i++) {
        LegacyQuad& q = s->quad[i];
        v3* dst[4] = {&q.a, &q.b, &q.c, &q.d};
        for (int k = 0; k < 4; k++) *dst[k] = add(mk(kV3S0QuadVerts[i][k][0], kV3S0QuadVerts[i][k][1], kV3S0QuadVerts[i][k][2]), T);
        q.n = normalize(cross(sub(q.c, q.a), sub(q.c, q.b)));  // :225
    }
    for (int i = 0; i < kV3S0Objects; i++) s->mat[i].IOR = 1.f;  // GetZeroedMaterial, :155-168
    s->mat[0].albedo = mk(0.7f, 0.7f, 0.7f);
    s->mat[1].albedo = mk(0.7f, 0.7f, 0.7f);
    s->mat[2].albedo = mk(0.7f, 0.7f, 0.7f);
    s->mat[3].albedo = mk(0.7f, 0.1f, 0.1f);
    s->mat[4].albedo = mk(0.1f, 0.7f, 0.1f);
    s->mat[5].emissive = muls(mk(1.0f, 0.9f, 0.7f), 20.0f);
    const float cx[3] = {-9.0f, 0.0f, 9.0f};
    for (int i = 0; i < kV3S0Spheres; i++) {
        const v3 c = add(mk(cx[i], -9.5f, 20.0f), T);
        s->sphere[i] = make_float4(c.x, c.y, c.z, 3.0f + 0.0f);
    }
    V4Material* m = &s->mat[kV3S0Quads];
    m[0].albedo = mk(0.9f, 0.9f, 0.5f); m[0].specularChance = 0.1f; m[0].specularRoughness = 0.2f; m[0].specularColor = mk(0.9f, 0.9f, 0.9f);
    m[1].albedo = mk(0.9f, 0.5f, 0.9f); m[1].specularChance = 0.3f; m[1].specularRoughness = 0.2f; m[1].specularColor = mk(0.9f, 0.9f, 0.9f);
    m[2].albedo = mk(0.f, 0.f, 1.f);    m[2].specularChance = 0.5f; m[2].specularRoughness = 0.4f; m[2].specularColor = mk(1.f, 0.f, 0.f);
    s->cameraPosition = mk(0.f, 0.f, 1.f * 40.f);  // :796
}

// ---- camera-ray culling ---------------------------------------------------------------------------
// A camera ray through fragCoord (fx, fy) has direction (tx, ty / aspect, camDist) with
// tx = fx / W * 2 - 1, ty = fy / H * 2 - 1 from the origin (v2.cpp:543-560), or
// (tx, ty * H / W, -camDist) from (0, 0, 40) (v4.cpp:1108-1121).  A primitive can only be hit by
// rays whose fragCoord lies inside the projection of its bounding box.  The boxes are projected in
// double precision and grown by 2 pixels + 0.1 %, three orders of magnitude more than the binary32
// error of the reference's edge tests, so "outside every rectangle" implies that every sign test
// of the reference fails by a wide margin (tests/test_cull_rects.py checks it against the oracle).
namespace {
struct Box { double lo[3], hi[3]; };

bool project_box(const Box& b, int profile, int W, int H, double camDist, float4* out, const double* camPos = nullptr)
{
    const double cx = camPos ? camPos[0] : 0.0, cy = camPos ? camPos[1] : 0.0, cz = camPos ? camPos[2] : 40.0;
    double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
    for (int k = 0; k < 8; k++) {
        const double px = (k & 1) ? b.hi[0] : b.lo[0], py = (k & 2) ? b.hi[1] : b.lo[1], pz = (k & 4) ? b.hi[2] : b.lo[2];
        double tx, ty;
        if (profile == kProfileV4) {
            const double depth = (cz - pz) / camDist;  // camera (default (0, 0, 40)) looking down -z
            if (depth < 1e-3) return false;
            tx = (px - cx) / depth;
            ty = (py - cy) / depth * ((double)W / (double)H);
        } else {
            const double depth = pz / camDist;  // camera at the origin looking down +z
            if (depth < 1e-3) return false;
            tx = px / depth;
            ty = py / depth * ((double)W / (double)H);  // ty / aspect = y / depth
        }
        const double fx = (tx + 1.0) * 0.5 * W, fy = (ty + 1.0) * 0.5 * H;
        x0 = fx < x0 ? fx : x0; x1 = fx > x1 ? fx : x1;
        y0 = fy < y0 ? fy : y0; y1 = fy > y1 ? fy : y1;
    }
    const double mx = 2.0 + 1e-3 * (x1 - x0 + W), my = 2.0 + 1e-3 * (y1 - y0 + H);
    *out = make_float4((float)(x0 - mx), (float)(y0 - my), (float)(x1 + mx), (float)(y1 + my));
    return true;
}

void grow(Box* b, const v3& p)
{
    const double c[3] = {p.x, p.y, p.z};
    for (int a = 0; a < 3; a++) {
        if (c[a] < b->lo[a]) b->lo[a] = c[a];
        if (c[a] > b->hi[a]) b->hi[a] = c[a];
    }
}
Box empty_box()
{
    Box b;
    for (int a = 0; a < 3; a++) { b.lo[a] = 1e300; b.hi[a] = -1e300; }
    return b;
}
Box sphere_box(const float4& s)
{
    Box b;
    const double c[3] = {s.x, s.y, s.z};
    for (int a = 0; a < 3; a++) { b.lo[a] = c[a] - 1.001 * s.w; b.hi[a] = c[a] + 1.001 * s.w; }
    return b;
}
}  // namespace

int compute_cull_rects_v4(const float* qv, int nq, const float* sp, int ns, const float cam_pos[3], float cam_dist, int width,
                          int height, float4* rects)
{
    if (nq + ns > kMaxCullRects || !(cam_dist > 0.f)) return -1;
    const double cp[3] = {cam_pos[0], cam_pos[1], cam_pos[2]};
    int n = 0;
    for (int i = 0; i < nq; i++) {
        Box b = empty_box();
        for (int k = 0; k < 4; k++) grow(&b, mk(qv[12 * i + 3 * k], qv[12 * i + 3 * k + 1], qv[12 * i + 3 * k + 2]));
        for (int a = 0; a < 3; a++) {
            const double pad = 1e-3 + 1e-5 * (std::fabs(b.lo[a]) + std::fabs(b.hi[a]));
            b.lo[a] -= pad;
            b.hi[a] += pad;
        }
        if (!project_box(b, kProfileV4, width, height, cam_dist, &rects[n++], cp)) return -1;
    }
    for (int i = 0; i < ns; i++)
        if (!project_box(sphere_box(make_float4(sp[4 * i], sp[4 * i + 1], sp[4 * i + 2], std::fabs(sp[4 * i + 3]))), kProfileV4, width,
                         height, cam_dist, &rects[n++], cp))
            return -1;
    return n;
}

int compute_cull_rects_cornell(const float* qv, const float* sp, int width, int height, float4* rects)
{
    const double camDist = camera_distance();
    int n = 0;
    for (int i = 0; i < kCornellQuads; i++) {
        Box b = empty_box();
        for (int k = 0; k < 4; k++) grow(&b, mk(qv[12 * i + 3 * k], qv[12 * i + 3 * k + 1], qv[12 * i + 3 * k + 2]));
        for (int a = 0; a < 3; a++) {
            const double pad = 1e-3 + 1e-5 * (std::fabs(b.lo[a]) + std::fabs(b.hi[a]));
            b.lo[a] -= pad;
            b.hi[a] += pad;
        }
        if (!project_box(b, kProfileV2, width, height, camDist, &rects[n++])) return -1;
    }
    for (int i = 0; i < kCornellSpheres; i++)
        if (!project_box(sphere_box(make_float4(sp[4 * i], sp[4 * i + 1], sp[4 * i + 2], std::fabs(sp[4 * i + 3]))), kProfileV2, width, height,
                         camDist, &rects[n++]))
            return -1;
    return n;
}

int compute_cull_rects(int profile, int width, int height, float4* rects)
{
    const double camDist = camera_distance();
    int n = 0;
    if (profile == kProfileV4 || profile == kProfileV3Redo) {  // same geometry, same camera model
        V4Scene s;
        build_v4_scene(&s);
        // the quad table only keeps V0 and edge bivectors; rebuild the vertex boxes from the source data
        const double T = 10.0;
        const double q[kV4Quads][2][3] = {{{-25, -12.5, -5 + T}, {25, -12.5, 5 + T}},
                                          {{-25, -10.5, 5}, {25, -1.5, 5}},
                                          {{-7.5, 12.5, -5 + T}, {7.5, 12.5, 5 + T}},
                                          {{-5, 12.4, -2.5 + T}, {5, 12.4, 2.5 + T}}};
        for (int i = 0; i < kV4Quads; i++) {
            Box b;
            for (int a = 0; a < 3; a++) { b.lo[a] = q[i][0][a] - 1e-3; b.hi[a] = q[i][1][a] + 1e-3; }
            if (!project_box(b, kProfileV4, width, height, camDist, &rects[n++])) return -1;
        }
        for (int i = 0; i < kV4Spheres; i++)
            if (!project_box(sphere_box(s.sphere[i]), kProfileV4, width, height, camDist, &rects[n++])) return -1;
    } else if (profile == kProfileV3RedoS0) {  // the box (walls enclose the spheres) and the light far behind it, seen from (0, 0, 40)
        V3RedoScene0 s;
        build_v3redo_scene0(&s);
        Box b = empty_box();
        for (int i = 0; i < kV3S0Quads - 1; i++) { grow(&b, s.quad[i].a); grow(&b, s.quad[i].b); grow(&b, s.quad[i].c); grow(&b, s.quad[i].d); }
        for (int i = 0; i < kV3S0Spheres; i++) { const Box sb = sphere_box(s.sphere[i]); grow(&b, mk((float)sb.lo[0], (float)sb.lo[1], (float)sb.lo[2])); grow(&b, mk((float)sb.hi[0], (float)sb.hi[1], (float)sb.hi[2])); }
        for (int a = 0; a < 3; a++) { b.lo[a] -= 1e-3; b.hi[a] += 1e-3; }
        if (!project_box(b, kProfileV4, width, height, camDist, &rects[n++])) return -1;
        Box l = empty_box();
        const LegacyQuad& L = s.quad[kV3S0Quads - 1];
        grow(&l, L.a); grow(&l, L.b); grow(&l, L.c); grow(&l, L.d);
        for (int a = 0; a < 3; a++) { l.lo[a] -= 1e-3; l.hi[a] += 1e-3; }
        if (!project_box(l, kProfileV4, width, height, camDist, &rects[n++])) return -1;
    } else {
        CornellScene s;
        build_cornell_scene(&s, profile == kProfileSimtTextured);
        // the whole box in one rectangle (walls enclose light and spheres), plus nothing else
        Box b = empty_box();
        for (int i = 0; i < kCornellQuads; i++) { grow(&b, s.quad[i].a); grow(&b, s.quad[i].b); grow(&b, s.quad[i].c); grow(&b, s.quad[i].d); }
        for (int i = 0; i < kCornellSpheres; i++) { const Box sb = sphere_box(s.sphere[i]); grow(&b, mk((float)sb.lo[0], (float)sb.lo[1], (float)sb.lo[2])); grow(&b, mk((float)sb.hi[0], (float)sb.hi[1], (float)sb.hi[2])); }
        for (int a = 0; a < 3; a++) { b.lo[a] -= 1e-3; b.hi[a] += 1e-3; }
        if (!project_box(b, profile, width, height, camDist, &rects[n++])) return -1;
    }
    return n;
}

}  // namespace b200pt
