// scene_setup.h -- host-side scene tables (see scene_setup.cpp).
#pragma once
#include "../csrc/pt_common.cuh"

namespace b200pt {
float camera_distance();
void build_cornell_scene(CornellScene* s, bool simt_textured_materials);
void build_v4_scene(V4Scene* s);
void build_v3redo_scene(V3RedoScene* s);
void build_v3redo_scene0(V3RedoScene0* s);
// AddQuadObjectToScene / AddSphereObjectToScene / AddMaterialToScene (v4.cpp:1368-1401) on caller data:
// quads as 4 vertices (12 floats each), spheres as xyz + radius, materials in the 17-float order of
// SceneMaterial (v4.cpp:351-362); like AddMaterialToScene, albedo.y / albedo.z are replaced by albedo.x.
bool build_v4_scene_from(V4Scene* s, const float* quad_vertices, int num_quads, const float* spheres, int num_spheres,
                         const float* materials, const float camera_position[3], float camera_distance);
// The Cornell-family scene (v2.cpp:320-454: six quads, three spheres, nine materials) from caller data: quads as 4
// vertices (12 floats each, already translated), spheres as xyz + radius, materials as 11 floats in LegacyMaterial order
// (albedo3, emissive3, specularColor3, percentSpecular, roughness).  The quad normals are the reference's per-ray
// expression normalize(cross(c - a, c - b)) (v2.cpp:166) evaluated once.  false: a coordinate outside +-1e6 / a radius
// outside [1e-3, 1e6] (the kernels' unchecked reciprocal / division sequences are proven for scene-scale operands only).
bool build_cornell_scene_from(CornellScene* s, const float* quad_vertices, const float* spheres, const float* materials);
int compute_cull_rects_cornell(const float* quad_vertices, const float* spheres, int width, int height, float4* rects);
// culling rectangles for an arbitrary v4-profile scene (quad vertices / spheres as above)
int compute_cull_rects_v4(const float* quad_vertices, int num_quads, const float* spheres, int num_spheres,
                          const float camera_position[3], float camera_distance, int width, int height, float4* rects);
// Conservative fragCoord-space rectangles (x0, y0, x1, y1) that contain the projection of every
// primitive of the profile's scene, expanded by a safety margin.  Returns the number of rectangles,
// or -1 when culling is not possible (a primitive reaches behind the camera).
int compute_cull_rects(int profile, int width, int height, float4* rects);
}  // namespace b200pt
