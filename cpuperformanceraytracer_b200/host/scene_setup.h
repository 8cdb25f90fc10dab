// scene_setup.h -- host-side scene tables (see scene_setup.cpp).
#pragma once
#include "../csrc/pt_common.cuh"

namespace b200pt {
float camera_distance();
void build_cornell_scene(CornellScene* s, bool simt_textured_materials);
void build_v4_scene(V4Scene* s);
void build_v3redo_scene(V3RedoScene* s);
// Conservative fragCoord-space rectangles (x0, y0, x1, y1) that contain the projection of every
// primitive of the profile's scene, expanded by a safety margin.  Returns the number of rectangles,
// or -1 when culling is not possible (a primitive reaches behind the camera).
int compute_cull_rects(int profile, int width, int height, float4* rects);
}  // namespace b200pt
