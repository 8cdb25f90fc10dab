// scene_setup.h -- host-side scene tables (see scene_setup.cpp).
#pragma once
#include "../csrc/pt_common.cuh"

namespace b200pt {
float camera_distance();
void build_cornell_scene(CornellScene* s, bool simt_textured_materials);
void build_v4_scene(V4Scene* s);
}  // namespace b200pt
