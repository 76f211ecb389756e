// Host-side BVH builder interface (see dtof_bvh.cpp).
#pragma once
#include <vector>

#include "dtof_layout.h"

namespace dtof {

struct GroupInput {
    std::vector<TriIsect> tris;   // in scene-description order, gid filled in
    std::vector<float> p1, p2;    // 3 floats per triangle: the other two vertices (bounds only)
    bool animated = false;
    float m0[12], m1[12];
};

struct BuiltScene {
    std::vector<BvhNode> nodes;   // all BLAS nodes followed by the TLAS nodes; child indices are absolute
    std::vector<TriIsect> tris;   // leaf order
    std::vector<int32_t> inst_root;
    std::vector<InstBox> inst_box;   // padded world bounds per instance
    int32_t root = 0;             // child reference of the TLAS root
    int max_depth = 0, tlas_depth = 0;
    bool has_geometry = false;
    float scene_lo[3] = { 0, 0, 0 }, scene_hi[3] = { 0, 0, 0 };
};

void build_scene_bvh(const std::vector<GroupInput> &groups, BuiltScene &out);

} // namespace dtof
