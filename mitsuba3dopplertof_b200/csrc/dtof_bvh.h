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
    std::vector<InstBox> group_box;  // object-space bounds per group (lo > hi: empty), input of a TLAS rebuild
    uint32_t tlas_begin = 0;      // index of the first TLAS node in `nodes`
    int32_t root = 0;             // child reference of the TLAS root
    int max_depth = 0, tlas_depth = 0;
    bool has_geometry = false;
    float scene_lo[3] = { 0, 0, 0 }, scene_hi[3] = { 0, 0, 0 };
};

void build_scene_bvh(const std::vector<GroupInput> &groups, BuiltScene &out);

// One instance of the TLAS: its motion, the object-space bounds of its group and the BLAS root.
struct TlasEntry {
    bool animated = false;
    float m0[12], m1[12];
    InstBox object_box;
    int32_t blas_root = 0;
};

// Builds the TLAS over `entries` into `tlas` (node indices absolute, the first node being `base`) and sets
// out.root / has_geometry / tlas_depth / scene bounds / inst_box. A TLAS over k non-empty instances always has k - 1
// nodes, so a rebuild for new keyframes (dtof_update_instances) overwrites the old nodes in place.
void build_tlas(const std::vector<TlasEntry> &entries, int32_t base, std::vector<BvhNode> &tlas, BuiltScene &out);

} // namespace dtof
