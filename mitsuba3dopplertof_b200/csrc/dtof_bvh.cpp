// Host-side BVH construction for the flattened scene: one binned-SAH BLAS per shape group in the
// group's own space, and a TLAS over instances whose bounds are the union of the two keyframe boxes
// (Instance::bbox, src/shapes/instance.cpp:101-114 -- exact for a linearly interpolated affine map).
// Replaces Embree's builder (ext/embree) for this path; output layout: dtof_layout.h.
#include "dtof_bvh.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace dtof {

namespace {

struct Box {
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    void grow(const float *p) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], p[a]);
            hi[a] = std::max(hi[a], p[a]);
        }
    }
    void grow(const Box &b) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
    bool valid() const { return lo[0] <= hi[0]; }
};

struct Ref {
    Box box;
    float c[3];
    uint32_t id;
};

// Conservative padding: Moeller-Trumbore accepts hits a few ulp outside the exact triangle, and the slab test
// (b - o) * idir carries ~3 ulp of relative error (absorbed by the widened far side in the kernel).
inline void pad_box(Box &b) {
    for (int a = 0; a < 3; ++a) {
        float m = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a]));
        float pad = 1e-5f * m + 1e-6f * (b.hi[a] - b.lo[a]) + 1e-30f;
        b.lo[a] -= pad;
        b.hi[a] += pad;
    }
}

// centre / half extent of [lo, hi], the half extent rounded up so that [m - h, m + h] contains the interval
inline void centre_half(float lo, float hi, float &m, float &h) {
    if (!(lo <= hi)) {   // empty box: never hit
        m = 0.f, h = -1.f;
        return;
    }
    m = 0.5f * lo + 0.5f * hi;
    h = std::max(hi - m, m - lo);
    h = h * (1.f + 4.f * FLT_EPSILON) + FLT_MIN;
}

inline void set_child(BvhNode &n, int which, const Box &b, int32_t ref) {
    if (which == 0) {
        centre_half(b.lo[0], b.hi[0], n.c0_mx, n.c0_hx);
        centre_half(b.lo[1], b.hi[1], n.c0_my, n.c0_hy);
        centre_half(b.lo[2], b.hi[2], n.c0_mz, n.c0_hz);
        n.child0 = ref;
    } else {
        centre_half(b.lo[0], b.hi[0], n.c1_mx, n.c1_hx);
        centre_half(b.lo[1], b.hi[1], n.c1_my, n.c1_hy);
        centre_half(b.lo[2], b.hi[2], n.c1_mz, n.c1_hz);
        n.child1 = ref;
    }
}

constexpr int kBins = 16;
static uint32_t kMaxLeaf = 4;   // set per scene in build_scene_bvh; DTOF_MAX_LEAF overrides (tuning experiments)
constexpr int kSahDepthLimit = 40;   // beyond this depth fall back to median splits (bounds the traversal stack)

// Generic top-down builder over `refs[lo,hi)`. `make_leaf(lo, hi)` returns the leaf reference.
// Returns the child reference of the subtree root and its (padded) box.
template <typename MakeLeaf>
int32_t build_range(std::vector<Ref> &refs, uint32_t lo, uint32_t hi, std::vector<BvhNode> &nodes, int depth,
                    uint32_t max_leaf, MakeLeaf &make_leaf, Box &out_box, int &max_depth) {
    max_depth = std::max(max_depth, depth);
    Box bounds, cbounds;
    for (uint32_t i = lo; i < hi; ++i) {
        bounds.grow(refs[i].box);
        cbounds.grow(refs[i].c);
    }
    out_box = bounds;
    uint32_t n = hi - lo;
    if (n <= max_leaf) {
        return make_leaf(lo, hi);
    }
    int axis = 0;
    float ext[3] = { cbounds.hi[0] - cbounds.lo[0], cbounds.hi[1] - cbounds.lo[1], cbounds.hi[2] - cbounds.lo[2] };
    if (ext[1] > ext[axis]) axis = 1;
    if (ext[2] > ext[axis]) axis = 2;
    uint32_t mid = lo;
    bool split_done = false;
    if (ext[axis] > 0.f && depth < kSahDepthLimit) {
        // binned SAH over all three axes
        float best_cost = FLT_MAX;
        int best_axis = -1, best_bin = -1;
        for (int a = 0; a < 3; ++a) {
            if (!(ext[a] > 0.f))
                continue;
            Box bb[kBins];
            uint32_t cnt[kBins] = {};
            float k = kBins * (1.f - 1e-6f) / ext[a];
            for (uint32_t i = lo; i < hi; ++i) {
                int b = std::min(kBins - 1, std::max(0, (int) ((refs[i].c[a] - cbounds.lo[a]) * k)));
                cnt[b]++;
                bb[b].grow(refs[i].box);
            }
            float right_area[kBins];
            uint32_t right_cnt[kBins];
            Box acc;
            uint32_t c = 0;
            for (int b = kBins - 1; b > 0; --b) {
                acc.grow(bb[b]);
                c += cnt[b];
                right_area[b] = acc.valid() ? acc.half_area() : 0.f;
                right_cnt[b] = c;
            }
            Box accl;
            uint32_t cl = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                accl.grow(bb[b]);
                cl += cnt[b];
                if (cl == 0 || right_cnt[b + 1] == 0)
                    continue;
                float cost = accl.half_area() * (float) cl + right_area[b + 1] * (float) right_cnt[b + 1];
                if (cost < best_cost) {
                    best_cost = cost;
                    best_axis = a;
                    best_bin = b;
                }
            }
        }
        if (best_axis >= 0) {
            float k = kBins * (1.f - 1e-6f) / ext[best_axis];
            float clo = cbounds.lo[best_axis];
            auto it = std::partition(refs.begin() + lo, refs.begin() + hi, [&](const Ref &r) {
                int b = std::min(kBins - 1, std::max(0, (int) ((r.c[best_axis] - clo) * k)));
                return b <= best_bin;
            });
            mid = (uint32_t) (it - refs.begin());
            split_done = mid > lo && mid < hi;
        }
    }
    if (!split_done) {   // object median along the widest centroid axis
        mid = lo + n / 2;
        std::nth_element(refs.begin() + lo, refs.begin() + mid, refs.begin() + hi,
                         [axis](const Ref &a, const Ref &b) { return a.c[axis] < b.c[axis]; });
    }
    int32_t me = (int32_t) nodes.size();
    nodes.push_back(BvhNode{});
    Box b0, b1;
    int32_t r0 = build_range(refs, lo, mid, nodes, depth + 1, max_leaf, make_leaf, b0, max_depth);
    int32_t r1 = build_range(refs, mid, hi, nodes, depth + 1, max_leaf, make_leaf, b1, max_depth);
    pad_box(b0);
    pad_box(b1);
    BvhNode nd{};
    set_child(nd, 0, b0, r0);
    set_child(nd, 1, b1, r1);
    nodes[me] = nd;
    return me;
}

inline void xf_point(const float *m, const float *p, float *o) {
    for (int r = 0; r < 3; ++r)
        o[r] = m[4 * r + 0] * p[0] + m[4 * r + 1] * p[1] + m[4 * r + 2] * p[2] + m[4 * r + 3];
}

// Large groups (multi-million-triangle meshes): the top of the tree is built serially down to subtrees of
// <= n / kTaskSplit references; the subtrees are independent jobs built by a thread pool into private node /
// triangle arrays (local indices), then appended to the output with their indices rebased.
constexpr uint32_t kParallelMin = 1u << 16;
constexpr uint32_t kTaskSplit = 256;

struct SubtreeJob {
    uint32_t lo, hi;
    int depth;
    std::vector<BvhNode> nodes;
    std::vector<TriIsect> tris;
    int32_t root = 0;
    int max_depth = 0;
};

int32_t build_group_parallel(const GroupInput &G, std::vector<Ref> &refs, BuiltScene &out, int &max_depth) {
    const uint32_t n = (uint32_t) refs.size();
    const uint32_t job_size = std::max<uint32_t>(n / kTaskSplit, 4096);
    std::vector<SubtreeJob> jobs;
    // top of the tree: a "leaf" of the top builder is a job; its reference is a placeholder patched below
    std::vector<BvhNode> top;
    auto make_job = [&](uint32_t lo, uint32_t hi) -> int32_t {
        SubtreeJob j;
        j.lo = lo, j.hi = hi, j.depth = 0;
        jobs.push_back(std::move(j));
        return ~(int32_t) (((uint32_t) (jobs.size() - 1) << 4) | 0u);   // count 0 marks a job placeholder here
    };
    Box bb;
    int top_depth = 0;
    std::vector<int> job_depth;
    // build_range does not pass the depth to make_leaf: recover each job's depth from the top tree afterwards
    int32_t top_root = build_range(refs, 0, n, top, 1, job_size, make_job, bb, top_depth);
    job_depth.assign(jobs.size(), 1);
    {   // depth of every placeholder in the top tree
        struct E { int32_t ref; int depth; };
        std::vector<E> st{ { top_root, 1 } };
        while (!st.empty()) {
            E e = st.back();
            st.pop_back();
            if (e.ref >= 0) {
                st.push_back({ top[e.ref].child0, e.depth + 1 });
                st.push_back({ top[e.ref].child1, e.depth + 1 });
            } else {
                job_depth[(uint32_t) ~e.ref >> 4] = e.depth;
            }
        }
    }
    unsigned n_threads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    std::atomic<size_t> next{ 0 };
    auto worker = [&]() {
        for (;;) {
            size_t k = next.fetch_add(1);
            if (k >= jobs.size())
                break;
            SubtreeJob &j = jobs[k];
            auto make_leaf = [&](uint32_t lo, uint32_t hi) -> int32_t {
                uint32_t first = (uint32_t) j.tris.size();
                for (uint32_t i = lo; i < hi; ++i)
                    j.tris.push_back(G.tris[refs[i].id]);
                return ~(int32_t) ((first << 4) | (hi - lo));
            };
            Box b;
            j.root = build_range(refs, j.lo, j.hi, j.nodes, job_depth[k], kMaxLeaf, make_leaf, b, j.max_depth);
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_threads; ++t)
        pool.emplace_back(worker);
    worker();
    for (auto &t : pool)
        t.join();
    // append: top nodes first, then every job's nodes / triangles with rebased indices
    const int32_t top_base = (int32_t) out.nodes.size();
    out.nodes.insert(out.nodes.end(), top.begin(), top.end());
    std::vector<int32_t> job_ref(jobs.size());
    for (size_t k = 0; k < jobs.size(); ++k) {
        SubtreeJob &j = jobs[k];
        const int32_t node_base = (int32_t) out.nodes.size();
        const uint32_t tri_base = (uint32_t) out.tris.size();
        auto rebase = [&](int32_t ref) -> int32_t {
            if (ref >= 0)
                return ref + node_base;
            uint32_t code = (uint32_t) ~ref;
            return ~(int32_t) ((((code >> 4) + tri_base) << 4) | (code & 15u));
        };
        for (BvhNode nd : j.nodes) {
            nd.child0 = rebase(nd.child0);
            nd.child1 = rebase(nd.child1);
            out.nodes.push_back(nd);
        }
        out.tris.insert(out.tris.end(), j.tris.begin(), j.tris.end());
        job_ref[k] = rebase(j.root);
        max_depth = std::max(max_depth, j.max_depth);
        std::vector<BvhNode>().swap(j.nodes);
        std::vector<TriIsect>().swap(j.tris);
    }
    auto patch = [&](int32_t ref) -> int32_t {
        return ref >= 0 ? ref + top_base : job_ref[(uint32_t) ~ref >> 4];
    };
    for (size_t i = 0; i < top.size(); ++i) {
        BvhNode &nd = out.nodes[top_base + i];
        nd.child0 = patch(nd.child0);
        nd.child1 = patch(nd.child1);
    }
    max_depth = std::max(max_depth, top_depth);
    return patch(top_root);
}

} // namespace

void build_scene_bvh(const std::vector<GroupInput> &groups, BuiltScene &out) {
    // Leaf size, measured on B200 (profiles/r01_tuning.md): scenes whose traversal data sits in shared memory are
    // instruction bound and want few triangle tests per leaf (2); multi-million-triangle scenes are L2/HBM bound and
    // want fewer, fatter nodes (4).
    size_t total_tris = 0;
    for (auto &g : groups)
        total_tris += g.tris.size();
    kMaxLeaf = total_tris <= 4096 ? 2 : 4;
    if (const char *e = getenv("DTOF_MAX_LEAF"))
        kMaxLeaf = (uint32_t) std::min(15, std::max(1, atoi(e)));
    out.nodes.clear();
    out.tris.clear();
    out.inst_root.assign(groups.size(), 0);
    out.max_depth = 0;
    std::vector<Box> group_boxes(groups.size());

    // ---- BLAS per group
    for (size_t g = 0; g < groups.size(); ++g) {
        const GroupInput &G = groups[g];
        uint32_t n = (uint32_t) G.tris.size();
        std::vector<Ref> refs(n);
        Box gb;
        for (uint32_t i = 0; i < n; ++i) {
            const TriIsect &t = G.tris[i];
            float p0[3] = { t.p0x, t.p0y, t.p0z };
            float p1[3] = { G.p1[3 * i], G.p1[3 * i + 1], G.p1[3 * i + 2] };
            float p2[3] = { G.p2[3 * i], G.p2[3 * i + 1], G.p2[3 * i + 2] };
            refs[i].box.grow(p0);
            refs[i].box.grow(p1);
            refs[i].box.grow(p2);
            for (int a = 0; a < 3; ++a)
                refs[i].c[a] = 0.5f * (refs[i].box.lo[a] + refs[i].box.hi[a]);
            refs[i].id = i;
            gb.grow(refs[i].box);
        }
        int depth = 0;
        int32_t root;
        if (n == 0) {
            // empty group: a leaf with zero triangles is not encodable -> point at an empty range via count 0
            root = ~(int32_t) (((uint32_t) out.tris.size()) << 4);
        } else if (n < kParallelMin) {
            auto make_leaf = [&](uint32_t lo, uint32_t hi) -> int32_t {
                uint32_t first = (uint32_t) out.tris.size();
                for (uint32_t i = lo; i < hi; ++i)
                    out.tris.push_back(G.tris[refs[i].id]);
                return ~(int32_t) ((first << 4) | (hi - lo));
            };
            Box bb;
            root = build_range(refs, 0, n, out.nodes, 1, kMaxLeaf, make_leaf, bb, depth);
        } else {
            root = build_group_parallel(G, refs, out, depth);
        }
        out.inst_root[g] = root;
        out.max_depth = std::max(out.max_depth, depth);
        group_boxes[g] = gb;   // empty (invalid) when the group has no triangles
    }

    // ---- TLAS over instances (one instance per leaf), appended after the BLAS nodes
    out.group_box.assign(groups.size(), InstBox{});
    std::vector<TlasEntry> entries(groups.size());
    for (size_t g = 0; g < groups.size(); ++g) {
        // object-space bounds of the group: dtof_update_instances rebuilds the TLAS from them for new keyframes
        const Box &b = group_boxes[g];
        InstBox ob{ 1.f, 1.f, 1.f, 0.f, -1.f, -1.f, -1.f, 0.f };   // lo > hi: empty group
        if (b.valid())
            ob = InstBox{ b.lo[0], b.lo[1], b.lo[2], 0.f, b.hi[0], b.hi[1], b.hi[2], 0.f };
        out.group_box[g] = ob;
        entries[g].animated = groups[g].animated;
        memcpy(entries[g].m0, groups[g].m0, sizeof(entries[g].m0));
        memcpy(entries[g].m1, groups[g].m1, sizeof(entries[g].m1));
        entries[g].object_box = ob;
        entries[g].blas_root = out.inst_root[g];
    }
    out.tlas_begin = (uint32_t) out.nodes.size();
    std::vector<BvhNode> tlas;
    build_tlas(entries, (int32_t) out.tlas_begin, tlas, out);
    out.nodes.insert(out.nodes.end(), tlas.begin(), tlas.end());
}

void build_tlas(const std::vector<TlasEntry> &entries, int32_t base, std::vector<BvhNode> &tlas, BuiltScene &out) {
    const size_t n = entries.size();
    std::vector<Box> inst_boxes(n);
    for (size_t g = 0; g < n; ++g) {
        const TlasEntry &E = entries[g];
        const InstBox &ob = E.object_box;
        Box wb;
        if (ob.lox <= ob.hix) {
            // world bounds of the instance: union of the group's box corners under both keyframes
            if (!E.animated) {
                float lo[3] = { ob.lox, ob.loy, ob.loz }, hi[3] = { ob.hix, ob.hiy, ob.hiz };
                wb.grow(lo);
                wb.grow(hi);
            } else {
                for (int c = 0; c < 8; ++c) {
                    float p[3] = { (c & 1) ? ob.hix : ob.lox, (c & 2) ? ob.hiy : ob.loy, (c & 4) ? ob.hiz : ob.loz };
                    float q[3];
                    xf_point(E.m0, p, q);
                    wb.grow(q);
                    xf_point(E.m1, p, q);
                    wb.grow(q);
                }
            }
        }
        inst_boxes[g] = wb;
    }
    std::vector<Ref> irefs;
    for (size_t g = 0; g < n; ++g) {
        if (!inst_boxes[g].valid())
            continue;
        Ref r;
        r.box = inst_boxes[g];
        for (int a = 0; a < 3; ++a)
            r.c[a] = 0.5f * (r.box.lo[a] + r.box.hi[a]);
        r.id = (uint32_t) g;
        irefs.push_back(r);
    }
    int tdepth = 0;
    tlas.clear();
    if (irefs.empty()) {
        out.root = ~(int32_t) 0x7fffffff;   // never dereferenced: has_geometry = false
        out.has_geometry = false;
    } else {
        // every leaf is first a placeholder ~(g << 4): local node indices (>= 0) can then be rebased to absolute ones
        // without touching the leaves. Afterwards: animated instance -> instance leaf (the placeholder itself);
        // static group -> its BLAS root is spliced in (single-level traversal)
        auto make_ileaf = [&](uint32_t lo, uint32_t) -> int32_t { return ~(int32_t) (irefs[lo].id << 4); };
        Box bb;
        int32_t root = build_range(irefs, 0, (uint32_t) irefs.size(), tlas, 1, 1, make_ileaf, bb, tdepth);
        auto patch = [&](int32_t ref) -> int32_t {
            if (ref >= 0)
                return ref + base;
            const uint32_t g = (uint32_t) ~ref >> 4;
            return entries[g].animated ? ref : entries[g].blas_root;
        };
        for (BvhNode &nd : tlas) {
            nd.child0 = patch(nd.child0);
            nd.child1 = patch(nd.child1);
        }
        out.root = patch(root);
        out.has_geometry = true;
        for (int a = 0; a < 3; ++a) {
            out.scene_lo[a] = bb.lo[a];
            out.scene_hi[a] = bb.hi[a];
        }
    }
    out.tlas_depth = tdepth;
    out.inst_box.assign(n, InstBox{});
    for (size_t g = 0; g < n; ++g) {
        Box b = inst_boxes[g];
        if (b.valid())
            pad_box(b);
        else
            b.lo[0] = b.lo[1] = b.lo[2] = 1.f, b.hi[0] = b.hi[1] = b.hi[2] = -1.f;   // never hit
        out.inst_box[g] = InstBox{ b.lo[0], b.lo[1], b.lo[2], 0.f, b.hi[0], b.hi[1], b.hi[2], 0.f };
    }
}

} // namespace dtof
