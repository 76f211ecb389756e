// Device-memory layout of a flattened scene (shared by the host builder and the CUDA kernels).
//
// Everything a ray touches during traversal is a 16-byte-aligned record fetched with 128-bit loads:
//   BvhNode   64 B  binary BVH node holding BOTH children's boxes (so one fetch decides both sides)
//   TriIsect  48 B  p0 + the two edges + the triangle's global id (Moeller-Trumbore inputs, mesh.h:342-365)
//   InstRec  128 B  per-instance motion: two 3x4 keyframes, key times, BLAS root, triangle range
// and, once per hit,
//   TriShade  96 B  p1, p2, vertex normals, uvs, mesh id/flags (Mesh::compute_surface_interaction inputs)
// These sizes are the per-unit figures of the traversal roofline.
#pragma once
#include <stdint.h>

namespace dtof {

// Child reference encoding: ref >= 0 -> inner node index; ref < 0 -> leaf with code = ~ref:
//   triangle leaf : code = (first_tri << 4) | count, count in [1, 15]
//   instance leaf : code = (instance  << 4) | 0      (animated instances only; the static group's BLAS is linked
//                                                     into the TLAS directly, so static geometry is single-level)
// A child box is stored per axis as centre m and half extent h (rounded up: [m - h, m + h] contains the padded box):
// the traversal's interval per axis is then c -/+ h |1/d| with c = (m - o) / d, without a min / max per axis
// (dtof_device.cuh: node_test). An empty box (never hit) has h < 0.
struct alignas(16) BvhNode {
    float c0_mx, c0_hx, c0_my, c0_hy;   // child 0 box x/y
    float c1_mx, c1_hx, c1_my, c1_hy;   // child 1 box x/y
    float c0_mz, c0_hz, c1_mz, c1_hz;   // z of both
    int32_t child0, child1, pad0, pad1;
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

struct alignas(16) TriIsect {
    float p0x, p0y, p0z;
    uint32_t gid;          // global triangle id (order of the scene description) -> deterministic tie-break + shading lookup
    float e1x, e1y, e1z, pad1;
    float e2x, e2y, e2z, pad2;
};
static_assert(sizeof(TriIsect) == 48, "TriIsect must be 48 bytes");

enum : uint32_t { TRI_HAS_NORMALS = 1u, TRI_HAS_UV = 2u, TRI_FLIP = 4u };

struct alignas(16) TriShade {   // indexed by gid
    float p0x, p0y, p0z;
    uint32_t mesh;
    float p1x, p1y, p1z;
    uint32_t flags;
    float p2x, p2y, p2z, pad;
    float n0x, n0y, n0z, n1x;
    float n1y, n1z, n2x, n2y;
    float n2z, uv0x, uv0y, uv1x;
    float uv1y, uv2x, uv2y, pad2;
};
static_assert(sizeof(TriShade) == 112, "TriShade must be 112 bytes");

struct alignas(16) InstRec {
    float m0[12];
    float m1[12];
    float t0, t1;
    int32_t root;          // child reference of the BLAS root
    uint32_t animated;
    uint32_t first_tri;    // range in the scene-order triangle array (flat traversal of tiny scenes)
    uint32_t n_tris;
    uint32_t pad0, pad1;
};
static_assert(sizeof(InstRec) == 128, "InstRec must be 128 bytes");

struct alignas(16) InstBox {   // padded world bounds over both keyframes (instance culling in the flat traversal)
    float lox, loy, loz, pad0;
    float hix, hiy, hiz, pad1;
};

struct MeshRec {
    uint32_t bsdf;
    int32_t emitter;
    uint32_t kind;         // dtof_shape_kind
    uint32_t n_faces;
    // area-emitter sampling (Rectangle::sample_position / Mesh::sample_position)
    float rect_to_world[12];
    float rect_n[3];
    float inv_area;        // rectangle: 1/area; mesh: area pmf normalisation
    float area_sum;
    uint32_t cdf_offset;   // into the area cdf/pmf arrays (mesh emitters)
    uint32_t valid_lo, valid_hi;
    uint32_t first_gid;    // global id of the mesh's face 0
    uint32_t pad[3];
};

struct alignas(16) BsdfRec {
    float r, g, b;         // diffuse reflectance / conductor specular_reflectance
    uint32_t flags;        // bit0: twosided, bit1: smooth diffuse lobe present, bit2: smooth conductor (delta reflection),
                           // bit3: smooth dielectric (delta reflection + refraction), bit4: thin dielectric (+ bit3),
                           // bit5: smooth plastic (+ bit1: its diffuse base is a smooth lobe),
                           // bit6: rough conductor (+ bit1: glossy lobe), bit7: its microfacet distribution is GGX
                           // (pad0 / pad1 = alpha_u / alpha_v), bit8: rough dielectric (+ bit1, bit7; eta_r = eta,
                           // k = specular_transmittance)
    float eta_r, eta_g, eta_b, pad0;   // conductor: complex index of refraction eta + i k; dielectric: eta_r = int_ior / ext_ior;
                                       // plastic: eta, fdr_int, 1 / eta^2, specular sampling weight (plastic.cpp:193-208)
    float k_r, k_g, k_b, pad1;         // dielectric: specular_transmittance; plastic: specular_reflectance, nonlinear (0 / 1)
};

struct alignas(16) SpotRec {   // per emitter index, read only for DTOF_EMITTER_SPOT (src/emitters/spot.cpp)
    float to_local[9];
    float cutoff_angle, cos_cutoff, cos_beam, inv_transition;
    float pad[3];
};

struct EmitterRec {
    uint32_t kind;         // dtof_emitter_kind
    uint32_t mesh;
    float px, py, pz;
    float vr, vg, vb;
};

} // namespace dtof
