// rt_bvh.h -- host-side BVH4 builder and the device memory layout of the flattened scene.
//
// The reference has NO acceleration structure: Raytracer::cast_ray / has_any_intersection scan
// every object (src/raytracing/raytracer.rs:48,180).  The BVH is new work; it must return exactly
// the hits of the linear scan, so node boxes are padded well beyond the rounding error of the
// primitive tests and the traversal's slab test only ever culls boxes that are clearly missed
// (DESIGN.md "Conservative traversal").
#pragma once
#include <cstdint>
#include <vector>

#include "rt_api.h"

namespace rtb {

// ---- device layout (all arrays are float4 = 16 B so every fetch is one LDG.128) -----------------
//
// Primitive record, 64 B = 4 x float4, stored in BVH leaf order ("slot" order):
//   triangle: q0 = (v1.x, v1.y, v1.z, e1.x)  q1 = (e1.y, e1.z, e2.x, e2.y)
//             q2 = (e2.z, g.x, g.y, g.z)  with g = cross(e1, e2) (the first row of the inverse,
//                  identical bit for bit to cross(-e1, -e2) evaluated per ray)
//   sphere:   q0 = (c.x, c.y, c.z, r*r)      q1, q2 unused
//   both:     q3 = (bits(prim_index), bits(meta), bits(cell), triangle: 64 eps * max|edge component|)
//             meta = material index | PRIM_TRANSMISSIVE | PRIM_SPHERE
//             cell = 0, or TRI_CELL(k, i, j, upper) when the slot is one cell of a large triangle
//                    that is referenced once per cell of a k x k barycentric subdivision
// prim_index is the reference iteration order (spheres first, then triangles): closest-hit ties go
// to the larger index (raytracer.rs:193-194).
constexpr uint32_t PRIM_MATERIAL_MASK = 0x00FFFFFFu;
constexpr uint32_t PRIM_TRANSMISSIVE = 1u << 30;
constexpr uint32_t PRIM_SPHERE = 1u << 31;
// cell tag: k in bits 0..6 (2..32), i in 8..14, j in 16..22, upper (inverted sub-triangle) in bit 24
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr uint32_t TRI_CELL(uint32_t k, uint32_t i, uint32_t j, uint32_t upper) { return k | (i << 8) | (j << 16) | (upper << 24); }

// BVH4 node, 128 B = 8 x float4:
//   n0..n5 = lo.x, lo.y, lo.z, hi.x, hi.y, hi.z of the four children (SoA: .x = child 0 ...)
//   n6     = bits of the four child references
//   n7     = unused (keeps the node 128-byte aligned = one L1 line)
// child reference: CHILD_EMPTY, or internal node index (bit 31 clear), or a leaf
//   (bit 31 set | (count-1) << 27 | first_slot), count in 1..MAX_LEAF.
constexpr uint32_t CHILD_EMPTY = 0x7FFFFFFFu;
constexpr uint32_t CHILD_LEAF = 1u << 31;
constexpr uint32_t MAX_LEAF = 4;
constexpr uint32_t LEAF_FIRST_MASK = (1u << 27) - 1;

struct FlatScene {
    std::vector<float> nodes;    // 32 floats per node
    std::vector<float> prims;    // 16 floats per primitive slot
    std::vector<float> shade_n;  // 4 floats per slot: stored triangle normal (spheres: 0)
    std::vector<float> mats;     // 12 floats per material: (color, metallic) (shininess, ior, opacity, boost) (bits has_opacity, bits transmissive, f0 against ior 1.0, 1 / ior)
    std::vector<float> lights;   // 8 floats per light: (pos, intensity) (color, 0)
    uint32_t n_nodes = 0, n_prims = 0, n_mats = 0, n_lights = 0;  // n_prims counts slots (references)
    uint32_t root_ref = CHILD_EMPTY;  // child reference of the root (a leaf ref for tiny scenes)
    // what a ray with a NaN direction "hits" in the reference (every triangle test passes with
    // t = NaN, so the linear scan keeps the first eligible triangle): slot of the lowest-index
    // triangle, and of the lowest-index transmissive triangle (per-ray back-face term on); -1 if none.
    int32_t nan_hit_slot = -1, nan_hit_slot_culled = -1;
    float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};
    float box_pad = 0;  // absolute padding applied to every primitive box
    uint32_t max_depth = 0;
};

// pointer / count checks of a scene description; materials and lights in device layout (shared with the device builder)
int check_scene_desc(const rt_scene_desc *desc);
void pack_materials_and_lights(const rt_scene_desc *desc, FlatScene &out, std::vector<uint8_t> &transmissive);

// Validates `desc` (indices, pointers) and builds the flattened scene. Returns RT_OK or an error
// (message via rt_set_error).
int build_flat_scene(const rt_scene_desc *desc, FlatScene &out);

}  // namespace rtb
