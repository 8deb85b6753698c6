// rt_host.hpp -- C++ mirror of the reference's host-side scene/renderer API for the render
// path (the reference is Rust; no Rust toolchain exists in this image, so the host side above
// the C ABI is C++ with the reference's names, argument meaning and behaviour).
//
// Mirrors: Material / TransmissionProperties (src/raytracing/material.rs:14-231),
// TriangleData (src/geometry/basic/triangle.rs:52-113), SphereData (sphere.rs:34-49),
// BoundedPlane (src/geometry/composite/bounded_plane.rs), PointLight::new + maximize_value
// (src/scene/lighting/light.rs:175-181, src/color.rs:125-131), Scene (src/scene/scene.rs),
// the crate constants (src/lib.rs:30-92), the Cargo feature set (Cargo.toml:62-83) and
// RaytracerRenderer::render (src/renderer/raytracer_renderer.rs:1360-1379).
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "rt_api.h"

namespace rth {

// ---------------------------------------------------------------- ultraviolet-style math (f32)
struct Vec3 {
    float x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    static Vec3 unit_x() { return {1, 0, 0}; }
    static Vec3 unit_y() { return {0, 1, 0}; }
    static Vec3 unit_z() { return {0, 0, 1}; }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(float s, Vec3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float dot(Vec3 a, Vec3 b) { return std::fmaf(a.x, b.x, std::fmaf(a.y, b.y, a.z * b.z)); }
inline Vec3 cross(Vec3 a, Vec3 b) {
    return {std::fmaf(a.y, b.z, -(a.z * b.y)), std::fmaf(a.z, b.x, -(a.x * b.z)), std::fmaf(a.x, b.y, -(a.y * b.x))};
}
inline Vec3 normalized(Vec3 a) {
    float r = 1.0f / std::sqrt(dot(a, a));
    return {a.x * r, a.y * r, a.z * r};
}
inline Vec3 mul_add(Vec3 a, Vec3 m, Vec3 add) {
    return {std::fmaf(a.x, m.x, add.x), std::fmaf(a.y, m.y, add.y), std::fmaf(a.z, m.z, add.z)};
}
inline Vec3 lerp(Vec3 a, Vec3 b, float t) { return a * (1.0f - t) + b * t; }

// Rotor3 = (s, xy, xz, yz); conventions verified against the reference render in SURVEY.md A.14
struct Rotor3 {
    float s = 1, xy = 0, xz = 0, yz = 0;
    static Rotor3 identity() { return {}; }
    static Rotor3 from_rotation_xy(float a) { return {std::cos(a * 0.5f), -std::sin(a * 0.5f), 0, 0}; }
    static Rotor3 from_rotation_xz(float a) { return {std::cos(a * 0.5f), 0, -std::sin(a * 0.5f), 0}; }
    static Rotor3 from_rotation_yz(float a) { return {std::cos(a * 0.5f), 0, 0, -std::sin(a * 0.5f)}; }
    static Rotor3 from_euler_angles(float roll, float pitch, float yaw) {
        return from_rotation_xz(yaw) * from_rotation_yz(pitch) * from_rotation_xy(roll);
    }
    Rotor3 operator*(const Rotor3 &q) const {
        Rotor3 r;
        r.s = s * q.s - xy * q.xy - xz * q.xz - yz * q.yz;
        r.xy = xy * q.s + s * q.xy + yz * q.xz - xz * q.yz;
        r.xz = xz * q.s + s * q.xz - yz * q.xy + xy * q.yz;
        r.yz = yz * q.s + s * q.yz + xz * q.xy - xy * q.xz;
        return r;
    }
    Vec3 rotate(Vec3 v) const {
        float fx = s * v.x + xy * v.y + xz * v.z;
        float fy = s * v.y - xy * v.x + yz * v.z;
        float fz = s * v.z - xz * v.x - yz * v.y;
        float fw = xy * v.z - xz * v.y + yz * v.x;
        return {s * fx + xy * fy + xz * fz + yz * fw, s * fy - xy * fx - xz * fw + yz * fz,
                s * fz + xy * fw - xz * fx - yz * fy};
    }
};
inline Vec3 rotated_by(Vec3 v, const Rotor3 &r) { return r.rotate(v); }

struct Isometry3 {
    Vec3 translation;
    Rotor3 rotation;
    Isometry3(Vec3 t, Rotor3 r) : translation(t), rotation(r) {}
    Vec3 transform_vec(Vec3 v) const { return rotation.rotate(v) + translation; }
};
struct Similarity3 {
    Vec3 translation;
    Rotor3 rotation;
    float scale = 1;
    Similarity3() = default;
    Similarity3(Vec3 t, Rotor3 r, float s) : translation(t), rotation(r), scale(s) {}
    static Similarity3 identity() { return {}; }
    Vec3 transform_vec(Vec3 v) const { return rotation.rotate(v * scale) + translation; }
};

using ColorType = Vec3;  // LinSrgb<f32>

// ---------------------------------------------------------------- crate constants (src/lib.rs)
struct SceneConstants {
    uint32_t window_width, window_height;
    float aspect, scene_width, scene_height, scene_depth, average_scene_dimension;
    float w2s_w, w2s_h, w2s_d;
    Vec3 render_ray_focus;
    explicit SceneConstants(uint32_t w, uint32_t h);
};

// ---------------------------------------------------------------- Cargo features (Cargo.toml:62-83)
struct Features {
    bool simd_render = false, anti_aliasing = false, anti_aliasing_rotation_scale = false,
         anti_aliasing_randomness = false, high_resolution = false, medium_resolution = false,
         soft_shadows = false, reflections = false, light_reflections = false, refractions = false,
         backface_culling = false, scene_backface_culling = false, save_rendering_image = false,
         realistic = false, high_quality_model = false, high_quality = false, extreme_quality = false;
    // "default" expands to the crate's default set; implications are resolved like cargo does.
    static bool parse(const std::string &csv, Features &out, std::string &err);
    uint32_t window_width() const { return high_resolution ? 1620 : (medium_resolution ? 1140 : 768); }
    uint32_t window_height() const { return high_resolution ? 1350 : (medium_resolution ? 950 : 640); }
    uint32_t reflection_max_depth() const { return high_quality ? (extreme_quality ? 21 : 13) : 9; }
    uint32_t refraction_max_depth() const { return high_quality ? (extreme_quality ? 21 : 18) : 8; }
    uint32_t point_light_multiplicator() const {
        return soft_shadows ? (high_quality ? (extreme_quality ? 28 : 19) : 10) : 1;
    }
    uint32_t antialiasing_samples_per_pixel() const { return extreme_quality ? 24 : 9; }
    uint32_t render_flags() const;
    // Output::get_feature_string (src/output/mod.rs:25-88) and the window title built from it (src/output/window.rs:58-64)
    std::string feature_string() const;
    std::string window_title() const { return "Raytracer - " + feature_string(); }
};

// ---------------------------------------------------------------- materials & geometry
struct TransmissionProperties {
    float refraction_index = 1.0f;  // Default::default()
    float opacity = 0.0f;
    bool has_opacity = false;
    float boost = 0.0f;
    static TransmissionProperties new_(float opacity, float refraction_index) { return {refraction_index, opacity, true, 0.0f}; }
    static TransmissionProperties new_with_boost(float opacity, float refraction_index, float boost) {
        return {refraction_index, opacity, true, boost};
    }
    static TransmissionProperties none() { return {0.0f, 0.0f, false, 0.0f}; }
    bool mask() const { return has_opacity && !(std::fabs(opacity - 0.0f) <= 1.1920929e-7f); }
};

struct Material {
    ColorType color{0, 0, 0};
    float metallic = 0, shininess = 0;
    TransmissionProperties transmission;
    Material() = default;
    Material(ColorType c, float m, float s, TransmissionProperties t) : color(c), metallic(m), shininess(s), transmission(t) {}
    static Material diffuse(ColorType c) { return Material(c, 0, 0, TransmissionProperties()); }
    rt_material flat() const;
};

struct TriangleData {
    Vec3 vertex1, vertex2, vertex3, normal;
    Material material;
    static TriangleData with_material(Vec3 v1, Vec3 v2, Vec3 v3, Material m) {
        Vec3 e1 = v2 - v1, e2 = v3 - v1;
        return {v1, v2, v3, normalized(cross(e1, e2)), m};
    }
    static TriangleData with_material_and_normal(Vec3 v1, Vec3 v2, Vec3 v3, Vec3 n, Material m) { return {v1, v2, v3, n, m}; }
};

struct SphereData {
    Vec3 center;
    float radius;
    Material material;
    static SphereData with_material(Vec3 c, float r, Material m) { return {c, r, m}; }
    static SphereData new_(Vec3 c, float r, ColorType col) { return {c, r, Material::diffuse(col)}; }
};

ColorType maximize_value(ColorType c);  // src/color.rs:125-131

struct PointLight {
    Vec3 position;
    ColorType color;
    float intensity;
    static PointLight new_(Vec3 p, ColorType c, float i) { return {p, maximize_value(c), i}; }
};

struct BoundedPlane {
    Vec3 center, up, left, normal;
    float width, height, depth;
    Material material;
    // panics in the reference -> returns false + message here
    static bool with_material(Vec3 normal, Vec3 center, Vec3 up, float width, float height, float depth, Material m,
                              BoundedPlane &out, std::string &err);
    void triangulate(Vec3 t1[3], Vec3 t2[3]) const;
    std::vector<TriangleData> to_basic_geometries() const;
};

// A polygon mesh as tobj hands it over (single_index corners): per face corner a position and
// (optionally) a normal index.
struct Mesh {
    std::vector<float> positions;  // 3 per vertex
    std::vector<float> normals;    // 3 per normal (may be empty)
    std::vector<int32_t> face_v;   // 3 per triangle
    std::vector<int32_t> face_vn;  // 3 per triangle, or empty
    std::vector<int32_t> face_material;  // per triangle index into materials, -1 = none
    std::vector<Material> materials;
};
bool load_obj(const std::string &path, Mesh &out, std::string &err);  // tobj::load_obj(triangulate, single_index)

struct Scene {
    // GeometryCollection: kind -> Vec (iteration order fixed to spheres, then triangles)
    std::vector<SphereData> spheres;
    std::vector<TriangleData> triangles;
    std::vector<PointLight> scene_lights;

    void add_sphere(const SphereData &s) { spheres.push_back(s); }
    void add_triangle(const TriangleData &t) { triangles.push_back(t); }
    void add_light(const PointLight &l) { scene_lights.push_back(l); }
    void merge(const Scene &o);
    static bool from_obj(const std::string &path, const Similarity3 *transform, Scene &out, std::string &err);
    static Scene from_mesh(const Mesh &mesh, const Similarity3 *transform);
    static Scene backface_culling(const Scene &scene, Vec3 view_direction);
    size_t object_count() const { return spheres.size() + triangles.size(); }

    // flatten into the C-ABI arrays (materials de-duplicated bitwise)
    void flatten(std::vector<rt_triangle> &t, std::vector<rt_sphere> &s, std::vector<rt_material> &m, std::vector<rt_light> &l) const;
};

// The reference's entry-point scenes (src/main.rs == examples/semesterbild.rs, examples/test_scene.rs,
// examples/test_text.rs). `text_mesh` is the parsed text.obj / text_lowres.obj.
Scene make_semesterbild_scene(const Features &f, const SceneConstants &k, const Mesh &text_mesh);
Scene make_test_scene(const SceneConstants &k);
Scene make_test_text_scene(const SceneConstants &k, const Mesh &text_mesh);

// ImageBuffer<W,H> (src/image_buffer.rs:9-46): row-major u32, zero-initialised.
struct ImageBuffer {
    uint32_t width, height;
    std::vector<uint32_t> buffer;
    ImageBuffer(uint32_t w, uint32_t h) : width(w), height(h), buffer((size_t)w * h, 0u) {}
    const uint32_t *get_u32_slice() const { return buffer.data(); }
};

// RaytracerRenderer::render: flattens the scene, creates the device scene, builds the sampler
// tables from `seed`, and calls rt_render. No CPU fallback: fails with the rt_status of the C ABI.
struct RaytracerRenderer {
    Features features;
    uint64_t seed = 0;
    int device = 0;
    uint32_t n_light_clouds = 4096;
    int render(ImageBuffer &buffer, const Scene &scene, rt_stats *stats = nullptr) const;
};

}  // namespace rth
