/*
 * rt_host.h -- C wrappers over the C++ host-side mirror of the reference's scene API
 * (hslu_i/ba_raytracing/f2501_raytracer_b200/csrc/host/rt_host.hpp).  They exist so that the
 * parity tests and bench.py (Python, ctypes) can build exactly the scenes the reference's entry
 * points build and flatten them into the rt_api.h arrays.  The Rust shim does NOT need these: it
 * flattens its own `Scene<Vec3>`.
 *
 * Reference interfaces mirrored: Scene::{new,add_sphere,add_triangle,add_light,from_obj,
 * backface_culling,merge} (src/scene/scene.rs:29-178), BoundedPlane::with_material +
 * to_basic_geometries (src/geometry/composite/bounded_plane.rs:49-216), PointLight::new
 * (src/scene/lighting/light.rs:175-181), the scenes of src/main.rs:26-320,
 * examples/test_scene.rs:22-343, examples/test_text.rs:24-52 and the Cargo feature set
 * (Cargo.toml:62-83).
 */
#ifndef RT_HOST_H
#define RT_HOST_H

#include "rt_api.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rth_scene rth_scene;

rth_scene *rth_scene_new(void);
void rth_scene_free(rth_scene *s);

int rth_scene_add_sphere(rth_scene *s, const float center[3], float radius, const rt_material *m);
/* TriangleData::with_material: normal = normalize(e1 x e2) */
int rth_scene_add_triangle(rth_scene *s, const float v1[3], const float v2[3], const float v3[3], const rt_material *m);
/* TriangleData::with_material_and_normal */
int rth_scene_add_triangle_with_normal(rth_scene *s, const float v1[3], const float v2[3], const float v3[3],
                                       const float normal[3], const rt_material *m);
/* BoundedPlane::with_material(...).to_basic_geometries(): 12 triangles */
int rth_scene_add_bounded_plane(rth_scene *s, const float normal[3], const float center[3], const float up[3],
                                float width, float height, float depth, const rt_material *m);
/* PointLight::new: colour goes through maximize_value */
int rth_scene_add_light(rth_scene *s, const float position[3], const float color[3], float intensity);
/* Scene::from_obj(path, Some(Similarity3::new(translation, Rotor3::from_euler_angles(roll,pitch,yaw), scale)))
 * merged into `s`. */
int rth_scene_add_obj(rth_scene *s, const char *path, const float translation[3], const float euler_rpy[3], float scale);
/* Same ingestion for an already parsed mesh (the committed assets/meshes fixtures). face_vn may be NULL. */
int rth_scene_add_mesh(rth_scene *s, const float *positions, uint32_t n_positions, const float *normals,
                       uint32_t n_normals, const int32_t *face_v, const int32_t *face_vn, uint32_t n_faces,
                       const rt_material *m, const float translation[3], const float euler_rpy[3], float scale);
/* Scene::backface_culling(scene, view_direction) in place */
int rth_scene_backface_culling(rth_scene *s, const float view_direction[3]);

int rth_scene_counts(const rth_scene *s, uint32_t *n_triangles, uint32_t *n_spheres, uint32_t *n_materials, uint32_t *n_lights);
int rth_scene_flatten(const rth_scene *s, rt_triangle *triangles, rt_sphere *spheres, rt_material *materials, rt_light *lights);

/* The reference's entry-point scenes: name = "semesterbild" (== "main"), "test_scene", "test_text".
 * The text mesh (text.obj / text_lowres.obj content) is passed parsed; ignored by test_scene. */
int rth_example_scene(const char *name, const char *features_csv, const float *positions, uint32_t n_positions,
                      const float *normals, uint32_t n_normals, const int32_t *face_v, const int32_t *face_vn,
                      uint32_t n_faces, rth_scene **out);

/* Resolve a Cargo feature list ("default", "high_quality,simd_render", "" ...) into an rt_config
 * (width, height, flags, depths, M, aa sizes; table pointers left NULL). Also reports which text mesh the
 * entry point loads (1 = text.obj, 0 = text_lowres.obj) and whether scene_backface_culling is on. */
int rth_config_from_features(const char *features_csv, rt_config *cfg, int *high_quality_model, int *scene_backface_culling,
                             int *anti_aliasing_randomness);

/* Output::get_feature_string() for a feature list (src/output/mod.rs:25-88), e.g.
 * "SIMD | Antialiasing ROS_SCL RNG | Reflections + Refractions | High Quality | Medium Resolution (1140x950x1045) | NO-OPT"
 * (with U+00D7 multiplication signs, UTF-8); with_title_prefix != 0: the window title "Raytracer - ..." the
 * reference sets (src/output/window.rs:58-64).  NUL-terminated into out[out_size]. */
int rth_feature_string(const char *features_csv, int with_title_prefix, char *out, uint32_t out_size);

/* RaytracerRenderer::render(&buffer, &scene) for a feature list: flatten, rt_scene_create on `device`,
 * sampler tables from `seed`, rt_render into out_argb (width*height of the feature set). */
int rth_render(const rth_scene *s, const char *features_csv, uint64_t seed, int device, uint32_t *out_argb,
               uint8_t *out_hit_mask, rt_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* RT_HOST_H */
