/*
 * rt_api.h -- C ABI of librt_b200.so, the B200 (sm_100a) render core that replaces the
 * per-pixel render loop of kije/HSLU_I.BA_RAYTRACING.F2501_raytracer.
 *
 * The reference exposes no FFI.  Its seam for this path is the trait method
 *     fn render(&self, buffer: &ImageBuffer<W,H>, scene: &Scene<Vec3>)
 *         (reference src/renderer/mod.rs:92-94, implemented at
 *          src/renderer/raytracer_renderer.rs:1360-1379)
 * called from src/main.rs:332, examples/semesterbild.rs:333, examples/test_scene.rs:354.
 * A Rust shim `impl Renderer<W,H,C> for GpuRaytracerRenderer<C>` (bindings/rust/gpu_renderer.rs,
 * INTEGRATION.md) flattens `&Scene<Vec3>` into the plain structs below and calls these entry
 * points.  Everything is plain pointers + sizes; no C++/torch types cross the boundary.
 *
 * Conventions: every function returns RT_OK (0) or a negative rt_status; nothing aborts or
 * throws across the boundary; rt_last_error() gives the thread-local message of the last
 * failure.  Host arrays are borrowed for the duration of the call only.  There is no CPU
 * fallback: without a CUDA device every compute entry point returns RT_ERR_NO_DEVICE.
 */
#ifndef RT_API_H
#define RT_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_API_VERSION 1

typedef enum rt_status {
    RT_OK = 0,
    RT_ERR_INVALID_ARGUMENT = -1,
    RT_ERR_NO_DEVICE = -2,
    RT_ERR_CUDA = -3,
    RT_ERR_OUT_OF_MEMORY = -4,
    RT_ERR_IO = -5
} rt_status;

/* ---- scene description (what the Rust shim flattens `Scene<Vec3>` into) ------------------- */

/* Material<f32> + TransmissionProperties<f32> (reference src/raytracing/material.rs:14-94).
 * `ior`, `opacity`, `boost` are the raw stored fields; `has_opacity` is the SimdOption mask of
 * `opacity` (0 for TransmissionProperties::none()/default()). */
typedef struct rt_material {
    float color[3];
    float metallic;
    float shininess;
    float ior;
    float opacity;
    uint32_t has_opacity;
    float boost;
} rt_material;

/* TriangleData<Vec3> (reference src/geometry/basic/triangle.rs:22-48): the three vertices and
 * the stored normal exactly as the reference holds them (the normal may be non-unit for OBJ
 * faces, src/scene/scene.rs:102-126).  edge1/edge2 are recomputed as v2-v1, v3-v1. */
typedef struct rt_triangle {
    float v1[3];
    float v2[3];
    float v3[3];
    float normal[3];
    uint32_t material; /* index into rt_scene_desc.materials */
} rt_triangle;

/* SphereData<Vec3> (reference src/geometry/basic/sphere.rs:20-48); r_sq = radius*radius. */
typedef struct rt_sphere {
    float center[3];
    float radius;
    uint32_t material;
} rt_sphere;

/* PointLight<Vec3> (reference src/scene/lighting/light.rs:160-181); `color` is the stored
 * colour, i.e. already passed through maximize_value (src/color.rs:125-131). */
typedef struct rt_light {
    float position[3];
    float color[3];
    float intensity;
} rt_light;

/* Primitive order = the reference's iteration order `GeometryCollection::get_all()`
 * (src/geometry/render_geometry.rs:146-148) with the kind order fixed to spheres, then
 * triangles; within a kind, insertion order.  Closest-hit ties go to the later primitive
 * (src/raytracing/raytracer.rs:193-194). */
typedef struct rt_scene_desc {
    const rt_triangle *triangles;
    uint32_t n_triangles;
    const rt_sphere *spheres;
    uint32_t n_spheres;
    const rt_material *materials;
    uint32_t n_materials;
    const rt_light *lights;
    uint32_t n_lights;
} rt_scene_desc;

/* ---- render configuration (the reference's cfg!() feature switches, made runtime) ---------- */

enum {
    RT_FLAG_ANTI_ALIASING = 1u << 0,            /* feature anti_aliasing */
    RT_FLAG_AA_ROTATION_SCALE = 1u << 1,        /* feature anti_aliasing_rotation_scale */
    RT_FLAG_REFLECTIONS = 1u << 2,              /* feature reflections */
    RT_FLAG_REFRACTIONS = 1u << 3,              /* feature refractions */
    RT_FLAG_BACKFACE_CULLING = 1u << 4,         /* feature backface_culling (per-ray term) */
    RT_FLAG_PACKET_SEMANTICS = 1u << 5,         /* feature simd_render: AA pattern + cloud sharing */
    RT_FLAG_COUNT_WORK = 1u << 6,               /* run the counter build (untimed; fills rt_stats counts) */
    RT_FLAG_ZERO_MISSES = 1u << 7,              /* write 0 for miss pixels instead of leaving them */
    RT_FLAG_PER_RAY_SHADOWS = 1u << 8           /* every shadow query walks the BVH itself (no shadow beams):
                                                   same image bit for bit; the counter build of this mode gives
                                                   the canonical per-ray test counts of the roofline */
};

typedef struct rt_config {
    uint32_t width;                 /* WINDOW_WIDTH  (src/lib.rs:30-60) */
    uint32_t height;                /* WINDOW_HEIGHT */
    uint32_t flags;                 /* RT_FLAG_* */
    uint32_t aa_samples_per_pixel;  /* ANTIALIASING_SAMPLES_PER_PIXEL: 9 or 24 (raytracer_renderer.rs:89-93) */
    const float *aa_table;          /* ANTIALIASING_SAMPLES: aa_table_len x [f32;2] (raytracer_renderer.rs:105-127) */
    uint32_t aa_table_len;          /* next_multiple_of(aa_samples_per_pixel, 8) */
    uint32_t light_multiplicator;   /* POINT_LIGHT_MULTIPLICATOR M: 1/10/19/28 (raytracer_renderer.rs:75-87) */
    const float *light_clouds;      /* n_light_clouds x M x [f32;3]: first M points of a Poisson3D
                                       process in [0,R)^3 (light.rs:183-225); NULL when M == 1 */
    uint32_t n_light_clouds;        /* K; one cloud is chosen per (pixel, light) by rt_cloud_index() */
    uint32_t reflection_max_depth;  /* RAYTRACE_REFLECTION_MAX_DEPTH 9/13/21 */
    uint32_t refraction_max_depth;  /* RAYTRACE_REFRACTION_MAX_DEPTH 8/18/21 */
    uint64_t seed;                  /* keys rt_cloud_index() */
    /* interleaved row tiles for multi-GPU: this call renders the image rows r with
       (r / tile_rows) % tile_world == tile_rank.  tile_world <= 1 renders every row. */
    uint32_t tile_rows;
    uint32_t tile_rank;
    uint32_t tile_world;
    /* Caller-managed change counter of the CONTENTS of aa_table / light_clouds. 0 = the tables are
       uploaded on every call. Non-zero = if the previous render call on this rt_scene had the same
       epoch, pointers and sizes, the device copy of light_clouds is reused (no host-to-device copy). */
    uint32_t tables_epoch;
} rt_config;

/* Work counters (RT_FLAG_COUNT_WORK) and timings. A "ray" is one closest-hit query
 * (Raytracer::cast_ray) or one shadow/transmittance query (Raytracer::has_any_intersection). */
typedef struct rt_stats {
    double kernel_ms;        /* device time of the render + resolve kernels (CUDA events) */
    double total_ms;         /* wall time of the call incl. copies */
    uint64_t closest_rays;
    uint64_t shadow_rays;
    uint64_t box_tests;      /* ray vs child-box tests */
    uint64_t triangle_tests;
    uint64_t sphere_tests;
    uint64_t kernel_launches; /* kernels launched by this call */
    uint64_t beam_triangle_tests; /* shadow-beam vs triangle culling tests (one per beam, not per ray) */
    uint64_t prepass_box_tests;      /* share of box/triangle/sphere_tests done by the primary-ray pre-pass */
    uint64_t prepass_triangle_tests; /* kernel (they are included in the totals above) */
    uint64_t prepass_sphere_tests;
} rt_stats;

typedef struct rt_scene rt_scene; /* opaque: device-resident SoA geometry + BVH */

/* ---- entry points -------------------------------------------------------------------------- */

int rt_api_version(void);
const char *rt_last_error(void);

/* Number of CUDA devices visible (0 without a GPU / driver). Never fails. */
int rt_device_count(void);

/* Measures the FP32 FMA throughput of `device` (TFLOP/s, 2 flop per FMA) with a register-resident
 * FFMA microbenchmark; the denominator of the FP32 roofline bench.py reports. */
int rt_measure_fp32_peak(int device, double *out_tflops);

/* Builds the BVH on the host, packs geometry into the device layout and uploads it to CUDA
 * device `device`.  Replaces nothing in the reference (it has no acceleration structure);
 * it is the flattening step the shim runs once per `&Scene<Vec3>`. */
int rt_scene_create(const rt_scene_desc *desc, int device, rt_scene **out_scene);
void rt_scene_destroy(rt_scene *scene);

/* What rt_scene_create built: sizes of the flattened arrays and where its time went (measurement only). */
typedef struct rt_scene_info {
    uint32_t n_nodes;        /* BVH4 nodes (128 B each) */
    uint32_t n_slots;        /* primitive slots = references (64 B each; >= primitives: large triangles are split) */
    uint32_t bvh_depth;      /* deepest BVH4 level */
    uint32_t builder;        /* 0 = host (binned SAH, multi-threaded), 1 = device (LBVH) */
    double build_ms;         /* BVH build + flattening */
    double upload_ms;        /* host-to-device copies (0 for the device builder: built in place) */
    uint64_t device_bytes;   /* nodes + slots + normals + materials + lights */
    uint64_t scratch_bytes;  /* per-scene render scratch currently allocated: samples, primary hits, chained records
                                (bounded window), path queue of the bounce configs, deferral scratch */
} rt_scene_info;
int rt_scene_get_info(const rt_scene *scene, rt_scene_info *out);

/* Threading / streams: an rt_scene owns scratch buffers and cached sampler tables that every render call on it
 * uses, so it is NOT re-entrant: call into one rt_scene from one host thread at a time (the reference renders from
 * a single render thread, src/main.rs:330-340).  Calls may use different CUDA streams (rt_render uses the scene's
 * own stream, rt_render_device the caller's): the library orders a call after the previous call on the same scene
 * with an event, so mixing them needs no device synchronisation by the caller. */

/* Replaces RaytracerRenderer::render (raytracer_renderer.rs:1369-1378).  `out_argb` is the
 * ImageBuffer storage: width*height u32, row-major, 0xFFRRGGBB for hit pixels, miss pixels
 * untouched (renderer/mod.rs:116-125) unless RT_FLAG_ZERO_MISSES.  `out_hit_mask` (optional,
 * width*height bytes) receives 1 where a pixel was written.  With row tiles only this rank's
 * rows are touched.  Host buffers; copies are inside the call. */
int rt_render(rt_scene *scene, const rt_config *cfg, uint32_t *out_argb, uint8_t *out_hit_mask,
              rt_stats *stats);

/* Progressive variant of rt_render: the frame is rendered in horizontal bands of `band_rows` rows, top to
 * bottom; after each band has been merged into `out_argb` / `out_hit_mask`, `on_band(user, first_row, n_rows)`
 * is called (may be NULL) -- the reference's window shows the image filling in while the render thread works
 * (src/output/window.rs:73-84, src/main.rs:342-347).  A non-zero return value of the callback stops the render
 * (rt_render_progressive then returns RT_OK; the remaining rows are untouched).  `cfg` must not use row tiles.
 * The image is exactly rt_render's.  `stats` (optional) sums the bands. */
/* Pipeline: while band k's pixels travel device-to-host on a copy stream and the host merges them and runs the
 * callback, band k+1 is already rendering (two device slabs, two pinned buffers), so the callback's work hides
 * behind the GPU like the reference's window loop hides behind its render thread.  The callback runs on the
 * calling thread.  rt_progressive_overlapped_bands: in the last progressive call on this scene, the number of
 * bands whose successor was still rendering when their merge began (measurement only). */
typedef int (*rt_band_callback)(void *user, uint32_t first_row, uint32_t n_rows);
int rt_render_progressive(rt_scene *scene, const rt_config *cfg, uint32_t *out_argb, uint8_t *out_hit_mask,
                          uint32_t band_rows, rt_band_callback on_band, void *user, rt_stats *stats);

uint32_t rt_progressive_overlapped_bands(const rt_scene *scene);

/* Same, but the outputs are DEVICE pointers and the work is enqueued on `cuda_stream`
 * (a cudaStream_t; NULL is the legacy default stream).  With row tiles the output is the rank's COMPACT slab:
 * local row j holds image row rt_tile_row(cfg, j).  The call returns after enqueueing unless
 * `stats` is non-NULL (then it synchronises to fill it). */
int rt_render_device(rt_scene *scene, const rt_config *cfg, void *d_out_argb, void *d_out_hit_mask,
                     void *cuda_stream, rt_stats *stats);

/* Validity of the frames enqueued with rt_render_device(stats == NULL): waits for the last one and returns
 * RT_ERR_CUDA if a bounded wait loop of the direct-lighting kernel gave up in ANY frame since the last check
 * (the device-side flag is sticky).  rt_render and rt_render_device(stats != NULL) check before they return; the
 * next render call on the scene also reports a failed earlier frame.  Call it before consuming asynchronous
 * output (e.g. before a gather). */
int rt_check_frame(rt_scene *scene);

/* Name of the dominant kernel (direct_kernel<...> / render_kernel<...>) the last render call launched. */
const char *rt_last_kernel(const rt_scene *scene);

/* Device durations (ms, CUDA events on the launching stream) of the DOMINANT kernel -- direct_kernel
 * (no bounces) or render_kernel -- of the last min(n, 64) render calls on this scene, oldest first;
 * *out_n receives how many were written.  Synchronises on the newest one.  Measurement only. */
int rt_kernel_times(rt_scene *scene, uint32_t n, float *out_ms, uint32_t *out_n);

/* Row-tile bookkeeping shared by every rank. */
uint32_t rt_tile_local_rows(const rt_config *cfg);              /* rows this rank renders */
uint32_t rt_tile_row(const rt_config *cfg, uint32_t local_row); /* image row of a local row */

/* Scatter one rank's compact slab (device) into the full frame (device): the step after the
 * NCCL gather.  `d_slab` holds rt_tile_local_rows rows of width u32 (or u8 when elem_size==1). */
int rt_scatter_tile_rows(const rt_config *cfg_of_rank, const void *d_slab, void *d_frame,
                         uint32_t elem_size, void *cuda_stream);

/* The same for ALL ranks in one launch: `d_slabs` holds the tile_world gathered slabs, slab r starting
 * at pixel r * slab_stride_pixels; writes the full frame and (optional) the hit mask, which is
 * `pixel != 0` (hit pixels carry alpha 0xFF, miss pixels of a slab are 0). */
int rt_place_tiles(const rt_config *cfg, const void *d_slabs, uint64_t slab_stride_pixels, void *d_frame,
                   void *d_hit_mask, void *cuda_stream);

/* ---- several GPUs, one process (what the Rust shim uses when gpu_count > 1) --------------------------------
 * The reference partitions the frame inside render() (src/image_buffer.rs:48-97: shuffled chunks on the rayon
 * pool).  rt_multi_* does the same behind the boundary: the BVH is built once and replicated, the frame is split
 * into interleaved row tiles (tile t -> device t % N; rt_config.tile_rows, 0 = 2 rows), every device is driven by
 * its own host thread, and each device's last kernel stores its pixels straight into the frame on the first
 * device through NVLink peer memory, from where one copy brings the frame to `out_argb`.  No torch, no second
 * process; `devices` = CUDA device indices (NULL: 0..n-1; n_devices 0: every visible device).  Same image as
 * rt_render on one device, bit for bit (sampler streams are keyed by pixel).  cfg must not set tile_world. */
typedef struct rt_multi rt_multi;
int rt_multi_create(const rt_scene_desc *desc, const int *devices, uint32_t n_devices, rt_multi **out);
void rt_multi_destroy(rt_multi *m);
uint32_t rt_multi_device_count(const rt_multi *m);
/* rt_stats.kernel_ms = the slowest device's kernels; counters are summed over the devices. */
int rt_multi_render(rt_multi *m, const rt_config *cfg, uint32_t *out_argb, uint8_t *out_hit_mask, rt_stats *stats);
/* out: the scene as built (build_ms = the one host BVH build, upload_ms = all replicas, in parallel);
 * out_peer_devices (optional): devices that reach the first device's memory over NVLink / PCIe peer access. */
int rt_multi_get_info(const rt_multi *m, rt_scene_info *out, uint32_t *out_peer_devices);

/* Debug/parity entry points: raw queries against the device BVH for arbitrary rays
 * (host arrays).  rays: n x 6 floats (origin, un-normalised direction).
 * closest: out_t[n], out_prim[n] (reference primitive index, 0xFFFFFFFF on miss).
 * shadow:  max_dist[n] in; out[n x 5] = {occluded, opacity, filter r,g,b}. */
int rt_trace_closest(rt_scene *scene, uint32_t flags, const float *rays, uint32_t n, float *out_t,
                     uint32_t *out_prim);
int rt_trace_shadow(rt_scene *scene, uint32_t flags, const float *rays, const float *max_dist,
                    uint32_t n, float *out);

/* Debug/parity (host only, no device needed): the per-frame constants and anti-aliasing sample offsets the kernels
 * receive for `cfg`, evaluated in f32 exactly as for a render call.
 * out_constants[8] = WINDOW_TO_SCENE_{WIDTH,HEIGHT,DEPTH}_FACTOR, RENDER_RAY_FOCUS.xyz, default_epsilon_distance
 * component, 1/samples (src/lib.rs:73-92, src/vector.rs:697-700, raytracer_renderer.rs:936-937).
 * out_offsets: *out_samples x [dx, dy] added to the pixel's scene coordinate for sample i
 * (raytracer_renderer.rs:876-916,1021-1138: bias direction i % 8 in packet mode, always the first one otherwise);
 * room for 24 pairs.  Either output may be NULL. */
int rt_frame_constants(const rt_config *cfg, float *out_constants, float *out_offsets, uint32_t *out_samples);

/* Debug/parity: builds the BVH of `desc` on the host only (no device needed) and returns an FNV-1a digest
 * of the flattened node / primitive / normal arrays, the BVH4 node count and depth.  The build is
 * multi-threaded (all host threads; RT_BVH_THREADS overrides) and the result must not depend on the
 * thread count. */
int rt_bvh_digest(const rt_scene_desc *desc, uint64_t *out_digest, uint32_t *out_nodes, uint32_t *out_depth);

/* ---- samplers (stand-ins for fast_poisson 1.0.2, seeded; host code) -------------------------- */

/* ANTIALIASING_SAMPLES table: [0,0], 8 x [1,1], then (randomness != 0) the first table_len-1
 * points of a Bridson Poisson-disk process on [0,1.2)^2 with radius 3/table_len, else [1,1];
 * truncated to table_len (raytracer_renderer.rs:105-127). */
int rt_sampler_aa_table(uint64_t seed, uint32_t samples_per_pixel, int randomness, float *out_table,
                        uint32_t table_len);
/* n_clouds clouds, each the first M points of a Bridson Poisson-disk process on [0,R)^3,
 * R = 1.725 + M/20, radius 4/M, M attempts per active point (light.rs:197-206). */
int rt_sampler_light_clouds(uint64_t seed, uint32_t m, uint32_t n_clouds, float *out_points);
/* Which cloud a (pixel, light) pair uses; identical in the oracle. */
uint32_t rt_cloud_index(uint64_t seed, uint32_t x, uint32_t y, uint32_t light, uint32_t n_clouds);

#ifdef __cplusplus
}
#endif
#endif /* RT_API_H */
