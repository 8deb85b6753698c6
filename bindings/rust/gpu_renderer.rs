//! `GpuRaytracerRenderer`: drop-in `impl Renderer` that sends the render loop to librt_b200.so.
//!
//! NOT BUILT in this repository (the image has no Rust toolchain); it documents the reference-side
//! binding a maintainer adds under `src/renderer/gpu_renderer.rs` of
//! kije/HSLU_I.BA_RAYTRACING.F2501_raytracer and is kept in sync with `include/rt_api.h`.
//! It must live inside the crate because `TriangleData::edge1/edge2` and `SphereData::r_sq` are
//! private (src/geometry/basic/triangle.rs:39-40, sphere.rs:24); it only needs the public vertices,
//! the stored normal, the material and one new accessor `SphereData::radius()`.
use crate::geometry::RenderGeometry;
use crate::image_buffer::ImageBuffer;
use crate::output::OutputColorEncoder;
use crate::renderer::Renderer;
use crate::scene::{Scene, SceneLightSource};
use std::ffi::CStr;
use std::marker::PhantomData;
use std::os::raw::{c_char, c_int, c_void};
use ultraviolet::Vec3;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct RtMaterial { color: [f32; 3], metallic: f32, shininess: f32, ior: f32, opacity: f32, has_opacity: u32, boost: f32 }
#[repr(C)]
pub struct RtTriangle { v1: [f32; 3], v2: [f32; 3], v3: [f32; 3], normal: [f32; 3], material: u32 }
#[repr(C)]
pub struct RtSphere { center: [f32; 3], radius: f32, material: u32 }
#[repr(C)]
pub struct RtLight { position: [f32; 3], color: [f32; 3], intensity: f32 }
#[repr(C)]
pub struct RtSceneDesc {
    triangles: *const RtTriangle, n_triangles: u32,
    spheres: *const RtSphere, n_spheres: u32,
    materials: *const RtMaterial, n_materials: u32,
    lights: *const RtLight, n_lights: u32,
}
#[repr(C)]
pub struct RtConfig {
    width: u32, height: u32, flags: u32, aa_samples_per_pixel: u32,
    aa_table: *const f32, aa_table_len: u32, light_multiplicator: u32,
    light_clouds: *const f32, n_light_clouds: u32, reflection_max_depth: u32, refraction_max_depth: u32,
    seed: u64, tile_rows: u32, tile_rank: u32, tile_world: u32, tables_epoch: u32,
}
#[repr(C)]
pub struct RtScene { _private: [u8; 0] }
#[repr(C)]
pub struct RtMulti { _private: [u8; 0] }

const RT_FLAG_ANTI_ALIASING: u32 = 1 << 0;
const RT_FLAG_AA_ROTATION_SCALE: u32 = 1 << 1;
const RT_FLAG_REFLECTIONS: u32 = 1 << 2;
const RT_FLAG_REFRACTIONS: u32 = 1 << 3;
const RT_FLAG_BACKFACE_CULLING: u32 = 1 << 4;
const RT_FLAG_PACKET_SEMANTICS: u32 = 1 << 5;

#[link(name = "rt_b200")]
extern "C" {
    fn rt_last_error() -> *const c_char;
    fn rt_scene_create(desc: *const RtSceneDesc, device: c_int, out: *mut *mut RtScene) -> c_int;
    fn rt_scene_destroy(scene: *mut RtScene);
    fn rt_render(scene: *mut RtScene, cfg: *const RtConfig, out_argb: *mut u32, out_hit_mask: *mut u8, stats: *mut c_void) -> c_int;
    // several GPUs, one process: the frame is partitioned inside the call, like ImageBuffer::process_chunks_parallel
    // (src/image_buffer.rs:48-97) partitions it over the rayon pool
    fn rt_device_count() -> c_int;
    fn rt_multi_create(desc: *const RtSceneDesc, devices: *const c_int, n_devices: u32, out: *mut *mut RtMulti) -> c_int;
    fn rt_multi_destroy(m: *mut RtMulti);
    fn rt_multi_render(m: *mut RtMulti, cfg: *const RtConfig, out_argb: *mut u32, out_hit_mask: *mut u8, stats: *mut c_void) -> c_int;
    fn rt_sampler_aa_table(seed: u64, samples_per_pixel: u32, randomness: c_int, out: *mut f32, table_len: u32) -> c_int;
    fn rt_sampler_light_clouds(seed: u64, m: u32, n_clouds: u32, out: *mut f32) -> c_int;
}

fn check(code: c_int) {
    if code != 0 {
        // the reference's render() returns () and panics on failure (src/main.rs:46,343)
        let msg = unsafe { CStr::from_ptr(rt_last_error()) }.to_string_lossy().into_owned();
        panic!("librt_b200: error {code}: {msg}");
    }
}

#[derive(Debug, Copy, Clone, Default)]
pub struct GpuRaytracerRenderer<C: OutputColorEncoder>(PhantomData<C>);

impl<const W: usize, const H: usize, C: OutputColorEncoder> Renderer<W, H, C> for GpuRaytracerRenderer<C>
where
    [(); W * H]:,
{
    fn render(&self, buffer: &ImageBuffer<W, H>, scene: &Scene<Vec3>) {
        // ---- flatten Scene<Vec3>; kind order fixed to spheres, then triangles (rt_api.h)
        let mut materials: Vec<RtMaterial> = Vec::new();
        let mut mat_index = |m: &crate::raytracing::Material<f32>| -> u32 {
            let t = &m.transmission;
            let flat = RtMaterial {
                color: [m.color.red, m.color.green, m.color.blue],
                metallic: m.metallic,
                shininess: m.shininess(),
                ior: t.raw_refraction_index(),      // new pub accessor: the stored field
                opacity: t.raw_opacity(),           // new pub accessor: *self.opacity.value()
                has_opacity: t.raw_has_opacity() as u32,
                boost: t.raw_boost(),
            };
            materials.push(flat);
            (materials.len() - 1) as u32
        };
        let (mut tris, mut spheres) = (Vec::new(), Vec::new());
        for g in scene.scene_objects.get_all() {
            if let RenderGeometry::Sphere(s) = g {
                spheres.push(RtSphere { center: *s.center.as_array(), radius: s.radius(), material: mat_index(&s.material) });
            }
        }
        for g in scene.scene_objects.get_all() {
            if let RenderGeometry::Triangle(t) = g {
                tris.push(RtTriangle {
                    v1: *t.vertex1.as_array(), v2: *t.vertex2.as_array(), v3: *t.vertex3.as_array(),
                    normal: *t.normal.as_array(), material: mat_index(&t.material),
                });
            }
        }
        let lights: Vec<RtLight> = scene.scene_lights.iter().map(|l| match l {
            SceneLightSource::PointLight(p) => RtLight {
                position: *p.position.as_array(), color: [p.color.red, p.color.green, p.color.blue], intensity: p.intensity,
            },
        }).collect();
        let desc = RtSceneDesc {
            triangles: tris.as_ptr(), n_triangles: tris.len() as u32,
            spheres: spheres.as_ptr(), n_spheres: spheres.len() as u32,
            materials: materials.as_ptr(), n_materials: materials.len() as u32,
            lights: lights.as_ptr(), n_lights: lights.len() as u32,
        };

        // ---- cfg!() feature switches -> rt_config (raytracer_renderer.rs:55-93)
        let mut flags = 0u32;
        if cfg!(feature = "anti_aliasing") { flags |= RT_FLAG_ANTI_ALIASING; }
        if cfg!(feature = "anti_aliasing_rotation_scale") { flags |= RT_FLAG_AA_ROTATION_SCALE; }
        if cfg!(feature = "reflections") { flags |= RT_FLAG_REFLECTIONS; }
        if cfg!(feature = "refractions") { flags |= RT_FLAG_REFRACTIONS; }
        if cfg!(feature = "backface_culling") { flags |= RT_FLAG_BACKFACE_CULLING; }
        if cfg!(feature = "simd_render") { flags |= RT_FLAG_PACKET_SEMANTICS; }
        let hq = cfg!(feature = "high_quality");
        let xq = cfg!(feature = "extreme_quality");
        let spp: u32 = if xq { 24 } else { 9 };
        let table_len = spp.next_multiple_of(8);
        let m: u32 = if cfg!(feature = "soft_shadows") { if hq { if xq { 28 } else { 19 } } else { 10 } } else { 1 };
        let seed: u64 = std::env::var("RT_SEED").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        // The library's seeded samplers; a shim that wants the crate's own fast_poisson points can
        // fill these two tables with `Poisson2D` / `Poisson3D::new().with_seed(..)` output instead.
        let mut aa = vec![0f32; 2 * table_len as usize];
        check(unsafe { rt_sampler_aa_table(seed, spp, cfg!(feature = "anti_aliasing_randomness") as c_int, aa.as_mut_ptr(), table_len) });
        let n_clouds = 4096u32;
        let mut clouds = vec![0f32; if m > 1 { (n_clouds * m * 3) as usize } else { 0 }];
        if m > 1 { check(unsafe { rt_sampler_light_clouds(seed, m, n_clouds, clouds.as_mut_ptr()) }); }
        let cfg = RtConfig {
            width: W as u32, height: H as u32, flags, aa_samples_per_pixel: spp,
            aa_table: aa.as_ptr(), aa_table_len: table_len, light_multiplicator: m,
            light_clouds: if m > 1 { clouds.as_ptr() } else { std::ptr::null() }, n_light_clouds: if m > 1 { n_clouds } else { 0 },
            reflection_max_depth: if hq { if xq { 21 } else { 13 } } else { 9 },
            refraction_max_depth: if hq { if xq { 21 } else { 18 } } else { 8 },
            seed, tile_rows: 0, tile_rank: 0, tile_world: 1, tables_epoch: 0,
        };

        // ---- render straight into the ImageBuffer: [AtomicU32; W*H] is layout-identical to [u32; W*H]
        // (#[repr(transparent)], src/image_buffer.rs:9-15); misses are left untouched like the reference.
        let out = buffer.get_u32_slice().as_ptr() as *mut u32;
        // RT_GPUS=n (default 1; 0 = every visible GPU): with more than one, rt_multi_* replicates the scene, deals
        // the image rows out in interleaved tiles and lets every GPU store its rows into the first GPU's frame
        // over NVLink; the call returns with the whole ImageBuffer filled.
        let visible = unsafe { rt_device_count() }.max(0) as u32;
        let gpus: u32 = std::env::var("RT_GPUS").ok().and_then(|s| s.parse().ok()).map(|n: u32| if n == 0 { visible } else { n.min(visible) }).unwrap_or(1);
        if gpus > 1 {
            let mut multi: *mut RtMulti = std::ptr::null_mut();
            check(unsafe { rt_multi_create(&desc, std::ptr::null(), gpus, &mut multi) });
            let rc = unsafe { rt_multi_render(multi, &cfg, out, std::ptr::null_mut(), std::ptr::null_mut()) };
            unsafe { rt_multi_destroy(multi) };
            check(rc);
            return;
        }
        let mut dev: *mut RtScene = std::ptr::null_mut();
        check(unsafe { rt_scene_create(&desc, 0, &mut dev) });
        let rc = unsafe { rt_render(dev, &cfg, out, std::ptr::null_mut(), std::ptr::null_mut()) };
        unsafe { rt_scene_destroy(dev) };
        check(rc);
    }
}
