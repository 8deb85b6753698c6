"""Host-side mirror of the reference's scene API (csrc/host): inventories of the entry-point
scenes, BoundedPlane expansion, OBJ ingestion, feature resolution, seeded samplers."""
import ctypes as C
import math
from pathlib import Path

import numpy as np
import pytest

REF_OBJ = Path("/root/reference/data/obj/text")


def test_feature_resolution(pkg):
    L = pkg._lib
    rc = pkg.RenderConfig.from_features("", 0)
    assert (rc.cfg.width, rc.cfg.height, rc.cfg.flags, rc.cfg.light_multiplicator) == (768, 640, 0, 1)
    assert (rc.cfg.reflection_max_depth, rc.cfg.refraction_max_depth) == (9, 8)
    rc = pkg.RenderConfig.from_features("default", 0)  # Cargo.toml:64
    assert (rc.cfg.width, rc.cfg.height) == (1140, 950)
    assert rc.cfg.flags == L.FLAG_ANTI_ALIASING | L.FLAG_AA_ROTATION_SCALE | L.FLAG_REFLECTIONS | L.FLAG_REFRACTIONS
    assert rc.cfg.light_multiplicator == 19 and rc.samples_per_pixel == 9
    assert (rc.cfg.reflection_max_depth, rc.cfg.refraction_max_depth) == (13, 18)
    assert rc.high_quality_model and rc.scene_backface_culling
    rc = pkg.RenderConfig.from_features("extreme_quality,realistic,high_resolution,simd_render", 0)
    assert (rc.cfg.width, rc.cfg.height) == (1620, 1350)
    assert rc.cfg.light_multiplicator == 28 and rc.samples_per_pixel == 24
    assert (rc.cfg.reflection_max_depth, rc.cfg.refraction_max_depth) == (21, 21)
    rc = pkg.RenderConfig.from_features("soft_shadows", 0)
    assert rc.cfg.light_multiplicator == 10
    with pytest.raises(pkg.RtError):
        pkg.RenderConfig.from_features("no_such_feature", 0)


def test_entry_point_scene_inventories(pkg):
    # 1 637 + 48 triangles + 9 spheres (lowres, no pre-filter ... 1639 faces in the file)
    f = pkg.example_scene("main", "").flatten()
    assert (len(f.triangles), len(f.spheres), len(f.lights)) == (1639 + 48, 9, 5)
    f = pkg.example_scene("semesterbild", "high_quality").flatten()
    assert (len(f.triangles), len(f.spheres), len(f.lights)) == (14521 + 48, 9, 5)
    f = pkg.example_scene("test_scene", "").flatten()
    assert (len(f.triangles), len(f.spheres), len(f.lights)) == (3 + 7 * 12, 4, 6)
    f = pkg.example_scene("test_text", "").flatten()
    assert (len(f.triangles), len(f.spheres), len(f.lights)) == (14521, 0, 2)
    # OBJ triangles carry the text.mtl material (Kd, illum 2 -> metallic 0, shininess Ps)
    m = f.materials[f.triangles["material"][0]]
    assert np.allclose(m["color"], [0.123533, 0.719221, 0.826410]) and m["metallic"] == 0
    assert m["shininess"] == np.float32(0.592264) and m["has_opacity"] == 0 and m["ior"] == 1.0


def test_bounded_plane_expansion(pkg):
    s = pkg.Scene()
    m = pkg.material([0.5, 0.5, 0.5])
    s.add_bounded_plane([0, 0, -1], [0.5, 0.5, 1.0], [0, 1, 0], 2.0, 1.0, 0.1, m)
    f = s.flatten()
    assert len(f.triangles) == 12  # front/back + 4 side plates, 2 triangles each
    n = f.triangles["normal"]
    assert sorted(map(tuple, np.unique(n, axis=0).tolist())) == sorted([(0, 0, -1), (0, 0, 1), (0, 1, 0), (0, -1, 0), (1, 0, 0), (-1, 0, 0)])
    v = np.concatenate([f.triangles["v1"], f.triangles["v2"], f.triangles["v3"]])
    assert np.allclose(v.min(0), [-0.5, 0.0, 0.95]) and np.allclose(v.max(0), [1.5, 1.0, 1.05])
    with pytest.raises(pkg.RtError):  # the reference asserts: up must be orthogonal to normal
        s.add_bounded_plane([0, 0, -1], [0, 0, 0], [0, 0.5, 0.5], 1.0, 1.0, 0.1, m)
    with pytest.raises(pkg.RtError):
        s.add_bounded_plane([0, 0, -1], [0, 0, 0], [0, 1, 0], 0.0, 1.0, 0.1, m)


def test_scene_backface_prefilter(pkg):
    s = pkg.Scene()
    opaque, glass = pkg.material([1, 1, 1]), pkg.material([1, 1, 1], opacity=0.5, ior=1.5)
    s.add_triangle([0, 0, 0], [1, 0, 0], [0, 1, 0], opaque)  # normal +z: dropped
    s.add_triangle([0, 0, 0], [0, 1, 0], [1, 0, 0], opaque)  # normal -z: kept
    s.add_triangle([0, 0, 0], [1, 0, 0], [0, 1, 0], glass)   # transmissive: kept
    s.add_sphere([0, 0, 0], 1.0, opaque)
    s.backface_culling([0, 0, 1])
    f = s.flatten()
    assert len(f.triangles) == 2 and len(f.spheres) == 1


def test_light_color_goes_through_maximize_value(pkg):
    s = pkg.Scene()
    s.add_light([0, 0, 0], [1, 1, 1], 1.0)
    s.add_light([0, 0, 0], [0.825, 0.675, 0.5], 1.0)
    s.add_light([0, 0, 0], [0.0, 0.25, 0.125], 0.5)
    c = s.flatten().lights["color"]
    assert np.allclose(c[0], 1.0)
    assert c[1].max() == pytest.approx(1.0, abs=1e-6) and c[1][0] > c[1][1] > c[1][2] > 0.3  # HSV value -> 1, hue/saturation kept
    assert c[2].max() == pytest.approx(1.0, abs=1e-6) and c[2][0] == 0.0


@pytest.mark.skipif(not REF_OBJ.exists(), reason="reference checkout not mounted")
@pytest.mark.parametrize("name", ["text", "text_lowres"])
def test_obj_loader_matches_committed_mesh_fixture(pkg, name):
    tf = dict(translation=(0.0135, 0.12, 0.8), euler_rpy=(0.0, -0.015, 0.0), scale=1.1)
    a = pkg.Scene()
    a.add_obj(REF_OBJ / f"{name}.obj", **tf)
    b = pkg.Scene()
    mesh = pkg.Mesh.load_rtmesh(pkg.scene.MESH_DIR / f"{name}.rtmesh")
    b.add_mesh(mesh, pkg.material([0.123533, 0.719221, 0.826410], 0.0, 0.592264), **tf)
    fa, fb = a.flatten(), b.flatten()
    assert fa.triangles.tobytes() == fb.triangles.tobytes()
    assert fa.materials.tobytes() == fb.materials.tobytes()


def test_obj_loader_errors(pkg, tmp_path):
    s = pkg.Scene()
    with pytest.raises(pkg.RtError) as e:
        s.add_obj(tmp_path / "missing.obj")
    assert e.value.code == pkg._lib.RT_ERR_IO
    p = tmp_path / "quad.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\n")  # polygon -> fan of 2; no vn -> geometric normal, white diffuse
    s.add_obj(p)
    f = s.flatten()
    assert len(f.triangles) == 2 and np.allclose(f.triangles["normal"], [0, 0, 1])
    assert np.allclose(f.materials[0]["color"], 1.0)


def test_seeded_samplers(lib, pkg):
    # AA table (raytracer_renderer.rs:105-127)
    t = np.zeros((16, 2), np.float32)
    assert lib.rt_sampler_aa_table(0, 9, 0, pkg._lib.ptr(t), 16) == 0
    assert t[0].tolist() == [0, 0] and (t[1:] == 1).all()
    assert lib.rt_sampler_aa_table(42, 9, 1, pkg._lib.ptr(t), 16) == 0
    assert t[0].tolist() == [0, 0] and (t[1:9] == 1).all()
    pts = t[9:]
    assert ((pts >= 0) & (pts < 1.2)).all()
    d = np.linalg.norm(pts[:, None] - pts[None], axis=-1) + np.eye(len(pts)) * 9
    assert d.min() >= 3.0 / 16 - 1e-6  # Poisson-disk radius
    t2 = np.zeros((16, 2), np.float32)
    lib.rt_sampler_aa_table(42, 9, 1, pkg._lib.ptr(t2), 16)
    assert np.array_equal(t, t2)  # deterministic in the seed
    lib.rt_sampler_aa_table(43, 9, 1, pkg._lib.ptr(t2), 16)
    assert not np.array_equal(t, t2)
    # soft-shadow clouds (light.rs:197-206): M points, min distance 4/M, inside [0, 1.725 + M/20)^3
    for M in (10, 19, 28):
        c = np.zeros((64, M, 3), np.float32)
        assert lib.rt_sampler_light_clouds(7, M, 64, pkg._lib.ptr(c)) == 0
        R = 1.725 + M / 20
        assert ((c >= 0) & (c < R)).all()
        for cl in c:
            d = np.linalg.norm(cl[:, None] - cl[None], axis=-1) + np.eye(M) * 9
            assert d.min() >= 4.0 / M - 1e-5
        assert len(np.unique(c.reshape(64, -1), axis=0)) == 64
    assert lib.rt_sampler_light_clouds(7, 0, 64, pkg._lib.ptr(c)) == pkg._lib.RT_ERR_INVALID_ARGUMENT


def test_camera_projection_of_scene_corners(pkg):
    """the plane corners of src/main.rs projected through the pinhole land on the
    silhouette of the reference render (rotor / isometry conventions)."""
    f = pkg.example_scene("main", "medium_resolution").flatten()
    W, H = 1140, 950
    sh = np.float32(H) / np.float32(W)
    sd = (1 + sh) / 2
    F = np.array([0.5, sh / 2, -1.9 * sd])
    def project(q):
        s = -F[2] / (q[2] - F[2])
        return (F[0] + s * (q[0] - F[0])) * W, (F[1] + s * (q[1] - F[1])) * H / sh
    back = f.triangles[-48:-36]  # first BoundedPlane = back wall
    v = np.concatenate([back["v1"], back["v2"], back["v3"]])
    px = np.array([project(q) for q in v])
    # back-wall top-left corner ~ (35, 68) and top-right ~ (688, 103) in output.png
    assert np.abs(px - np.array([34.5, 67])).sum(1).min() < 6
    assert np.abs(px - np.array([687.5, 104])).sum(1).min() < 6


def test_parallel_bvh_build_is_independent_of_the_thread_count(pkg, lib, monkeypatch):
    """rt_scene_create builds the BVH with all host threads (subtrees set aside as tasks, spliced back in);
    the flattened node / primitive arrays must be those of the sequential build, bit for bit."""
    import ctypes as C

    from hslu_i.ba_raytracing.f2501_raytracer_b200.synthetic import large_mesh_scene

    def digest(flat, threads):
        monkeypatch.setenv("RT_BVH_THREADS", str(threads))
        d = flat.desc()
        h, n, depth = C.c_uint64(), C.c_uint32(), C.c_uint32()
        pkg._lib.check(lib.rt_bvh_digest(C.byref(d), C.byref(h), C.byref(n), C.byref(depth)))
        return h.value, n.value, depth.value

    for flat in (pkg.example_scene("semesterbild", "high_quality").flatten(), large_mesh_scene("high_quality", nx=256, ny=256).flatten()):
        ref = digest(flat, 1)
        assert ref[1] > 0 and 0 < ref[2] <= 20
        assert digest(flat, 2) == ref and digest(flat, 7) == ref


def test_file_output_writes_the_framebuffer_as_rgb8_png(pkg, tmp_path):
    """FileOutput::render_buffer (reference src/output/file.rs:27-49): RGB8 PNG of the u32 buffer, untouched pixels black."""
    PIL = pytest.importorskip("PIL.Image")
    buf = pkg.ImageBuffer(7, 5)
    rng = np.random.default_rng(3)
    buf.buffer[:] = rng.integers(0, 1 << 24, 35, dtype=np.uint32) | np.uint32(0xFF000000)
    buf.buffer[::4] = 0  # untouched
    out = tmp_path / "frame.png"
    pkg.FileOutput(out).render_buffer(buf)
    img = PIL.open(out)
    assert img.mode == "RGB" and img.size == (7, 5)
    assert np.array_equal(np.asarray(img), buf.rgb())


def test_feature_string_and_window_title(pkg, tmp_path):
    """Output::get_feature_string (reference src/output/mod.rs:25-88) and the window title built from it
    (src/output/window.rs:58-64): expected strings written out by hand from the format! call."""
    X = "×"
    cases = {
        "": f"Non-SIMD | Non-Antialiasing | Non-Realistic | Standard Quality | Small Resolution (768{X}640{X}704) | NO-OPT",
        "default,simd_render": f"SIMD | Antialiasing ROS_SCL RNG | Reflections + Refractions | High Quality | Medium Resolution (1140{X}950{X}1045) | NO-OPT",
        "simd_render,backface_culling,medium_resolution": f"SIMD | Non-Antialiasing | Non-Realistic | Standard Quality | Medium Resolution (1140{X}950{X}1045) | Backface Culling",
        "high_quality": f"Non-SIMD | Antialiasing | Non-Realistic | High Quality | Small Resolution (768{X}640{X}704) | NO-OPT",
        "extreme_quality,realistic,high_resolution,simd_render": f"SIMD | Antialiasing | Reflections + Refractions | Extreme Quality | High Resolution (1620{X}1350{X}1485) | NO-OPT",
        "reflections": f"Non-SIMD | Non-Antialiasing | Reflections | Standard Quality | Small Resolution (768{X}640{X}704) | NO-OPT",
        # format!("Antialiasing {} {}", "", "RNG").trim() keeps the inner double blank; "ROS_SCL" alone loses its trailing one
        "anti_aliasing_randomness": f"Non-SIMD | Antialiasing  RNG | Non-Realistic | Standard Quality | Small Resolution (768{X}640{X}704) | NO-OPT",
        "anti_aliasing_rotation_scale": f"Non-SIMD | Antialiasing ROS_SCL | Non-Realistic | Standard Quality | Small Resolution (768{X}640{X}704) | NO-OPT",
    }
    for feat, want in cases.items():
        assert pkg.get_feature_string(feat) == want
        assert pkg.window_title(feat) == "Raytracer - " + want
    with pytest.raises(pkg.RtError):
        pkg.get_feature_string("no_such_feature")
    # FileOutput(path, features=...): the title travels in a tEXt chunk; the pixels are the same file without it
    buf = pkg.ImageBuffer(16, 8)
    buf.buffer[::3] = 0xFF336699
    a, b = tmp_path / "plain.png", tmp_path / "titled.png"
    pkg.FileOutput(a).render_buffer(buf)
    pkg.FileOutput(b, features="default,simd_render").render_buffer(buf)
    raw_a, raw_b = a.read_bytes(), b.read_bytes()
    title = ("Raytracer - " + cases["default,simd_render"]).encode("latin-1")
    assert b"tEXt" not in raw_a and b"tEXtTitle\x00" + title in raw_b
    i = raw_b.index(b"tEXt") - 4
    n = int.from_bytes(raw_b[i:i + 4], "big")
    assert raw_b[:i] + raw_b[i + 12 + n:] == raw_a
