"""Pins the CPU oracle (and the host-side scene builders feeding it) to what the reference itself
provides for this path: its committed render output.png (tests/golden/output_png_kats.json,
extracted by tests/golden/make_output_png_kats.py) plus hand-computable known answers for every
formula on the path (the reference ships no tests for any of them)."""
import ctypes as C
import json
import math
from pathlib import Path

import numpy as np
import pytest

GOLD = json.loads((Path(__file__).parent / "golden" / "output_png_kats.json").read_text())


def rgb_of(argb):
    return [(int(argb) >> 16) & 255, (int(argb) >> 8) & 255, int(argb) & 255]


@pytest.fixture(scope="module")
def main_scene_default(pkg):
    # crate defaults: text.obj, scene_backface_culling (Cargo.toml:64); geometry does not depend on the quality flags
    return pkg.example_scene("main", "default,simd_render").flatten()


def test_default_scene_inventory(main_scene_default):
    # 14 520 text triangles after the +z pre-filter + 48 plane triangles, 9 spheres, 5 lights
    assert len(main_scene_default.triangles) == 14568
    assert len(main_scene_default.spheres) == 9
    assert len(main_scene_default.lights) == 5


def test_smooth_pixels_reproduce_reference_render(pkg, oracle_mod, main_scene_default):
    """Centre sample, M = 1, depth 13/18: must match output.png to the LSB in smooth regions."""
    feat = "realistic,scene_backface_culling,medium_resolution,high_quality_model"
    rc = pkg.RenderConfig.from_features(feat, 0)
    rc.cfg.reflection_max_depth, rc.cfg.refraction_max_depth = 13, 18
    W = GOLD["width"]
    pts = GOLD["smooth"]
    pix = np.array([p["y"] * W + p["x"] for p in pts], np.uint32)
    desc = main_scene_default.desc()
    argb, mask, _, _ = oracle_mod.render(desc, rc.cfg, oracle_mod.BRUTE, 0, pix)
    exact = 0
    for p, i in zip(pts, pix):
        assert mask[i] == 1
        d = max(abs(a - b) for a, b in zip(rgb_of(argb[i]), p["rgb"]))
        assert d <= 1, (p, rgb_of(argb[i]))
        exact += d == 0
    assert exact >= 10  # the three +-1 pixels sit next to penumbrae


def test_full_default_feature_set_close_to_reference_render(pkg, oracle_mod, main_scene_default):
    """default + simd_render (16-sample packet AA with rotation/randomness, 95 soft-shadow rays,
    depth 13/18) with OUR seeded samplers: stochastic pixels land within a few LSB of output.png."""
    feat = "default,simd_render"
    rc = pkg.RenderConfig.from_features(feat, 0)
    assert (rc.cfg.width, rc.cfg.height) == (1140, 950)
    assert rc.samples_per_pixel == 16 and rc.cfg.light_multiplicator == 19
    W = GOLD["width"]
    pts = GOLD["stochastic"] + GOLD["smooth"][:4]
    pix = np.array([p["y"] * W + p["x"] for p in pts], np.uint32)
    desc = main_scene_default.desc()
    argb, mask, _, _ = oracle_mod.render(desc, rc.cfg, oracle_mod.BVH, 0, pix)
    worst = 0
    for p, i in zip(pts, pix):
        d = max(abs(a - b) for a, b in zip(rgb_of(argb[i]), p["rgb"]))
        worst = max(worst, d)
        assert d <= 12, (p, rgb_of(argb[i]))
    assert worst <= 12


def test_hit_mask_grid_matches_reference_render(pkg, oracle_mod, main_scene_default):
    """Primary-hit mask (centre sample) vs the non-black mask of output.png on a 12-px grid; the
    only disagreements allowed are AA-dilated silhouette pixels (13 of 7505)."""
    feat = "scene_backface_culling,medium_resolution,high_quality_model"
    rc = pkg.RenderConfig.from_features(feat, 0)
    g = GOLD["mask_grid"]
    ref = np.array(g["mask"], np.uint8)
    H, W = GOLD["height"], GOLD["width"]
    ys, xs = np.mgrid[g["y0"]:H:g["step"], g["x0"]:W:g["step"]]
    pix = (ys * W + xs).astype(np.uint32).ravel()
    desc = main_scene_default.desc()
    _, mask, _, _ = oracle_mod.render(desc, rc.cfg, oracle_mod.BVH, 0, pix)
    got = mask[pix].reshape(ref.shape)
    assert (got != ref).sum() <= 20
    assert not ((got == 1) & (ref == 0)).any()  # we never hit where the reference render is black
    assert abs(got.mean() - GOLD["non_black_fraction"]) < 0.005


# ---------------------------------------------------------------- formula-level known answers
def tiny_scene(pkg, with_sphere=True, with_tri=True, mat=None, cull=False):
    s = pkg.Scene()
    m = mat if mat is not None else pkg.material([0.5, 0.5, 0.5])
    if with_sphere:
        s.add_sphere([0, 0, 5], 1.0, m)
    if with_tri:
        s.add_triangle([-1, -1, 10], [1, -1, 10], [-1, 1, 10], m)
    return s.flatten()


def test_sphere_kats(pkg, oracle_mod):
    flat = tiny_scene(pkg, with_tri=False)
    d = flat.desc()
    rays = np.array([
        [0, 0, 0, 0, 0, 1],      # head-on: t0 = 4
        [0, 0, 5, 0, 0, 1],      # from the centre: t0 < 0, t1 = 1 (inside-sphere branch)
        [0, 0, 5, 0, 0, -2],     # direction is normalised by Ray::new
        [0, 1, 0, 0, 0, 1],      # tangent: disc = 0, t = 5
        [0, 1.0001, 0, 0, 0, 1], # just misses
        [0, 0, 7, 0, 0, 1],      # sphere behind the origin
        [0, 0, 4, 0, 0, 1],      # on the surface: t0 = 0 is accepted (no epsilon, sphere.rs:111)
    ], np.float32)
    t, prim = oracle_mod.trace_closest(d, rays)
    assert t[0] == 4.0 and prim[0] == 0
    assert t[1] == 1.0
    assert t[2] == 1.0
    assert t[3] == 5.0
    assert math.isinf(t[4]) and prim[4] == 0xFFFFFFFF
    assert math.isinf(t[5])
    assert t[6] == 0.0 and prim[6] == 0


def test_triangle_kats(pkg, oracle_mod):
    flat = tiny_scene(pkg, with_sphere=False)
    d = flat.desc()
    eps = np.float32(1.1920929e-7)
    rays = np.array([
        [-0.5, -0.5, 0, 0, 0, 1],   # interior: t = 10
        [0, 0, 0, 0, 0, 1],         # on the hypotenuse u + v = 1: rejected (triangle.rs:192)
        [-1, -1, 0, 0, 0, 1],       # on vertex 1: u = v = 0 accepted
        [-1.001, -0.5, 0, 0, 0, 1], # u < 0
        [-0.5, -0.5, 10, 0, 0, 1],  # origin in the plane: t = 0 <= eps rejected
        [-0.5, -0.5, 20, 0, 0, 1],  # behind
        [-0.5, -0.5, 0, 1, 0, 0],   # parallel to the plane: det = 0
    ], np.float32)
    t, prim = oracle_mod.trace_closest(d, rays)
    assert t[0] == 10.0 and prim[0] == 0
    assert math.isinf(t[1])
    assert t[2] == 10.0
    assert math.isinf(t[3]) and math.isinf(t[4]) and math.isinf(t[5]) and math.isinf(t[6])


def test_closest_hit_tie_goes_to_later_primitive(pkg, oracle_mod):
    s = pkg.Scene()
    m0, m1 = pkg.material([1, 0, 0]), pkg.material([0, 1, 0])
    s.add_triangle([-1, -1, 3], [1, -1, 3], [-1, 1, 3], m0)
    s.add_triangle([-1, -1, 3], [1, -1, 3], [-1, 1, 3], m1)  # coincident duplicate
    flat = s.flatten()
    for mode in (oracle_mod.BRUTE, oracle_mod.BVH):
        t, prim = oracle_mod.trace_closest(flat.desc(), np.array([[-0.5, -0.5, 0, 0, 0, 1]], np.float32), mode=mode)
        assert t[0] == 3.0 and prim[0] == 1  # `<=` in raytracer.rs:193-194


def test_backface_term(pkg, oracle_mod):
    """feature backface_culling: a primitive is visible iff d.n < 0.75 or it is transmissive."""
    opaque = pkg.material([0.5, 0.5, 0.5])
    glass = pkg.material([0.5, 0.5, 0.5], opacity=0.5, ior=1.5)
    ray = np.array([[-0.5, -0.5, 0, 0, 0, 1]], np.float32)
    for mat, expect_hit in ((opaque, False), (glass, True)):
        s = pkg.Scene()
        # normal (0,0,1) points along the ray: d.n = 1 >= 0.75
        s.add_triangle([-1, -1, 10], [1, -1, 10], [-1, 1, 10], mat)
        flat = s.flatten()
        assert np.allclose(flat.triangles["normal"][0], [0, 0, 1])
        t, _ = oracle_mod.trace_closest(flat.desc(), ray, flags=pkg._lib.FLAG_BACKFACE_CULLING)
        assert math.isinf(t[0]) != expect_hit
        t, _ = oracle_mod.trace_closest(flat.desc(), ray, flags=0)
        assert t[0] == 10.0


def test_shadow_query_kats(pkg, oracle_mod):
    glass = pkg.material([0.2, 0.4, 0.6], opacity=0.5, ior=1.5)
    s = pkg.Scene()
    s.add_sphere([0, 0, 5], 1.0, glass)
    flat = s.flatten()
    rays = np.array([[0, 0, 0, 0, 0, 1], [0, 0, 0, 0, 0, 1], [0, 3, 0, 0, 0, 1]], np.float32)
    out = oracle_mod.trace_shadow(flat.desc(), rays, np.array([10, 3.9, 10], np.float32))
    # through the sphere (counted ONCE: the sphere test returns only its nearest root)
    f0 = ((1 - 1.5) / (1 + 1.5)) ** 2
    io = 0.5 * (1 - f0)
    assert out[0, 0] == 0.0
    assert out[0, 1] == pytest.approx(1 - (1 - io), rel=1e-6)
    assert out[0, 2:] == pytest.approx([1 - 0.2 * 0.5, 1 - 0.4 * 0.5, 1 - 0.6 * 0.5], rel=1e-6)
    # light in front of the sphere / ray passing beside it: untouched
    for r in (1, 2):
        assert list(out[r]) == [0.0, 1.0, 1.0, 1.0, 1.0]
    # an opaque blocker occludes completely
    s2 = pkg.Scene()
    s2.add_sphere([0, 0, 5], 1.0, pkg.material([1, 1, 1]))
    out2 = oracle_mod.trace_shadow(s2.flatten().desc(), rays[:1], np.array([10], np.float32))
    assert out2[0, 0] == 1.0 and out2[0, 1] == 0.0


def test_encode_kats(oracle_mod):
    lib = oracle_mod.load()
    enc = lambda r, g, b: lib.oracle_encode(r, g, b)
    assert enc(0, 0, 0) == 0xFF000000
    assert enc(1, 1, 1) == 0xFFFFFFFF
    assert enc(2.0, -1.0, 0.5) == 0xFFFF0080  # clamp high, clamp negative to 0, 127.5 -> 128 (ties to even)
    assert enc(0.5 / 255, 1.5 / 255, 2.5 / 255) == 0xFF000202  # round half to even: 0.5->0, 1.5->2, 2.5->2
    assert enc(float("nan"), float("inf"), -float("inf")) == 0xFFFFFF00  # NaN.min(255) = 255 in palette's conversion


def test_cloud_index_matches_product(lib, oracle_mod):
    olib = oracle_mod.load()
    rng = np.random.default_rng(3)
    for _ in range(2000):
        seed, x, y, l, n = int(rng.integers(0, 2**63)), int(rng.integers(0, 4000)), int(rng.integers(0, 4000)), int(rng.integers(0, 8)), int(rng.integers(1, 9000))
        a, b = lib.rt_cloud_index(seed, x, y, l, n), olib.oracle_cloud_index(seed, x, y, l, n)
        assert a == b and a < n


def test_nan_ray_semantics(pkg, oracle_mod):
    """A refraction child under total internal reflection gets a NaN direction (zero vector
    normalised). Every triangle test then reports valid with t = NaN and the scan keeps the first
    triangle in order; spheres report no hit ('NaN rays')."""
    s = pkg.Scene()
    m = pkg.material([0.5, 0.5, 0.5])
    s.add_sphere([0, 0, 5], 1.0, m)
    s.add_triangle([-1, -1, 10], [1, -1, 10], [-1, 1, 10], m)
    s.add_triangle([-1, -1, 20], [1, -1, 20], [-1, 1, 20], m)
    flat = s.flatten()
    ray = np.array([[0, 0, 0, float("nan"), 0, 1]], np.float32)
    for mode in (oracle_mod.BRUTE, oracle_mod.BVH):
        t, prim = oracle_mod.trace_closest(flat.desc(), ray, mode=mode)
        assert math.isnan(t[0]) and prim[0] == 1  # first triangle (primitive 0 is the sphere)
        out = oracle_mod.trace_shadow(flat.desc(), ray, np.array([100], np.float32), mode=mode)
        assert list(out[0]) == [0.0, 1.0, 1.0, 1.0, 1.0]


def test_dense_grid_of_reference_render_distribution(pkg, oracle_mod, main_scene_default):
    """1748 pixels of output.png on a 25-px grid vs the oracle under the full default + simd_render feature
    set with OUR seeded samplers (the reference's are seeded from OS entropy): the hit mask must agree
    everywhere and the colours must agree up to sampler noise -- measured when the fixture was made:
    62 % exact, 88 % within 1 LSB, 96 % within 5 LSB, mean |diff| 0.65 LSB."""
    rc = pkg.RenderConfig.from_features("default,simd_render", 0)
    g = GOLD["dense_grid"]
    H, W = GOLD["height"], GOLD["width"]
    ys, xs = np.mgrid[g["y0"]:H:g["step"], g["x0"]:W:g["step"]]
    pix = (ys * W + xs).astype(np.uint32).ravel()
    ref = np.array(g["rgb"], np.int32)
    assert len(ref) == len(pix) == 1748
    argb, mask, _, _ = oracle_mod.render(main_scene_default.desc(), rc.cfg, oracle_mod.BVH, 0, pix)
    got = np.stack([(argb[pix] >> 16) & 255, (argb[pix] >> 8) & 255, argb[pix] & 255], -1).astype(np.int32)
    ref_mask = ref.sum(-1) > 0
    assert np.array_equal(ref_mask, mask[pix] == 1)
    d = np.abs(got[ref_mask] - ref[ref_mask]).max(-1)
    assert (d == 0).mean() >= 0.55 and (d <= 1).mean() >= 0.85 and (d <= 5).mean() >= 0.95 and (d <= 20).mean() >= 0.99
    assert np.abs(got[ref_mask] - ref[ref_mask]).mean() <= 1.0
