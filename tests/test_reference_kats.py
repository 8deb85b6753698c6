"""Hand-computed known-answer tests for the parts of the path that `output.png` (default,simd_render) does NOT pin:
the scalar anti-aliasing pattern (9 samples, bias direction 0 for every sample), the packet pattern with and
without rotation/scale, the crate constants at all three resolutions, the `soft_shadows` (M = 10) and
`extreme_quality` (S = 24, M = 28, depth 21/21) constants, and the per-ray `backface_culling` term.

Every expected value is evaluated HERE, in numpy float32, from the reference source lines cited next to it --
not by the oracle and not by the product -- so a misreading shared by oracle and kernels cannot pass.
No GPU needed: the product side is queried through rt_frame_constants / rth_config_from_features (host code of
librt_b200.so), the oracle through its raw-query entry points."""
import ctypes as C

import numpy as np
import pytest

F = np.float32
EPS = F(1.1920929e-7)


def frame_constants(pkg, rc):
    lib = pkg._lib.load()
    k = np.zeros(8, F)
    off = np.zeros((24, 2), F)
    n = C.c_uint32(0)
    pkg._lib.check(lib.rt_frame_constants(C.byref(rc.cfg), pkg._lib.ptr(k), pkg._lib.ptr(off), C.byref(n)))
    return k, off[: n.value]


@pytest.mark.parametrize("features,W,H", [("", 768, 640), ("medium_resolution", 1140, 950), ("high_resolution", 1620, 1350)])
def test_crate_constants(pkg, features, W, H):
    """src/lib.rs:73-92 and src/vector.rs:697-700, evaluated in f32 in the order the consts are written."""
    rc = pkg.RenderConfig.from_features(features, 0)
    assert (rc.cfg.width, rc.cfg.height) == (W, H)
    k, _ = frame_constants(pkg, rc)
    aspect = F(H) / F(W)                                   # WINDOW_ASPECT_RATIO
    sw = F(1.0)
    sh = sw * aspect                                       # SCENE_HEIGHT
    sd = (sw + sh) / F(2.0)                                # SCENE_DEPTH
    avg = (sw + sh + sd) / F(3.0)                          # AVERAGE_SCENE_DIMENSION
    depth_px = (W + H) // 2                                # WINDOW_SCENE_DEPTH (usize)
    want = [sw / F(W), sh / F(H), sd / F(depth_px), sw / F(2.0), sh / F(2.0), F(-1.9) * sd, EPS * (F(100.0) * avg), F(1.0)]
    assert np.array_equal(k.view(np.uint32), np.array(want, F).view(np.uint32))


def hand_aa_offsets(W, H, table, packet, rotation):
    """raytracer_renderer.rs:876-916 (directions) and :1021-1138 (bundling): sample i sits at
    coords + (table[i] * W2S * scale) * bias[i % LANES]; LANES = 8 in the packet path, 1 in the scalar path (the
    `directions.iter().cycle()` restarts for every chunk of LANES samples, so a 1-lane chunk always gets bias[0] = t)."""
    aspect = F(H) / F(W)
    w2s_w, w2s_h = F(1.0) / F(W), (F(1.0) * aspect) / F(H)
    if rotation:
        ang = np.arctan(F(0.5)).astype(F)
        sin, cos = np.sin(ang).astype(F), np.cos(ang).astype(F)
        # x.mul_add(cos, y * sin), x.mul_add(-sin, y * cos) on unit_x / unit_y: products with 0 and 1 are exact
        xr = np.array([cos, sin, 0], F)
        yr = np.array([-sin, cos, 0], F)
        scale = F(np.sqrt(F(5.0))) / F(2.05)
    else:
        xr, yr = np.array([1, 0, 0], F), np.array([0, 1, 0], F)
        scale = F(0.85)
    t, l, r, b = -yr, -xr, xr, yr
    dirs = [t, l, b, r, t + l, t + r, b + l, b + r]

    def normalized(v):  # ultraviolet: v * (1 / sqrt(dot)), dot = fma chain (exactly representable products here or not: f32 throughout)
        v = v.astype(F)
        d = F(np.float64(v[0]) * np.float64(v[0]) + F(np.float64(v[1]) * np.float64(v[1]) + np.float64(F(v[2] * v[2]))))  # fma(x,x,fma(y,y,z*z)) with z = 0
        return (v * (F(1.0) / np.sqrt(d, dtype=F))).astype(F)

    dirs = [normalized(v) for v in dirs]
    out = []
    for i, p in enumerate(table):
        bias = dirs[i % 8] if packet else dirs[0]
        rx, ry = F(p[0]) * w2s_w * scale, F(p[1]) * w2s_h * scale
        out.append([rx * bias[0], ry * bias[1]])
    return np.array(out, F)


def test_scalar_aa_pattern_is_nine_samples_with_the_first_bias(pkg):
    """`high_quality` without simd_render: S = 9, deterministic table [0,0] + 8 x [1,1], no rotation: bias[0] = t =
    (0,-1,0), so eight samples coincide 0.85 pixel heights ABOVE the centre and dx is (+-)0 for all of them."""
    rc = pkg.RenderConfig.from_features("high_quality", 0)
    assert rc.samples_per_pixel == 9 and rc.cfg.aa_table_len == 16
    assert np.array_equal(rc.aa_table, np.array([[0, 0]] + [[1, 1]] * 15, F))
    k, off = frame_constants(pkg, rc)
    want = hand_aa_offsets(768, 640, rc.aa_table[:9], packet=False, rotation=False)
    assert off.shape == (9, 2) and np.array_equal(np.abs(off).view(np.uint32), np.abs(want).view(np.uint32))
    assert (off[:, 0] == 0).all() and off[0, 1] == 0 and (off[1:, 1] == -(F(1) * k[1] * F(0.85))).all()
    assert k[7] == F(1.0) / F(9.0)  # weight of a sample (raytracer_renderer.rs:936-937)


@pytest.mark.parametrize("features,S", [("high_quality,simd_render", 16), ("default,simd_render", 16), ("extreme_quality,simd_render", 24),
                                        ("extreme_quality,anti_aliasing_rotation_scale,anti_aliasing_randomness", 24)])
def test_packet_and_extreme_aa_patterns(pkg, features, S):
    rc = pkg.RenderConfig.from_features(features, 5)
    packet = "simd_render" in features
    rotation = "default" in features or "rotation" in features
    assert rc.samples_per_pixel == S
    k, off = frame_constants(pkg, rc)
    want = hand_aa_offsets(rc.cfg.width, rc.cfg.height, rc.aa_table[:S], packet=packet, rotation=rotation)
    # the device table may differ from the hand evaluation in the sign of a zero only
    assert off.shape == (S, 2) and np.allclose(off, want, rtol=0, atol=0) and np.array_equal(np.abs(off).view(np.uint32), np.abs(want).view(np.uint32))
    assert k[7] == F(1.0) / F(S)
    if packet and not rotation:  # bias j = i % 8: samples 1..8 of the deterministic table go t, l, b, r, tl, tr, bl, br
        px, py = F(1) * k[0] * F(0.85), F(1) * k[1] * F(0.85)
        assert off[2, 1] == py and off[3, 0] == px and off[1, 0] == -px and off[8, 0] == 0 and off[8, 1] == -py  # i = 8 -> bias 0 again


def test_quality_constants(pkg):
    """raytracer_renderer.rs:55-93: depths 9/8, 13/18, 21/21; M = 1 / 10 / 19 / 28; S = 9 / 24."""
    cases = {
        "reflections,refractions": (9, 8, 1, 9),
        "soft_shadows": (9, 8, 10, 9),
        "high_quality": (13, 18, 19, 9),
        "high_quality,realistic": (13, 18, 19, 9),
        "extreme_quality": (21, 21, 28, 24),
    }
    for feat, (refl, refr, M, S) in cases.items():
        rc = pkg.RenderConfig.from_features(feat, 0, n_light_clouds=4)
        c = rc.cfg
        assert (c.reflection_max_depth, c.refraction_max_depth, c.light_multiplicator, c.aa_samples_per_pixel) == (refl, refr, M, S), feat
        assert c.aa_table_len == -(-S // 8) * 8
    # Cargo.toml:62-83: extreme_quality implies high_quality, which implies anti_aliasing + soft_shadows + high_quality_model
    rc = pkg.RenderConfig.from_features("extreme_quality", 0, n_light_clouds=4)
    L = pkg._lib
    assert rc.cfg.flags & L.FLAG_ANTI_ALIASING and rc.high_quality_model
    assert not rc.cfg.flags & (L.FLAG_REFLECTIONS | L.FLAG_REFRACTIONS | L.FLAG_PACKET_SEMANTICS)


@pytest.mark.parametrize("M", [10, 19, 28])
def test_light_cloud_geometry(pkg, M):
    """light.rs:197-206: M points of a Poisson-disk process on [0, R)^3, R = 1.725 + M / 20, radius 4 / M."""
    lib = pkg._lib.load()
    c = np.zeros((32, M, 3), F)
    assert lib.rt_sampler_light_clouds(11, M, 32, pkg._lib.ptr(c)) == 0
    R, r = 1.725 + M / 20.0, 4.0 / M
    assert (c >= 0).all() and (c < R).all()
    for cloud in c:
        d = np.linalg.norm(cloud[:, None, :].astype(np.float64) - cloud[None, :, :], axis=-1) + np.eye(M) * 10
        assert d.min() >= r * (1 - 1e-6)


def test_per_ray_backface_term(pkg, oracle_mod):
    """triangle.rs:154-168 / sphere.rs:137-151: with `backface_culling` a hit only counts when dot(d, n_stored) < 0.75
    or the material is transmissive.  Hand cases: a triangle facing the ray (dot = -1), facing away (dot = +1: culled),
    tilted to dot = 0.6 (kept) and 0.8 (culled), and the same back-facing triangle made of glass (kept)."""
    L = pkg._lib
    opaque, glass = pkg.material([0.5, 0.5, 0.5]), pkg.material([0.5, 0.5, 0.5], opacity=0.5, ior=1.5)
    v = ([-1, -1, 2], [1, -1, 2], [-1, 1, 2])

    def hit(normal, mat, flags):
        s = pkg.Scene()
        s.add_triangle(*v, mat, normal=normal)
        flat = s.flatten()
        t, prim = oracle_mod.trace_closest(flat.desc(), np.array([[-0.5, -0.5, 0, 0, 0, 1]], F), flags, oracle_mod.BRUTE)
        return prim[0] != 0xFFFFFFFF

    n_of = lambda dot: [float(np.sqrt(1 - dot * dot)), 0.0, dot]  # dot(d, n) with d = +z
    for dot, culled in [(-1.0, False), (1.0, True), (0.6, False), (0.8, True), (0.7499, False), (0.7501, True)]:
        assert hit(n_of(dot), opaque, 0)                                   # without the feature every orientation hits
        assert hit(n_of(dot), opaque, L.FLAG_BACKFACE_CULLING) == (not culled), dot
        assert hit(n_of(dot), glass, L.FLAG_BACKFACE_CULLING)              # transmissive: never culled
