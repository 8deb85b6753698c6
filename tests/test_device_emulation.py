"""The product's device-side traversal and shadow-beam code (csrc/rt_device.cuh), compiled unchanged for the
HOST by tests/emu/device_emu.cpp, against the CPU oracle and against itself -- without a GPU:

* BVH4 traversal (trace_closest / trace_shadow over the arrays rt_bvh.cpp flattens) == the oracle's linear scan;
* a shadow beam's candidate list (make_beam + beam_collect + beam_filter + trace_shadow_list) gives every one of
  the M copies the result of its own BVH walk, bit for bit, on the shipped scenes.

Test infrastructure only: the emulation answers raw queries, it is not a CPU path of the product."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
EMU_DIR = ROOT / "tests" / "emu"
CSRC = ROOT / "hslu_i" / "ba_raytracing" / "f2501_raytracer_b200" / "csrc"
F = np.float32


def _build_emu(lib_name, defines=()):
    lib_path = EMU_DIR / lib_name
    srcs = [EMU_DIR / "device_emu.cpp", CSRC / "rt_bvh.cpp", CSRC / "rt_error.cpp"]
    deps = srcs + [CSRC / "rt_device.cuh", CSRC / "rt_bvh.h", ROOT / "include" / "rt_api.h"]
    if not lib_path.exists() or lib_path.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        cuda_inc = next((p for p in (Path("/usr/local/cuda/include"),) if (p / "cuda_runtime.h").exists()), None)
        if cuda_inc is None:
            pytest.skip("CUDA headers not found")
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-march=x86-64-v3", "-shared", "-pthread",
                        *defines, f"-I{cuda_inc}", f"-I{ROOT / 'include'}", f"-I{CSRC}", *map(str, srcs), "-o", str(lib_path)], check=True)
    lib = C.CDLL(str(lib_path))
    vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
    lib.emu_scene_create.restype = vp
    lib.emu_scene_create.argtypes = [vp]
    lib.emu_scene_destroy.argtypes = [vp]
    lib.emu_trace_closest.argtypes = [vp, u32, vp, u32, vp, vp]
    lib.emu_trace_shadow.argtypes = [vp, u32, vp, vp, u32, vp]
    lib.emu_light_box_behind.argtypes = [vp, vp, vp, vp, vp, u32, u32, vp, vp]
    lib.emu_beam_shadows.restype = C.c_int
    lib.emu_beam_shadows.argtypes = [vp, u32, vp, vp, vp, f32, vp, u32, vp, vp, vp]
    return lib


@pytest.fixture(scope="module", params=["product", "round1_filter"])
def emu(request):
    """"product": the device code as the product builds it (beam-filter switches RT_SPHERE_MISS_CULL / RT_SELF_CULL on).
    "round1_filter": the same code with both switches off (the round-1 filter) and the BVH walks in their node-or-leaf
    form (RT_WW_* = 0; the product's walks descend to a leaf first), which stays buildable for A/B runs."""
    if request.param == "product":
        return _build_emu("libdevice_emu.so")
    return _build_emu("libdevice_emu_r1.so", ("-DRT_SPHERE_MISS_CULL=0", "-DRT_SELF_CULL=0", "-DRT_WW_CLOSEST=0", "-DRT_WW_SHADOW_INLINE=0"))


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def random_rays(rng, n):
    o = rng.uniform(-0.3, 1.3, (n, 3)).astype(F)
    tgt = rng.uniform(0.0, 1.0, (n, 3)).astype(F)
    return np.ascontiguousarray(np.concatenate([o, tgt - o], 1), F)


@pytest.mark.parametrize("scene_name,features,n", [("test_scene", "backface_culling", 20000), ("main", "", 4000)])
def test_device_traversal_code_equals_linear_scan(pkg, oracle_mod, emu, scene_name, features, n):
    flat = pkg.example_scene(scene_name, features).flatten()
    desc = flat.desc()
    h = emu.emu_scene_create(C.addressof(desc))
    assert h
    try:
        rng = np.random.default_rng(21)
        rays = random_rays(rng, n)
        md = rng.uniform(0.05, 2.5, n).astype(F)
        for flags in (0, pkg._lib.FLAG_BACKFACE_CULLING):
            t = np.zeros(n, F)
            prim = np.zeros(n, np.uint32)
            emu.emu_trace_closest(h, flags, ptr(rays), n, ptr(t), ptr(prim))
            t_o, p_o = oracle_mod.trace_closest(desc, rays, flags, oracle_mod.BRUTE)
            assert np.array_equal(prim, p_o) and np.array_equal(t.view(np.uint32), t_o.view(np.uint32))
            sh = np.zeros((n, 5), F)
            emu.emu_trace_shadow(h, flags, ptr(rays), ptr(md), n, ptr(sh))
            s_o = oracle_mod.trace_shadow(desc, rays, md, flags, oracle_mod.BRUTE)
            assert np.array_equal(sh[:, 0], s_o[:, 0])
            free = s_o[:, 0] == 0
            assert np.allclose(sh[free, 1:], s_o[free, 1:], rtol=0, atol=2e-6)  # traversal order vs scan order
            assert (p_o != 0xFFFFFFFF).mean() > 0.3 and 0.05 < (~free).mean() < 0.95
    finally:
        emu.emu_scene_destroy(h)


def shading_points(pkg, emu, h, rc, pixels):
    """Primary hit points of pixel centres, as the kernels compute them (get_pixel_color: origin on z = 0)."""
    W, H = rc.cfg.width, rc.cfg.height
    sh = F(H) / F(W)
    sd = (F(1) + sh) / F(2)
    focus = np.array([0.5, sh / F(2), F(-1.9) * sd], F)
    o = np.stack([(pixels % W).astype(F) * (F(1) / F(W)), (pixels // W).astype(F) * (sh / F(H)), np.zeros(len(pixels), F)], 1).astype(F)
    rays = np.ascontiguousarray(np.concatenate([o, o - focus], 1), F)
    t = np.zeros(len(pixels), F)
    prim = np.zeros(len(pixels), np.uint32)
    emu.emu_trace_closest(h, 0, ptr(rays), len(pixels), ptr(t), ptr(prim))
    d = rays[:, 3:] / np.sqrt((rays[:, 3:] ** 2).sum(1, dtype=F))[:, None]
    hit = prim != 0xFFFFFFFF
    return (o[hit] + d[hit] * t[hit, None]).astype(F), pixels[hit]


@pytest.mark.parametrize("scene_name,features,extra_pixels", [
    ("semesterbild", "high_quality,simd_render", [190 * 768 + 226, 187 * 768 + 226, 243 * 768 + 706, 171 * 768 + 734, 8 * 768 + 666]),
    ("test_scene", "simd_render,backface_culling,soft_shadows", []),
    ("test_scene", "simd_render,soft_shadows", []),
    ("large_mesh", "high_quality,simd_render", []),
])
def test_shadow_beam_lists_equal_per_ray_traversal(pkg, emu, scene_name, features, extra_pixels):
    if scene_name == "large_mesh":
        from hslu_i.ba_raytracing.f2501_raytracer_b200.synthetic import large_mesh_scene

        flat = large_mesh_scene(features, nx=192, ny=192).flatten()
    else:
        flat = pkg.example_scene(scene_name, features).flatten()
    rc = pkg.RenderConfig.from_features(features, 3, n_light_clouds=256)
    M = rc.cfg.light_multiplicator
    assert M > 1
    W, H = rc.cfg.width, rc.cfg.height
    desc = flat.desc()
    h = emu.emu_scene_create(C.addressof(desc))
    try:
        rng = np.random.default_rng(5)
        pixels = np.unique(np.concatenate([rng.choice(W * H, 1000, replace=False), np.array(extra_pixels, np.int64)])).astype(np.int64)
        P, _ = shading_points(pkg, emu, h, rc, pixels)
        assert len(P) > 600
        sh = F(H) / F(W)
        w2s = np.array([F(1) / F(W), sh / F(H), ((F(1) + sh) / F(2)) / F((W + H) // 2)], F)
        eps_d = F(1.1920929e-7) * (F(100) * ((F(1) + sh + (F(1) + sh) / F(2)) / F(3)))
        clouds = rc.light_clouds.astype(F)  # (K, M, 3)
        lo, hi = clouds.reshape(-1, 3).min(0), clouds.reshape(-1, 3).max(0)
        flags = rc.cfg.flags & pkg._lib.FLAG_BACKFACE_CULLING
        totals = np.zeros(8, np.uint64)
        overflow = occluded = rays = covered = dark_beams = 0
        for i, p in enumerate(P):
            for li, light in enumerate(flat.lights):
                lpos = np.asarray(light["position"], F)
                lmin, lmax = (lpos + lo * w2s).astype(F), (lpos + hi * w2s).astype(F)
                copies = np.ascontiguousarray((lpos + clouds[(i * 7 + li * 13) % len(clouds)] * w2s).astype(F))
                assert (copies >= lmin).all() and (copies <= lmax).all()  # the bounds are exact in float arithmetic
                out_list, out_ray, counts = np.zeros((M, 5), F), np.zeros((M, 5), F), np.zeros(8, np.uint64)
                pc = np.ascontiguousarray(p)
                n_cand = emu.emu_beam_shadows(h, flags, ptr(pc), ptr(lmin), ptr(lmax), eps_d, ptr(copies), M, ptr(out_list), ptr(out_ray), ptr(counts))
                # an occluded query carries no other state (shade_pair returns at once; the last-occluder probe may
                # decide it before any transmissive hit is accumulated); everything else must agree bit for bit
                assert np.array_equal(out_list[:, 0], out_ray[:, 0]), (i, li, n_cand)
                free = out_ray[:, 0] == 0
                assert np.array_equal(out_list[free].view(np.uint32), out_ray[free].view(np.uint32)), (i, li, n_cand)
                assert (out_list[~free, 1] == 0).all() and (out_ray[~free, 1] == 0).all()
                overflow += n_cand == -1
                covered += n_cand == -2  # BEAM_COVERED: one opaque triangle occludes the whole beam (checked above: all occluded)
                dark_beams += bool((out_ray[:, 0] == 1).all())
                occluded += int(out_ray[:, 0].sum())
                rays += M
                totals += counts
        assert 0.05 < occluded / rays < 0.95
        assert overflow < 0.05 * len(P) * len(flat.lights)
        print(f"\n{scene_name}: {len(P) * len(flat.lights)} beams, {dark_beams} fully occluded, {covered} of them decided by the cover test, {overflow} overflowed")
        if not flags:
            assert covered > 0
        # the point of the exercise: one hull walk + a short list instead of M BVH walks
        assert totals[0] < totals[3] / 5 and totals[2] + totals[1] < totals[4]
    finally:
        emu.emu_scene_destroy(h)


def test_cover_test_never_claims_a_beam_with_a_free_ray(pkg, emu):
    """BEAM_COVERED (one opaque triangle / sphere occludes every ray of the beam, so none is traced) on random
    single-primitive scenes with beams aimed at the primitive's silhouette: whenever the device code claims it,
    every copy's own shadow query (trace_shadow, reference arithmetic) must be occluded."""
    rng = np.random.default_rng(77)
    mat = pkg.material([0.5, 0.5, 0.5])
    eps_d = F(1.0927e-5)
    M = 24
    claimed = {"tri": 0, "sphere": 0, "sphere_inside": 0}
    seen = {"tri": 0, "sphere": 0, "sphere_inside": 0}
    for case in range(1500):
        kind = ("tri", "sphere", "sphere_inside")[case % 3]
        s = pkg.Scene()
        if kind == "tri":
            scale = 10 ** rng.uniform(-2.0, -0.2)
            v1 = rng.uniform(0.2, 0.8, 3)
            v2, v3 = v1 + rng.normal(0, scale, 3), v1 + rng.normal(0, scale, 3)
            if np.linalg.norm(np.cross(v2 - v1, v3 - v1)) < 1e-10:
                continue
            s.add_triangle(v1.astype(F), v2.astype(F), v3.astype(F), mat)
            u, v = rng.uniform(-0.2, 1.2, 2)
            X = v1 + u * (v2 - v1) + v * (v3 - v1)  # aim: a point of the plane in / near the triangle
            P = X + rng.normal(0, 1, 3) * 10 ** rng.uniform(-1.5, 0)
            Cc = P + (X - P) * rng.uniform(1.02, 3.0)
        else:
            c = rng.uniform(0.3, 0.7, 3)
            r = 10 ** rng.uniform(-1.7, -0.6)
            s.add_sphere(c.astype(F), F(r), mat)
            dirn = rng.normal(0, 1, 3)
            dirn /= np.linalg.norm(dirn)
            if kind == "sphere":
                side = np.cross(dirn, rng.normal(0, 1, 3))
                side /= np.linalg.norm(side)
                X = c + side * r * rng.uniform(0, 1.2)  # aim inside / at / just outside the silhouette
                P = X - dirn * r * rng.uniform(1.1, 15)
                Cc = X + dirn * r * rng.uniform(0.3, 10)
            else:  # the shading point ON the sphere, the light beyond it (or not)
                P = c + dirn * np.float64(F(r))
                away = rng.normal(0, 1, 3)
                away /= np.linalg.norm(away)
                Cc = P + away * r * rng.uniform(0.5, 10)
        half = 10 ** rng.uniform(-3.3, -1.8, 3)
        lmin, lmax = (Cc - half).astype(F), (Cc + half).astype(F)
        copies = np.ascontiguousarray((lmin + rng.random((M, 3)).astype(F) * (lmax - lmin)).astype(F))
        copies[:8] = [[(lmin, lmax)[(k >> a) & 1][a] for a in range(3)] for k in range(8)]
        flat = s.flatten()
        desc = flat.desc()
        h = emu.emu_scene_create(C.addressof(desc))
        try:
            out_list, out_ray, counts = np.zeros((M, 5), F), np.zeros((M, 5), F), np.zeros(8, np.uint64)
            pc = np.ascontiguousarray(P.astype(F))
            n_cand = emu.emu_beam_shadows(h, 0, ptr(pc), ptr(lmin), ptr(lmax), eps_d, ptr(copies), M, ptr(out_list), ptr(out_ray), ptr(counts))
        finally:
            emu.emu_scene_destroy(h)
        seen[kind] += 1
        if n_cand == -2:
            claimed[kind] += 1
            assert (out_ray[:, 0] == 1).all(), f"case {case} ({kind}): covered claimed but {int((out_ray[:, 0] == 0).sum())} copies are unoccluded"
        else:
            assert np.array_equal(out_list[:, 0], out_ray[:, 0])
    assert all(claimed[k] > 20 for k in claimed), (claimed, seen)


def test_beam_degenerate_inputs_equal_per_ray_traversal(pkg, emu):
    """Corners of the beam construction on the semesterbild scene: a degenerate light box (one point), a shading
    point inside the light box (the hull widens both ways on every axis), axis-aligned beams, a NaN shading point."""
    feat = "high_quality,simd_render"
    flat = pkg.example_scene("semesterbild", feat).flatten()
    desc = flat.desc()
    h = emu.emu_scene_create(C.addressof(desc))
    eps_d = F(1.0927e-5)
    M = 8
    rng = np.random.default_rng(12)
    try:
        cases = []
        for _ in range(60):
            P = rng.uniform(0.05, 0.95, 3).astype(F)
            L = rng.uniform(0.05, 0.95, 3).astype(F)
            cases.append((P, L, L))                                                   # all copies in one point
            cases.append((P, (P - F(0.002)).astype(F), (P + F(0.003)).astype(F)))     # the point inside the light box
            a = rng.integers(3)
            L2 = P.copy()
            L2[a] = F(rng.uniform(0.05, 0.95))
            cases.append((P, L2, (L2 + F(0.003)).astype(F)))                          # beam along one axis
        cases.append((np.array([np.nan, 0.5, 0.5], F), np.array([0.2, 0.2, 0.2], F), np.array([0.21, 0.21, 0.21], F)))
        for P, lmin, lmax in cases:
            copies = np.ascontiguousarray((lmin + rng.random((M, 3)).astype(F) * (lmax - lmin)).astype(F))
            out_list, out_ray, counts = np.zeros((M, 5), F), np.zeros((M, 5), F), np.zeros(8, np.uint64)
            pc, lo, hi = np.ascontiguousarray(P), np.ascontiguousarray(lmin), np.ascontiguousarray(lmax)
            emu.emu_beam_shadows(h, 0, ptr(pc), ptr(lo), ptr(hi), eps_d, ptr(copies), M, ptr(out_list), ptr(out_ray), ptr(counts))
            assert np.array_equal(out_list[:, 0], out_ray[:, 0]), (P, lmin, lmax)
            free = out_ray[:, 0] == 0
            assert np.array_equal(out_list[free].view(np.uint32), out_ray[free].view(np.uint32))
    finally:
        emu.emu_scene_destroy(h)


def test_a_light_box_behind_the_surface_has_no_facing_copy(emu):
    """RT_FACING_FIRST (rt_kernels.cuh): direct_kernel builds no beam towards a light whose box of copies is entirely
    behind the surface (light_box_behind, rt_device.cuh), relying on every copy then failing shade_pair's own
    N.L > 0 test.  Random and adversarial cases -- grazing boxes, boxes touching the tangent plane, corner copies,
    huge and tiny scales, NaN normals -- with the per-copy test evaluated by shade_pair's operations."""
    rng = np.random.default_rng(5)
    n, m = 200000, 12
    P = rng.uniform(-1, 1, (n, 3)).astype(F)
    N = rng.normal(size=(n, 3)).astype(F)
    N /= np.linalg.norm(N, axis=1, keepdims=True).astype(F)
    N[::17] *= F(0.37)  # (interpolated shading normals are not unit length)
    centre = P + rng.normal(size=(n, 3)).astype(F) * rng.choice([1e-3, 0.05, 1.0, 30.0], (n, 1)).astype(F)
    half = np.abs(rng.normal(size=(n, 3))).astype(F) * rng.choice([0.0, 1e-4, 0.02, 0.5], (n, 1)).astype(F)
    # a third of the cases: the box pushed so that its most-facing corner lies (almost) in the tangent plane
    k = n // 3
    reach = np.sum(np.abs(N[:k]) * half[:k], axis=1) + np.sum(N[:k] * (centre[:k] - P[:k]), axis=1)
    nn = np.sum(N[:k] * N[:k], axis=1)
    centre[:k] -= (N[:k] * (reach / nn)[:, None] * rng.choice([0.999, 1.0, 1.0001, 1.001, 1.5], (k, 1))).astype(F)
    scale = rng.choice([1e-3, 1.0, 1e3], (n, 1)).astype(F)
    P, centre, half = (P * scale).astype(F), (centre * scale).astype(F), (half * scale).astype(F)
    lmin, lmax = (centre - half).astype(F), (centre + half).astype(F)
    u = rng.random((n, m, 3)).astype(F)
    u[:, :8] = np.array([[(i >> 0) & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)], F)  # the corners first
    copies = np.ascontiguousarray(np.clip(lmin[:, None, :] + u * (lmax - lmin)[:, None, :], lmin[:, None, :], lmax[:, None, :]).astype(F))
    N[5] = np.nan
    P[6, 1] = np.nan
    behind, facing = np.zeros(n, np.int32), np.zeros(n, np.int32)
    P, N, lmin, lmax = (np.ascontiguousarray(a, F) for a in (P, N, lmin, lmax))
    emu.emu_light_box_behind(ptr(P), ptr(N), ptr(lmin), ptr(lmax), ptr(copies), n, m, ptr(behind), ptr(facing))
    assert behind[5] == 0 and behind[6] == 0
    assert not np.any((behind == 1) & (facing > 0))
    assert behind.mean() > 0.2 and (facing == 0).mean() < 0.7  # both outcomes are exercised
    # the test is not vacuous where it matters: most boxes without a facing copy are recognised
    assert ((behind == 1) & (facing == 0)).sum() > 0.8 * (facing == 0).sum()
