"""The oracle's private BVH mode (used only to make full-frame oracle images affordable) must be
bit-identical to its linear-scan mode, which is the restatement of the reference."""
import numpy as np
import pytest


def random_rays(rng, n, lo=-0.3, hi=1.3):
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    tgt = rng.uniform(0.0, 1.0, (n, 3)).astype(np.float32)
    return np.concatenate([o, tgt - o], 1).astype(np.float32)


@pytest.mark.parametrize("scene_name,features", [("test_scene", "backface_culling"), ("main", "")])
def test_raw_queries_identical(pkg, oracle_mod, scene_name, features):
    flat = pkg.example_scene(scene_name, features).flatten()
    rng = np.random.default_rng(11)
    n = 3000 if scene_name == "main" else 20000
    rays = random_rays(rng, n)
    for flags in (0, pkg._lib.FLAG_BACKFACE_CULLING):
        t0, p0 = oracle_mod.trace_closest(flat.desc(), rays, flags, oracle_mod.BRUTE)
        t1, p1 = oracle_mod.trace_closest(flat.desc(), rays, flags, oracle_mod.BVH)
        assert np.array_equal(p0, p1)
        assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
        md = rng.uniform(0.05, 2.5, n).astype(np.float32)
        s0 = oracle_mod.trace_shadow(flat.desc(), rays, md, flags, oracle_mod.BRUTE)
        s1 = oracle_mod.trace_shadow(flat.desc(), rays, md, flags, oracle_mod.BVH)
        occ = s0[:, 0] == 1
        assert np.array_equal(s0[:, 0], s1[:, 0])
        assert np.array_equal(s0[~occ].view(np.uint32), s1[~occ].view(np.uint32))  # bitwise on unoccluded rays
        assert (p0 != 0xFFFFFFFF).mean() > 0.3 and 0.05 < occ.mean() < 0.95


def test_rendered_pixels_identical_with_bounces_and_soft_shadows(pkg, oracle_mod):
    feat = "realistic,anti_aliasing,soft_shadows,simd_render"
    flat = pkg.example_scene("main", feat).flatten()
    rc = pkg.RenderConfig.from_features(feat, 5, n_light_clouds=64)
    rng = np.random.default_rng(5)
    pix = rng.choice(rc.cfg.width * rc.cfg.height, 24, replace=False).astype(np.uint32)
    a0, m0, c0, n0 = oracle_mod.render(flat.desc(), rc.cfg, oracle_mod.BRUTE, 0, pix, want_rgb=True)
    a1, m1, c1, n1 = oracle_mod.render(flat.desc(), rc.cfg, oracle_mod.BVH, 0, pix, want_rgb=True)
    assert np.array_equal(m0, m1) and np.array_equal(a0, a1)
    assert np.array_equal(c0.view(np.uint32), c1.view(np.uint32))
    assert n0["closest_rays"] == n1["closest_rays"] and n0["shadow_rays"] == n1["shadow_rays"]
    assert n1["triangle_tests"] < n0["triangle_tests"] / 20


def test_packet8_mode_identical_to_scalar_scan(pkg, oracle_mod):
    """mode 2 (8-ray AVX2 packets, the `simd_render` stand-in timed by bench.py) == mode 0 per lane."""
    feat = "high_quality,simd_render"
    flat = pkg.example_scene("main", "").flatten()  # lowres text keeps the brute-force scan short
    rc = pkg.RenderConfig.from_features(feat, 5, n_light_clouds=64)
    rng = np.random.default_rng(8)
    pix = rng.choice(rc.cfg.width * rc.cfg.height, 16, replace=False).astype(np.uint32)
    pix = np.concatenate([pix, np.array([100 * 768 + 30, 320 * 768 + 384], np.uint32)])
    a0, m0, c0, n0 = oracle_mod.render(flat.desc(), rc.cfg, oracle_mod.BRUTE, 0, pix, want_rgb=True)
    a2, m2, c2, n2 = oracle_mod.render(flat.desc(), rc.cfg, 2, 0, pix, want_rgb=True)
    assert np.array_equal(m0, m2) and np.array_equal(a0, a2)
    assert np.array_equal(c0.view(np.uint32), c2.view(np.uint32))
    assert n0["closest_rays"] == n2["closest_rays"]
    assert n2["shadow_rays"] >= n0["shadow_rays"]  # masked (missed) lanes still ride along in a packet
