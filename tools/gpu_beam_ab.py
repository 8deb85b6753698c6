#!/usr/bin/env python3
"""A/B of shadow beams vs per-ray shadow traversal (run under gpurun): images must be identical."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene, RenderConfig, example_scene
from hslu_i.ba_raytracing.f2501_raytracer_b200 import _lib as L

WL = {"c3": ("semesterbild", "high_quality,simd_render"), "c3s": ("semesterbild", "high_quality"),
      "c1": ("main", "default,simd_render"), "soft": ("main", "anti_aliasing,soft_shadows"),
      "c2soft": ("test_scene", "simd_render,backface_culling,soft_shadows")}
for name in sys.argv[1:] or ["c3"]:
    scene, feat = WL[name]
    flat = example_scene(scene, feat).flatten()
    dev = DeviceScene(flat)
    out = {}
    for tag, extra in (("beam", 0), ("ray", L.FLAG_PER_RAY_SHADOWS)):
        rc = RenderConfig.from_features(feat, 0, extra_flags=extra)
        dev.render(rc)
        runs = [dev.render(rc) for _ in range(4)]
        ms = min(r[1]["kernel_ms"] for r in runs)
        st = dev.render(RenderConfig.from_features(feat, 0, extra_flags=extra | L.FLAG_COUNT_WORK))[1]
        rays = st["closest_rays"] + st["shadow_rays"]
        out[tag] = runs[0][0]
        print(f"{name}/{tag}: kernel_ms={ms:.3f} Grays/s={rays/ms/1e6:.2f} rays={rays} box/ray={st['box_tests']/rays:.2f} "
              f"tri/ray={st['triangle_tests']/rays:.2f} sph/ray={st['sphere_tests']/rays:.3f} beamtri/ray={st['beam_triangle_tests']/rays:.3f}", flush=True)
    same = np.array_equal(out["beam"].buffer, out["ray"].buffer) and np.array_equal(out["beam"].hit_mask, out["ray"].hit_mask)
    print(f"{name}: images identical = {same}; differing pixels = {int((out['beam'].buffer != out['ray'].buffer).sum())}", flush=True)
    dev.close()
