#!/usr/bin/env python3
"""Kernel-time / work-counter probe for tuning (run under gpurun): prints ms and tests per ray."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene, RenderConfig, example_scene
from hslu_i.ba_raytracing.f2501_raytracer_b200 import _lib as L

WL = {"c3": ("semesterbild", "high_quality,simd_render"), "c4": ("semesterbild", "realistic,high_resolution"),
      "c2": ("test_scene", "simd_render,backface_culling,medium_resolution"), "c1": ("main", "default,simd_render"),
      "c1lite": ("main", "")}
for name in sys.argv[1:] or ["c3"]:
    scene, feat = WL[name]
    flat = example_scene(scene, feat).flatten()
    dev = DeviceScene(flat)
    rc = RenderConfig.from_features(feat, 0)
    dev.render(rc)
    ms = min(dev.render(rc)[1]["kernel_ms"] for _ in range(5))
    st = dev.render(RenderConfig.from_features(feat, 0, extra_flags=L.FLAG_COUNT_WORK))[1]
    rays = st["closest_rays"] + st["shadow_rays"]
    print(f"{name}: kernel_ms={ms:.3f} Grays/s={rays/ms/1e6:.2f} rays={rays} box/ray={st['box_tests']/rays:.2f} tri/ray={st['triangle_tests']/rays:.2f} sph/ray={st['sphere_tests']/rays:.2f}", flush=True)
    dev.close()
