#!/usr/bin/env python3
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene, RenderConfig, example_scene
from hslu_i.ba_raytracing.f2501_raytracer_b200 import _lib as L
feat = "high_quality,simd_render"
flat = example_scene("semesterbild", feat).flatten()
dev = DeviceScene(flat)
rc = RenderConfig.from_features(feat, 0)
rcc = RenderConfig.from_features(feat, 0, extra_flags=L.FLAG_COUNT_WORK)
slab = torch.zeros(768 * 640, dtype=torch.int32, device="cuda")
mask = torch.zeros(768 * 640, dtype=torch.uint8, device="cuda")
world, tile_rows = 8, 2
for order in (range(world), reversed(range(world))):
    for r in order:
        rcr = rc.with_tiles(tile_rows, r, world)
        dev.render_device(rcr, slab.data_ptr(), mask.data_ptr(), None, want_stats=True)
        t = min(dev.render_device(rcr, slab.data_ptr(), mask.data_ptr(), None, want_stats=True)["kernel_ms"] for _ in range(3))
        st = dev.render_device(rcc.with_tiles(tile_rows, r, world), slab.data_ptr(), mask.data_ptr(), None, want_stats=True)
        rays = st["closest_rays"] + st["shadow_rays"]
        print(f"rank {r}: ms={t:.3f} rays={rays/1e6:.2f}M box={st['box_tests']/1e9:.3f}G tri={st['triangle_tests']/1e9:.3f}G  ns/ray={t*1e6/rays:.4f}", flush=True)
