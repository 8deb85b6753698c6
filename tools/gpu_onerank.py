#!/usr/bin/env python3
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene, RenderConfig, example_scene
feat = "high_quality,simd_render"
flat = example_scene("semesterbild", feat).flatten()
dev = DeviceScene(flat)
rc = RenderConfig.from_features(feat, 0)
slab = torch.zeros(768 * 640, dtype=torch.int32, device="cuda")
mask = torch.zeros(768 * 640, dtype=torch.uint8, device="cuda")
for cfg in [rc, rc.with_tiles(2, 1, 8), rc.with_tiles(80, 3, 8), rc.with_tiles(2, 1, 16)]:
    for _ in range(2):
        dev.render_device(cfg, slab.data_ptr(), mask.data_ptr(), None, want_stats=True)
