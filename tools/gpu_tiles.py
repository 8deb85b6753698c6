#!/usr/bin/env python3
"""Per-rank kernel time of the row-tile partition, emulated on one GPU (run under gpurun)."""
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
import os
for world, tile_rows in [tuple(map(int, a.split('x'))) for a in (sys.argv[1:] or ['1x2', '8x2', '4x2'])]:
    times = []
    for r in range(world):
        rcr = rc.with_tiles(tile_rows, r, world) if world > 1 else rc
        dev.render_device(rcr, slab.data_ptr(), mask.data_ptr(), None, want_stats=True)
        t = min(dev.render_device(rcr, slab.data_ptr(), mask.data_ptr(), None, want_stats=True)["kernel_ms"] for _ in range(3))
        times.append(t)
    print(f"world={world} tile_rows={tile_rows}: max={max(times):.3f} min={min(times):.3f} sum={sum(times):.3f} ideal={sum(times)/world:.3f}", [round(t, 2) for t in times], flush=True)
