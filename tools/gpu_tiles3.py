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
st = torch.cuda.current_stream().cuda_stream
def run(cfgs, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for c in cfgs:
            dev.render_device(c, slab.data_ptr(), mask.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
print("full frame:", run([rc]))
for world, tr in [(2, 2), (4, 2), (8, 2), (8, 8), (16, 2)]:
    cfgs = [rc.with_tiles(tr, r, world) for r in range(world)]
    print(f"world={world} tile_rows={tr} back-to-back total:", run(cfgs))
# same rank rendered 8 times back to back
print("rank1 x8:", run([rc.with_tiles(2, 1, 8)] * 8), "rank6 x8:", run([rc.with_tiles(2, 6, 8)] * 8))
