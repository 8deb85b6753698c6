import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import hslu_i.ba_raytracing.f2501_raytracer_b200 as pkg
feat = "high_quality,simd_render"
flat = pkg.example_scene("semesterbild", feat).flatten()
dev = pkg.DeviceScene(flat)
rc = pkg.RenderConfig.from_features(feat, 4, width=384, height=320, n_light_clouds=512)
full, _ = dev.render(rc)
full2, _ = dev.render(rc)
print("full deterministic:", np.array_equal(full.buffer, full2.buffer))
for world, tr in [(2, 2), (4, 8)]:
    comp = pkg.ImageBuffer(rc.cfg.width, rc.cfg.height)
    for r in range(world):
        dev.render(rc.with_tiles(tr, r, world), comp)
    diff = comp.buffer != full.buffer
    ys, xs = np.nonzero(diff.reshape(320, 384))
    print(world, tr, "ndiff", diff.sum(), "mask diff", (comp.hit_mask != full.hit_mask).sum(), "rows", np.unique(ys)[:20], "cols", np.unique(xs)[:20])
    for y, x in list(zip(ys, xs))[:8]:
        print("  ", x, y, hex(int(comp.buffer[y * 384 + x])), hex(int(full.buffer[y * 384 + x])))
