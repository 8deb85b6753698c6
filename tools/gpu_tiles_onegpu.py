#!/usr/bin/env python3
"""One GPU renders each rank's share of an N-way row-tiled C3 frame in turn: separates content
imbalance / small-launch inefficiency from GPU-to-GPU variation (run under gpurun)."""
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
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
W, H = rc.cfg.width, rc.cfg.height
slab = torch.zeros(W * H, dtype=torch.int32, device="cuda")
mask = torch.zeros(W * H, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for tr in [int(x) for x in (sys.argv[2:] or ["1", "8"])]:
    doms, rays = [], []
    for r in range(world):
        rcr = rc.with_tiles(tr, r, world) if world > 1 else rc
        for _ in range(2):
            dev.render_device(rcr, slab.data_ptr(), mask.data_ptr(), None)
        for _ in range(5):
            flush.zero_()
            dev.render_device(rcr, slab.data_ptr(), mask.data_ptr(), None)
        torch.cuda.synchronize()
        doms.append(float(dev.kernel_times(5).mean()))
        rcc = RenderConfig.from_features(feat, 0, extra_flags=L.FLAG_COUNT_WORK).with_tiles(tr, r, world) if world > 1 else RenderConfig.from_features(feat, 0, extra_flags=L.FLAG_COUNT_WORK)
        st = dev.render_device(rcc, slab.data_ptr(), mask.data_ptr(), None, want_stats=True)
        rays.append(st["shadow_rays"])
    print(f"world={world} tile_rows={tr} dominant_ms={[round(x, 3) for x in doms]} sum={sum(doms):.3f} Mrays={[round(x / 1e6, 2) for x in rays]} "
          f"ns/ray={[round(1e6 * d / x, 4) for d, x in zip(doms, rays)]}", flush=True)
dev.close()
