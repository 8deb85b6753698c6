#!/usr/bin/env python3
"""Per-warp timeline of direct_kernel (debug build with -DRT_DBG_TAIL; tuning experiments only)."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene, RenderConfig, example_scene
from hslu_i.ba_raytracing.f2501_raytracer_b200 import _lib as L

lib = L.load()
feat = "high_quality,simd_render"
flat = example_scene("semesterbild", feat).flatten()
dev = DeviceScene(flat)
rc = RenderConfig.from_features(feat, 0)
W, H = rc.cfg.width, rc.cfg.height
slab = torch.zeros(W * H, dtype=torch.int32, device="cuda")
mask = torch.zeros(W * H, dtype=torch.uint8, device="cuda")
NW = 148 * 7 * 4  # (the DEFER variant runs 6 CTAs per SM: its last 592 rows stay zero)
for world, r in ((1, 0), (8, 0), (8, 6)):
    rcr = rc.with_tiles(8, r, world) if world > 1 else rc
    for _ in range(3):
        dev.render_device(rcr, slab.data_ptr(), mask.data_ptr(), None)
    torch.cuda.synchronize()
    out = np.zeros((NW, 12), np.uint64)
    lib.rt_dbg_tail(C.c_void_p(out.ctypes.data), NW)
    t0 = out[:, 0].min()
    start = (out[:, 0] - t0).astype(np.float64) / 1e3
    end = (out[:, 1] - t0).astype(np.float64) / 1e3
    last = (out[:, 2] - t0).astype(np.float64) / 1e3
    spin = out[:, 3].astype(np.float64) / 1e3
    units = out[:, 4].astype(np.float64)
    q = lambda a: [round(float(np.percentile(a, p)), 1) for p in (0, 1, 10, 50, 90, 99, 100)]
    print(f"world={world} rank={r} kernel_ms={dev.kernel_times(1)[0]:.3f}")
    print("  warp start us  pct(0,1,10,50,90,99,100):", q(start))
    print("  warp end us    :", q(end))
    print("  last unit len  :", q(end - last))
    print("  spin us / warp :", q(spin), "total spin share", round(float(spin.sum() / (end - start).sum()), 4))
    print("  units / warp   :", q(units), " deferred units total", int(out[:, 7].sum()),
          "phase-B us / warp:", q(out[:, 11].astype(np.float64) / 1e3), flush=True)
    worst = out[:, 5].astype(np.float64) / 1e3
    print("  worst unit of each warp, us:", q(worst), " warps with a unit > 300 us:", int((worst > 300).sum()), " > 600 us:", int((worst > 600).sum()))
    order = np.argsort(-worst)[:6]
    n_items = rcr.local_rows() * W * 16
    n_batches = (n_items + 31) // 32
    for w in order:
        info = int(out[w, 6])
        unit, lanes, ncand = info >> 32, (info >> 8) & 0xFF, info & 0xFF
        g, b = divmod(unit, n_batches)
        pix = b * 2
        tile, within = divmod(pix, 32)
        ty, tx = divmod(tile, (W + 7) // 8)
        print(f"   worst unit of warp {w}: {worst[w]:.0f} us, group {g}, batch {b} -> local pixel ({tx * 8 + within % 8}, {ty * 4 + within // 8}), shared {lanes & 1}, max n_cand {ncand}; phases us: setup+chain+beam {out[w, 8] / 1e3:.0f}, own copies {out[w, 9] / 1e3:.0f}, wait helpers {out[w, 10] / 1e3:.0f}")
dev.close()
