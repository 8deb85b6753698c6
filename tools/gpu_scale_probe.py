#!/usr/bin/env python3
"""Strong-scaling probe (run under torchrun on N GPUs): step time and per-rank kernel time of the
C3 frame for several row-tile heights.  Prints one line per tile height on rank 0."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene, RenderConfig, example_scene
from hslu_i.ba_raytracing.f2501_raytracer_b200.distributed import TiledRenderer

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
feat = "high_quality,simd_render"
flat = example_scene("semesterbild", feat).flatten()
dev = DeviceScene(flat, device=local)
rc = RenderConfig.from_features(feat, 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
steps = int(os.environ.get("STEPS", 10))
for tr in [int(x) for x in (sys.argv[1:] or ["1", "2", "4", "8"])]:
    tiled = TiledRenderer(dev, rc, rank, world, tr)
    for _ in range(3):
        tiled.step()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for i in range(steps):
        flush.zero_()
        ev[i][0].record()
        tiled.render_local()
        ev[i][1].record()
        tiled.gather()
        ev[i][2].record()
    torch.cuda.synchronize()
    step = sum(e[0].elapsed_time(e[2]) for e in ev) / steps
    rend = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
    dom = float(dev.kernel_times(steps).mean())
    t = torch.tensor([step, rend, dom], dtype=torch.float64, device="cuda")
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
    else:
        allt = [t]
    if rank == 0:
        steps_ms = [float(a[0]) for a in allt]
        rend_ms = [float(a[1]) for a in allt]
        dom_ms = [float(a[2]) for a in allt]
        print(f"world={world} tile_rows={tr} step_ms_max={max(steps_ms):.3f} render_ms per rank={[round(x, 3) for x in rend_ms]} "
              f"dominant_ms per rank={[round(x, 3) for x in dom_ms]} sum_dominant={sum(dom_ms):.3f}", flush=True)
dev.close()
if world > 1:
    dist.destroy_process_group()
