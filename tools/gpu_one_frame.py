#!/usr/bin/env python3
"""Renders N frames of a bench workload through rt_render (run under gpurun / ncu): prints kernel ms per frame."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
flat, rc, features, desc = bench.build_workload(name, 0)
if len(sys.argv) > 4:  # reduced frame size (ncu captures of the multi-second configs): same scene, fewer pixels
    from hslu_i.ba_raytracing.f2501_raytracer_b200 import RenderConfig

    rc = RenderConfig.from_features(features, 0, width=int(sys.argv[3]), height=int(sys.argv[4]))
dev = DeviceScene(flat)
print(name, desc, dev.bvh_info(), flush=True)
for i in range(frames):
    _, st = dev.render(rc)
    print(f"frame {i}: kernel_ms={st['kernel_ms']:.3f} total_ms={st['total_ms']:.3f} kernel={dev.dominant_kernel_name()}", flush=True)
dev.close()
