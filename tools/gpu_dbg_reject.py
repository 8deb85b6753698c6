#!/usr/bin/env python3
"""Where tri_test exits (debug build with -DRT_DBG_REJECT; tuning experiments only)."""
import ctypes as C
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene, RenderConfig, example_scene
from hslu_i.ba_raytracing.f2501_raytracer_b200 import _lib as L

lib = L.load()
flat = example_scene("semesterbild", "high_quality,simd_render").flatten()
dev = DeviceScene(flat)
for tag, extra in (("beam", 0), ("ray", L.FLAG_PER_RAY_SHADOWS)):
    rc = RenderConfig.from_features("high_quality,simd_render", 0, extra_flags=extra)
    out = (C.c_ulonglong * 8)()
    lib.rt_dbg_counters(out)
    dev.render(rc)
    lib.rt_dbg_counters(out)
    print(tag, "t_reject", out[1], "uv_reject", out[2], "hit", out[3], flush=True)
