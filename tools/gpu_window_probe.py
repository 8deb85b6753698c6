#!/usr/bin/env python3
"""Chain-window probe (run under gpurun): renders a 384x320 headline frame with several RT_CHAIN_WINDOW values and
reports the kernel time or the watchdog error."""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene, RenderConfig, RtError, example_scene  # noqa: E402

feat = "high_quality,simd_render"
flat = example_scene("semesterbild", feat).flatten()
dev = DeviceScene(flat)
rc = RenderConfig.from_features(feat, 2, width=384, height=320, n_light_clouds=512)
os.environ["RT_DEFER_BELOW"] = "0"
os.environ["RT_CHAIN_WINDOW"] = "1000000"
ref, _ = dev.render(rc)
for w in [int(x) for x in (sys.argv[1:] or ["8288", "8289", "8300", "9000", "10000", "12432", "16576", "8288"])]:
    os.environ["RT_CHAIN_WINDOW"] = str(w)
    for rep in range(2):
        t0 = time.perf_counter()
        try:
            got, st = dev.render(rc)
            ok = bool(np.array_equal(got.buffer, ref.buffer))
            print(f"window {w} rep {rep}: kernel_ms={st['kernel_ms']:.3f} identical={ok}", flush=True)
        except RtError as e:
            print(f"window {w} rep {rep}: ERROR after {time.perf_counter() - t0:.2f} s: {e}", flush=True)
dev.close()
