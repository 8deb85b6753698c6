#!/usr/bin/env python3
"""Run-to-run determinism probe (run under gpurun): renders a workload several times under different switches and
reports how many pixels differ from the first render of the first variant."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
flat, rc, features, desc = bench.build_workload(name, 0)
dev = DeviceScene(flat)
ref = None
for tag, env in [("default", {}), ("default again", {}), ("no deferral", {"RT_DEFER_BELOW": "0"}), ("always deferral", {"RT_DEFER_BELOW": "1000000"}),
                 ("megakernel", {"RT_WAVEFRONT": "0"}), ("small chunks", {"RT_WAVE_CHUNK": "1048576"}),
                 ("per-light units, one window", {"RT_LIGHTS_PER_UNIT": "1", "RT_CHAIN_WINDOW": "100000000"}),
                 ("per-light units, windows", {"RT_LIGHTS_PER_UNIT": "1", "RT_DEFER_BELOW": "0"}), ("two lights per unit", {"RT_LIGHTS_PER_UNIT": "2", "RT_DEFER_BELOW": "0"})]:
    for k in ("RT_DEFER_BELOW", "RT_WAVEFRONT", "RT_WAVE_CHUNK", "RT_CHAIN_WINDOW", "RT_LIGHTS_PER_UNIT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for rep in range(2):
        buf, st = dev.render(rc)
        if ref is None:
            ref = buf
        diff = buf.buffer != ref.buffer
        d = np.zeros(len(buf.buffer), np.int32)
        for sh in (16, 8, 0):
            d = np.maximum(d, np.abs(((buf.buffer >> sh) & 255).astype(np.int32) - ((ref.buffer >> sh) & 255).astype(np.int32)))
        ys, xs = np.nonzero(diff.reshape(rc.cfg.height, rc.cfg.width))
        where = list(zip(xs[:4].tolist(), ys[:4].tolist()))
        print(f"{name} {tag} rep {rep}: kernel_ms={st['kernel_ms']:.2f} differing pixels={int(diff.sum())} max channel diff={int(d.max())} mask equal={bool(np.array_equal(buf.hit_mask, ref.hit_mask))} first={where} kernel={dev.dominant_kernel_name()[:60]}", flush=True)
dev.close()
