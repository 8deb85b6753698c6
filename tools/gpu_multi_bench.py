#!/usr/bin/env python3
"""rt_multi_render (several GPUs, ONE process, C ABI only -- no torch) on 1..N devices: ms per frame end to end
(host buffers in and out) and the slowest device's kernel time; checks the image against the single-device frame.

    python tools/gpu_multi_bench.py [workload] [frames]       (run under gpurun --gpus N)
"""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene, ImageBuffer, MultiDeviceScene  # noqa: E402
from hslu_i.ba_raytracing.f2501_raytracer_b200 import _lib as L  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 20
flat, rc, features, desc = bench.build_workload(name, 0)
n_dev = L.load().rt_device_count()
single = DeviceScene(flat)
want, _ = single.render(rc)
single.close()
out = {"workload": desc, "devices_visible": n_dev, "runs": []}
base = None
for n in [k for k in (1, 2, 4, 8) if k <= n_dev]:
    t0 = time.perf_counter()
    multi = MultiDeviceScene(flat, n_devices=n)
    create_ms = 1e3 * (time.perf_counter() - t0)
    buf = ImageBuffer(rc.cfg.width, rc.cfg.height)
    for _ in range(3):
        multi.render(rc, buf)
    kernel = []
    t0 = time.perf_counter()
    for _ in range(frames):
        _, st = multi.render(rc, buf)
        kernel.append(st["kernel_ms"])
    e2e_ms = 1e3 * (time.perf_counter() - t0) / frames
    same = bool(np.array_equal(buf.buffer[want.hit_mask == 1], want.buffer[want.hit_mask == 1]) and np.array_equal(buf.hit_mask, want.hit_mask))
    info = multi.info()
    multi.close()
    base = base or e2e_ms
    run = {"devices": n, "peer_devices": info["peer_devices"], "e2e_ms_per_frame": e2e_ms, "slowest_device_kernel_ms": float(np.mean(kernel)),
           "efficiency_e2e": base / (n * e2e_ms), "create_ms": create_ms, "image_identical_to_single_device": same}
    out["runs"].append(run)
    print(json.dumps(run), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/multi_bench.json").write_text(json.dumps(out, indent=1))
