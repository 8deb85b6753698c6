#!/usr/bin/env python3
"""gpu_probe.py ROOT WORKLOAD FRAMES: renders FRAMES frames of a bench workload with the library of the checkout at ROOT
(`.` = this tree, or a worktree of an older commit built in place; RT_LIB_PATH selects another library) and prints the
minimum / median of the frames' summed kernel times.  Run under gpurun (A/B runs, timing bisects)."""
import sys, statistics
from pathlib import Path
root = Path(sys.argv[1]).resolve()
sys.path.insert(0, str(root))
import bench
from hslu_i.ba_raytracing.f2501_raytracer_b200 import DeviceScene
name = sys.argv[2]
flat, rc, features, desc = bench.build_workload(name, 0)
dev = DeviceScene(flat)
ks = []
for i in range(int(sys.argv[3])):
    _, st = dev.render(rc)
    ks.append(st['kernel_ms'])
ks = ks[3:]
print(root.name, name, "min %.3f med %.3f" % (min(ks), statistics.median(ks)), flush=True)
dev.close()
