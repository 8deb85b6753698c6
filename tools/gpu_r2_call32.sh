#!/bin/bash
# A/B: BVH walks that descend to a leaf before testing primitives (RT_WHILE_WHILE) against the node-or-leaf loop
mkdir -p gpurun_out
{
for lib in build/librt_ifif.so ""; do
  echo "== lib=${lib:-product (while-while)}"
  export RT_LIB_PATH=${lib:-hslu_i/ba_raytracing/f2501_raytracer_b200/librt_b200.so}
  timeout 200 python tools/gpu_probe.py . c4 50 2>&1 | tail -n 1
  timeout 200 python tools/gpu_probe.py . c2 50 2>&1 | tail -n 1
  timeout 200 python tools/gpu_probe.py . c3 16 2>&1 | tail -n 1
  timeout 200 python tools/gpu_probe.py . c1 5 2>&1 | tail -n 1
  timeout 200 python tools/gpu_probe.py . c5 4 2>&1 | tail -n 1
  timeout 200 python tools/gpu_tiles_onegpu.py 8 2 2>&1 | grep world | cut -c1-140
done
} > gpurun_out/r2c32_ab.txt 2>&1
cat gpurun_out/r2c32_ab.txt
unset RT_LIB_PATH
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "linear_scan or c2_ or c4_ or wavefront or deferred or nan" > gpurun_out/r2c32_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2c32_pytest.log
