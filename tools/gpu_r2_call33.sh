#!/bin/bash
# while-while walks per call site: A = closest + inlined shadow (product), B = inlined shadow only, C = closest only
mkdir -p gpurun_out
{
for lib in "" build/librt_wwB.so build/librt_wwC.so; do
  echo "== lib=${lib:-product (A)}"
  export RT_LIB_PATH=${lib:-hslu_i/ba_raytracing/f2501_raytracer_b200/librt_b200.so}
  timeout 200 python tools/gpu_probe.py . c4 40 2>&1 | tail -n 1
  timeout 200 python tools/gpu_probe.py . c2 40 2>&1 | tail -n 1
  timeout 200 python tools/gpu_probe.py . c3 14 2>&1 | tail -n 1
  timeout 200 python tools/gpu_tiles_onegpu.py 8 2 2>&1 | grep world | cut -c1-110
done
} > gpurun_out/r2c33_ab.txt 2>&1
cat gpurun_out/r2c33_ab.txt
