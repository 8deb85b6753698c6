#!/bin/bash
# A/B: sums only touched by contributing pairs; 6 CTAs/SM; candidate capacity of a beam 32 / 48 / 64 / 96
mkdir -p gpurun_out
{
for lib in build/librt_base2.so "" build/librt_ctas6.so build/librt_cand32.so build/librt_cand64.so build/librt_cand96.so; do
  echo "== lib=${lib:-product}"
  export RT_LIB_PATH=${lib:-hslu_i/ba_raytracing/f2501_raytracer_b200/librt_b200.so}
  timeout 300 python tools/gpu_probe.py . c3 24 2>&1 | tail -n 1
  timeout 300 python tools/gpu_probe.py . c1 6 2>&1 | tail -n 1
  timeout 300 python tools/gpu_probe.py . c5 5 2>&1 | tail -n 1
  timeout 300 python tools/gpu_tiles_onegpu.py 8 8 2>&1 | grep world | cut -c1-140
done
} > gpurun_out/r2c22_ab.txt 2>&1
cat gpurun_out/r2c22_ab.txt
