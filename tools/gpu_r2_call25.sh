#!/bin/bash
# DEFER variant: resident CTAs per SM (register budget) on a rank's eighth of C3
mkdir -p gpurun_out
{
for lib in "" build/librt_d7.so build/librt_d5.so; do
  echo "== lib=${lib:-product (6 CTAs/SM)}"
  export RT_LIB_PATH=${lib:-hslu_i/ba_raytracing/f2501_raytracer_b200/librt_b200.so}
  timeout 300 python tools/gpu_tiles_onegpu.py 8 8 2>&1 | grep world | cut -c1-140
  RT_DEFER_BELOW=1000000 timeout 300 python tools/gpu_probe.py . c3 16 2>&1 | tail -n 1
done
} > gpurun_out/r2c25_defer_ctas.txt 2>&1
cat gpurun_out/r2c25_defer_ctas.txt
