#!/bin/bash
# Small launches (a rank's share of an N-way tiled C3 frame, C4): deferring per-light units vs all lights in a unit
mkdir -p gpurun_out
{
for v in "" "RT_DEFER_BELOW=0" "RT_DEFER_BELOW=0 RT_LIGHTS_PER_UNIT=1" "RT_DEFER_BELOW=0 RT_LIGHTS_PER_UNIT=2"; do
  echo "== ${v:-default}"
  env $v timeout 300 python tools/gpu_tiles_onegpu.py 8 8 2>&1 | tail -n 1
  env $v timeout 300 python tools/gpu_tiles_onegpu.py 4 8 2>&1 | tail -n 1
  env $v timeout 300 python tools/gpu_tiles_onegpu.py 2 8 2>&1 | tail -n 1
  env $v timeout 300 python tools/gpu_probe.py . c4 60 2>&1 | tail -n 1
  env $v timeout 300 python tools/gpu_probe.py . c2 60 2>&1 | tail -n 1
done
} > gpurun_out/r2c17_small.txt 2>&1
cat gpurun_out/r2c17_small.txt
