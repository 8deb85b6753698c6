#!/bin/bash
# Is the per-rank spread of the 8-way tiled C3 frame aliasing between the tile period (world x tile_rows rows) and the image content?
mkdir -p gpurun_out
{
for v in "RT_SMALL_LAUNCH=defer" "RT_SMALL_LAUNCH=plain"; do
  echo "== $v"
  env $v timeout 600 python tools/gpu_tiles_onegpu.py 8 4 8 12 16 20 28 40 2>&1 | grep world | cut -c1-140
done
} > gpurun_out/r2c19_tiles.txt 2>&1
cat gpurun_out/r2c19_tiles.txt
