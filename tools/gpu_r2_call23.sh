#!/bin/bash
# DEFER: which candidate count makes a unit a heavy one (RT_DEFER_CAND), on a rank's eighth / quarter of the C3 frame
mkdir -p gpurun_out
{
for c in 48 40 32 24 16 8 4; do
  echo "== RT_DEFER_CAND=$c"
  RT_DEFER_CAND=$c timeout 300 python tools/gpu_tiles_onegpu.py 8 8 2>&1 | grep world | cut -c1-140
  RT_DEFER_CAND=$c RT_DEFER_BELOW=100 timeout 300 python tools/gpu_tiles_onegpu.py 4 8 2>&1 | grep world | cut -c1-140
done
echo "== plain, world 4"
timeout 300 python tools/gpu_tiles_onegpu.py 4 8 2>&1 | grep world | cut -c1-140
} > gpurun_out/r2c23_defer_cand.txt 2>&1
cat gpurun_out/r2c23_defer_cand.txt
