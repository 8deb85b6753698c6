#!/bin/bash
mkdir -p gpurun_out
{ timeout 300 python tools/gpu_tiles_onegpu.py 8 2 4 8 2>&1 | grep world | cut -c1-150; timeout 300 python tools/gpu_tiles_onegpu.py 4 2 8 2>&1 | grep world | cut -c1-150; timeout 300 python tools/gpu_tiles_onegpu.py 2 2 8 2>&1 | grep world | cut -c1-150; } > gpurun_out/r2c27_tiles.txt 2>&1
cat gpurun_out/r2c27_tiles.txt
