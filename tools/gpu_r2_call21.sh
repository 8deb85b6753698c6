#!/bin/bash
mkdir -p gpurun_out
{
for v in "RT_SMALL_LAUNCH=plain" "RT_SMALL_LAUNCH=defer"; do
echo "== $v"
env $v RT_LIB_PATH=build/librt_dbg.so timeout 600 python tools/gpu_dbg_tail.py 2>&1 | tail -n 45
done
} > gpurun_out/r2c21_dbg.txt 2>&1
cat gpurun_out/r2c21_dbg.txt
