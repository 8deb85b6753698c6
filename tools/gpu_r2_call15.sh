#!/bin/bash
# ncu capture of the headline kernel after RT_FACING_FIRST, plus a bench line
mkdir -p gpurun_out
( time python bench.py --steps 20 --warmup 5 --scenes none ) > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:direct_kernel -c 1 --launch-skip 1 -f -o gpurun_out/r2_direct_c3_facing python tools/gpu_one_frame.py c3 3 > gpurun_out/r2c15_ncu_c3.log 2>&1
tail -n 2 gpurun_out/r2c15_ncu_c3.log; cat gpurun_out/r2c15_bench.json | cut -c1-600
