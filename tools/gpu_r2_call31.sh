#!/bin/bash
# ncu captures of the wavefront kernels on C4 (bounce config): lighting of the point queue, path walk
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:direct_kernel -c 1 --launch-skip 1 -f -o gpurun_out/r2_direct_points_c4 python tools/gpu_one_frame.py c4 3 > gpurun_out/r2c31_ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:paths_kernel -c 1 --launch-skip 1 -f -o gpurun_out/r2_paths_c4 python tools/gpu_one_frame.py c4 3 > gpurun_out/r2c31_ncu_b.log 2>&1
tail -n 2 gpurun_out/r2c31_ncu_a.log gpurun_out/r2c31_ncu_b.log
