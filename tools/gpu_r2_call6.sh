#!/bin/bash
# round-2 call 6: chain-window probe; deferral A/B on one rank's eighth (help between units, wait backoff)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/gpu_window_probe.py > gpurun_out/r2c6_window_probe.log 2>&1
python tools/gpu_tiles_onegpu.py 8 2 > gpurun_out/r2c6_tiles_base.log 2>&1
for v in nohelp help32 nobackoff nohelp_nobackoff; do
  RT_LIB_PATH=$GRAFT_REPO_ROOT/build/librt_$v.so python tools/gpu_tiles_onegpu.py 8 2 > gpurun_out/r2c6_tiles_$v.log 2>&1
done
python tools/gpu_one_frame.py c3 4 > gpurun_out/r2c6_c3_base.log 2>&1
RT_LIB_PATH=$GRAFT_REPO_ROOT/build/librt_nobackoff.so python tools/gpu_one_frame.py c3 4 > gpurun_out/r2c6_c3_nobackoff.log 2>&1
cat gpurun_out/r2c6_window_probe.log; tail -n 1 gpurun_out/r2c6_tiles_*.log; tail -n 2 gpurun_out/r2c6_c3_*.log
