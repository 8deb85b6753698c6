#!/bin/bash
# round-2 call 7: fixed window protocol (probe), deferral tickets A/B on one rank's eighth, then the full suite and the bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/gpu_window_probe.py 8288 8289 9000 10000 16576 8288 65536 > gpurun_out/r2c7_window_probe.log 2>&1
python tools/gpu_tiles_onegpu.py 8 2 > gpurun_out/r2c7_tiles_help8.log 2>&1
for v in nohelp help32 help4; do
  RT_LIB_PATH=$GRAFT_REPO_ROOT/build/librt_$v.so python tools/gpu_tiles_onegpu.py 8 2 > gpurun_out/r2c7_tiles_$v.log 2>&1
done
RT_LIB_PATH=$GRAFT_REPO_ROOT/build/librt_nohelp.so python tools/gpu_tiles_onegpu.py 4 2 > gpurun_out/r2c7_tiles4_nohelp.log 2>&1
python tools/gpu_tiles_onegpu.py 4 2 > gpurun_out/r2c7_tiles4_help8.log 2>&1
cat gpurun_out/r2c7_window_probe.log; tail -n 1 gpurun_out/r2c7_tiles*.log
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/r2c7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c7_pytest.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err
echo "bench rc=$?" >> gpurun_out/r2c7_bench.err
tail -n 5 gpurun_out/r2c7_pytest.log; tail -n 3 gpurun_out/r2c7_bench.err
