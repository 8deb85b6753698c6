#!/bin/bash
# DEFER with all lights in a unit: tests of the deferral paths, determinism, a rank's eighth / quarter of C3, full frames
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "defer or chained or more_lights or asynchronous or multi_gpu or tiles" > gpurun_out/r2c24_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2c24_pytest.log
{
for v in "RT_LIGHTS_PER_UNIT=1" "RT_LIGHTS_PER_UNIT=5"; do
  echo "== $v"
  env $v timeout 300 python tools/gpu_tiles_onegpu.py 8 8 2>&1 | grep world | cut -c1-140
  env $v RT_DEFER_BELOW=100 timeout 300 python tools/gpu_tiles_onegpu.py 4 8 2>&1 | grep world | cut -c1-140
done
echo "== plain world 4"; timeout 300 python tools/gpu_tiles_onegpu.py 4 8 2>&1 | grep world | cut -c1-140
for c in 16 24 32; do echo "== RT_DEFER_CAND=$c"; RT_DEFER_CAND=$c timeout 300 python tools/gpu_tiles_onegpu.py 8 8 2>&1 | grep world | cut -c1-140; done
timeout 300 python tools/gpu_probe.py . c3 24 2>&1 | tail -n 1
RT_DEFER_BELOW=1000000 timeout 300 python tools/gpu_probe.py . c3 24 2>&1 | tail -n 1
timeout 600 python tools/gpu_determinism.py c3 2>&1 | grep -v "rep 1" | cut -c1-200
} > gpurun_out/r2c24_defer.txt 2>&1
cat gpurun_out/r2c24_defer.txt
