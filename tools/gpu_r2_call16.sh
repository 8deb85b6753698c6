#!/bin/bash
# A/B: lights per unit (1 = chained per-light units as before, 5 = all lights of the frame in one unit), all configs
mkdir -p gpurun_out
{
for lpu in 1 2 3 5; do
  echo "RT_LIGHTS_PER_UNIT=$lpu"
  RT_LIGHTS_PER_UNIT=$lpu timeout 300 python tools/gpu_probe.py . c3 28 2>&1 | tail -n 1
done
for w in c1 c4 c5; do
  for lpu in 1 99; do
    echo "RT_LIGHTS_PER_UNIT=$lpu"
    RT_LIGHTS_PER_UNIT=$lpu timeout 600 python tools/gpu_probe.py . $w 8 2>&1 | tail -n 1
  done
done
} > gpurun_out/r2c16_ab.txt 2>&1
cat gpurun_out/r2c16_ab.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c16_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2c16_pytest.log
