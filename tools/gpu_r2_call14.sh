#!/bin/bash
# A/B: facing test before the shadow query (RT_FACING_FIRST) and register bounds of direct_kernel; then the GPU suite.
mkdir -p gpurun_out
{
for lib in build/librt_nofacing.so "" build/librt_ctas6.so build/librt_ctas8.so build/librt_ctas5.so; do
  echo "lib=${lib:-product}"
  RT_LIB_PATH=${lib:-hslu_i/ba_raytracing/f2501_raytracer_b200/librt_b200.so} timeout 300 python tools/gpu_probe.py . c3 28 2>&1 | tail -n 1
done
for w in c1 c2 c4 c5; do
  for lib in build/librt_nofacing.so ""; do
    echo "lib=${lib:-product}"
    RT_LIB_PATH=${lib:-hslu_i/ba_raytracing/f2501_raytracer_b200/librt_b200.so} timeout 600 python tools/gpu_probe.py . $w 8 2>&1 | tail -n 1
  done
done
} > gpurun_out/r2c14_ab.txt 2>&1
cat gpurun_out/r2c14_ab.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c14_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2c14_pytest.log
