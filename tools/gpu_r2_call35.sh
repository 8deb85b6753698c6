#!/bin/bash
# large trees (device-built and host-built) with the final BVH walks
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_bvh_builder or large_synthetic" > gpurun_out/r2c35_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2c35_pytest.log
