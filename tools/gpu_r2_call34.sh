#!/bin/bash
# parity subset on the final build (BVH walks: closest + inlined shadow descend to a leaf first, outlined walks node-or-leaf)
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "linear_scan or c2_ or c4_ or wavefront or deferred or nan or beams_equal or more_lights" > gpurun_out/r2c34_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2c34_pytest.log
