#!/bin/bash
# round-2 call 11: determinism probe (C1), cooperative overflow lanes: tests + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/gpu_determinism.py c1 > gpurun_out/r2c11_determinism_c1.log 2>&1
cat gpurun_out/r2c11_determinism_c1.log
python tools/gpu_one_frame.py c3 4 > gpurun_out/r2c11_c3.log 2>&1
python tools/gpu_one_frame.py c5 2 > gpurun_out/r2c11_c5.log 2>&1
python tools/gpu_one_frame.py c1 3 > gpurun_out/r2c11_c1.log 2>&1
python tools/gpu_tiles_onegpu.py 8 2 > gpurun_out/r2c11_tiles8_defer.log 2>&1
RT_DEFER_BELOW=0 python tools/gpu_tiles_onegpu.py 8 2 > gpurun_out/r2c11_tiles8_nodefer.log 2>&1
RT_DEFER_BELOW=0 python tools/gpu_tiles_onegpu.py 4 2 > gpurun_out/r2c11_tiles4_nodefer.log 2>&1
tail -n 2 gpurun_out/r2c11_c*.log; tail -n 1 gpurun_out/r2c11_tiles*.log
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/r2c11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c11_pytest.log
tail -n 5 gpurun_out/r2c11_pytest.log
