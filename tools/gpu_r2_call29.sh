#!/bin/bash
# forward-linked point lists (single-pass gather): bounce-path tests, C4 / C1 / C5 timing, C4 launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wavefront or ragged or bounce or c1 or c4 or extreme or packet or progressive" > gpurun_out/r2c29_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2c29_pytest.log
{
timeout 300 python tools/gpu_probe.py . c4 60 2>&1 | tail -n 1
timeout 300 python tools/gpu_probe.py . c1 6 2>&1 | tail -n 1
timeout 300 python tools/gpu_probe.py . c5 4 2>&1 | tail -n 1
} > gpurun_out/r2c29_probe.txt 2>&1
cat gpurun_out/r2c29_probe.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c29_launches_c4.csv python tools/gpu_one_frame.py c4 3 > gpurun_out/r2c29_ncu_launches_c4.log 2>&1
timeout 300 python tools/gpu_determinism.py c4 2>&1 | grep -v "rep 1" | cut -c1-150 | head -6
