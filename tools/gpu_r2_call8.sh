#!/bin/bash
# round-2 call 8: light deferral + cascade finishing: one rank's share (8-way, 4-way, 2-way), then the full suite and the bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/gpu_tiles_onegpu.py 8 2 > gpurun_out/r2c8_tiles8.log 2>&1
python tools/gpu_tiles_onegpu.py 4 2 > gpurun_out/r2c8_tiles4.log 2>&1
python tools/gpu_tiles_onegpu.py 2 2 > gpurun_out/r2c8_tiles2.log 2>&1
tail -n 1 gpurun_out/r2c8_tiles*.log
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/r2c8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c8_pytest.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err
echo "bench rc=$?" >> gpurun_out/r2c8_bench.err
tail -n 6 gpurun_out/r2c8_pytest.log; tail -n 3 gpurun_out/r2c8_bench.err
