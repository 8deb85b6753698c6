#!/bin/bash
# round-2 call 5: full GPU suite after the window / early-help / material changes; bench; chain-window sweep; one rank's eighth
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/r2c5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c5_pytest.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
echo "bench rc=$?" >> gpurun_out/r2c5_bench.err
for w in 16384 32768 65536 131072 1000000; do
  echo "window $w" >> gpurun_out/r2c5_window.log
  RT_CHAIN_WINDOW=$w python tools/gpu_one_frame.py c3 5 2>&1 | tail -n 3 >> gpurun_out/r2c5_window.log
done
python tools/gpu_tiles_onegpu.py 8 2 > gpurun_out/r2c5_tiles_defer.log 2>&1
python tools/gpu_tiles_onegpu.py 4 2 > gpurun_out/r2c5_tiles4_defer.log 2>&1
tail -n 4 gpurun_out/r2c5_pytest.log; tail -n 3 gpurun_out/r2c5_bench.err; cat gpurun_out/r2c5_window.log gpurun_out/r2c5_tiles_defer.log
