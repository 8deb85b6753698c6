#!/bin/bash
# round-2 call 2: the full GPU test suite on the new code + the default bench (with the per-scene table) + reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nproc > gpurun_out/r2c2_nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
echo "bench rc=$?" >> gpurun_out/r2c2_bench.err
( time python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r2c2_ref.json 2> gpurun_out/r2c2_ref.err
tail -n 5 gpurun_out/r2c2_pytest.log; tail -n 3 gpurun_out/r2c2_bench.err
