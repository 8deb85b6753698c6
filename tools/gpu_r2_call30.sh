#!/bin/bash
# final bench line of the round (all scenes), after the single-pass gather
mkdir -p gpurun_out
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c30_bench.json 2> gpurun_out/r2c30_bench.err
echo "bench rc=$?" >> gpurun_out/r2c30_bench.err
tail -n 4 gpurun_out/r2c30_bench.err; cut -c1-300 gpurun_out/r2c30_bench.json
