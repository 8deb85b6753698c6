#!/bin/bash
# round-2 call 3: full GPU suite, default bench with the per-scene table, ncu captures (C3 direct_kernel, C5 render_kernel)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/r2c3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c3_pytest.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
echo "bench rc=$?" >> gpurun_out/r2c3_bench.err
RT_BVH_VERBOSE=1 python tools/gpu_one_frame.py c5 1 > gpurun_out/r2c3_c5_frame.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c3_launches.csv python bench.py --steps 3 --warmup 3 --scenes none --no-cpu-baseline > gpurun_out/r2c3_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:direct_kernel -c 1 --launch-skip 1 -f -o gpurun_out/r2_direct_c3 python tools/gpu_one_frame.py c3 3 > gpurun_out/r2c3_ncu_c3.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:render_kernel -c 1 -f -o gpurun_out/r2_render_c5 python tools/gpu_one_frame.py c5 1 > gpurun_out/r2c3_ncu_c5.log 2>&1
tail -n 6 gpurun_out/r2c3_pytest.log; tail -n 3 gpurun_out/r2c3_bench.err; tail -n 3 gpurun_out/r2c3_ncu_c5.log
