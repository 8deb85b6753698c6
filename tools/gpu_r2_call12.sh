#!/bin/bash
# round-2 call 12: after the padding-round fix: determinism probe, full suite, bench, final ncu captures of the headline kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/gpu_determinism.py c1 > gpurun_out/r2c12_determinism_c1.log 2>&1
cat gpurun_out/r2c12_determinism_c1.log
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r2c12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c12_pytest.log
tail -n 5 gpurun_out/r2c12_pytest.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err
echo "bench rc=$?" >> gpurun_out/r2c12_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c12_launches.csv python bench.py --steps 3 --warmup 3 --scenes none --no-cpu-baseline > gpurun_out/r2c12_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:direct_kernel -c 1 --launch-skip 1 -f -o gpurun_out/r2_direct_c3_final python tools/gpu_one_frame.py c3 3 > gpurun_out/r2c12_ncu_c3.log 2>&1
tail -n 3 gpurun_out/r2c12_bench.err; tail -n 2 gpurun_out/r2c12_ncu_c3.log
