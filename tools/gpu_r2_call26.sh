#!/bin/bash
# round-2 final single-GPU pass: smoke, full GPU suite, bench (both arms), launch lists, ncu capture of the headline kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c26_smoke.log 2>&1; echo "smoke rc=$?"
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r2c26_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c26_pytest.log
tail -n 5 gpurun_out/r2c26_pytest.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c26_bench.json 2> gpurun_out/r2c26_bench.err
echo "bench rc=$?" >> gpurun_out/r2c26_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2c26_ref.json 2> gpurun_out/r2c26_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c26_launches.csv python bench.py --steps 3 --warmup 3 --scenes none --no-cpu-baseline > gpurun_out/r2c26_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c26_launches_c4.csv python tools/gpu_one_frame.py c4 3 > gpurun_out/r2c26_ncu_launches_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:direct_kernel -c 1 --launch-skip 1 -f -o gpurun_out/r2_direct_c3_final2 python tools/gpu_one_frame.py c3 3 > gpurun_out/r2c26_ncu_c3.log 2>&1
tail -n 3 gpurun_out/r2c26_bench.err; tail -n 2 gpurun_out/r2c26_ncu_c3.log; cut -c1-400 gpurun_out/r2c26_bench.json; cat gpurun_out/r2c26_ref.json | cut -c1-300
