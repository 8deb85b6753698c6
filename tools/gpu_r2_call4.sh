#!/bin/bash
# round-2 call 4: full GPU suite on the wavefront / windowed-chain / device-builder tree, bench, A/B switches, ncu of C5 (reduced frame)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r2c4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c4_pytest.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
echo "bench rc=$?" >> gpurun_out/r2c4_bench.err
# A/B: wavefront vs megakernel
for w in c1 c4; do
  python tools/gpu_one_frame.py $w 3 > gpurun_out/r2c4_wave_$w.log 2>&1
  RT_WAVEFRONT=0 python tools/gpu_one_frame.py $w 3 > gpurun_out/r2c4_mega_$w.log 2>&1
done
# A/B: top BVH levels in shared memory (N2), L2 persistence window on the node array
python tools/gpu_one_frame.py c3 4 > gpurun_out/r2c4_ab_base_c3.log 2>&1
RT_NVCC_EXTRA="-DRT_SMEM_TOP_NODES=85" RT_LIB_PATH=$GRAFT_REPO_ROOT/build/librt_smem.so python tools/gpu_one_frame.py c3 4 > gpurun_out/r2c4_ab_smem_c3.log 2>&1
python tools/gpu_one_frame.py c5 2 810 675 > gpurun_out/r2c4_ab_base_c5.log 2>&1
RT_NVCC_EXTRA="-DRT_SMEM_TOP_NODES=85" RT_LIB_PATH=$GRAFT_REPO_ROOT/build/librt_smem.so python tools/gpu_one_frame.py c5 2 810 675 > gpurun_out/r2c4_ab_smem_c5.log 2>&1
RT_L2_PERSIST=1 python tools/gpu_one_frame.py c5 2 810 675 > gpurun_out/r2c4_ab_l2_c5.log 2>&1
# one rank's eighth of the headline frame: DEFER variant at 6 vs 7 CTAs/SM, and no deferral
python tools/gpu_tiles_onegpu.py 8 2 > gpurun_out/r2c4_tiles_defer6.log 2>&1
RT_NVCC_EXTRA="-DRT_DEFER_MIN_CTAS=7" RT_LIB_PATH=$GRAFT_REPO_ROOT/build/librt_defer7.so python tools/gpu_tiles_onegpu.py 8 2 > gpurun_out/r2c4_tiles_defer7.log 2>&1
RT_DEFER_BELOW=0 python tools/gpu_tiles_onegpu.py 8 2 > gpurun_out/r2c4_tiles_nodefer.log 2>&1
# ncu: C5's lighting kernel at a quarter of the frame in each axis (same 4.2 M-triangle scene)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:direct_kernel -c 1 -f -o gpurun_out/r2_direct_points_c5 python tools/gpu_one_frame.py c5 1 405 337 > gpurun_out/r2c4_ncu_c5.log 2>&1
tail -n 4 gpurun_out/r2c4_pytest.log; tail -n 3 gpurun_out/r2c4_bench.err; tail -n 2 gpurun_out/r2c4_ncu_c5.log
