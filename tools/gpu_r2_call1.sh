#!/bin/bash
# round-2 call 1: A/B of the three beam switches (product build vs -DRT_SPHERE_MISS_CULL -DRT_SELF_CULL -DRT_EMPTY_LIST_SKIP)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export RT_NVCC_EXTRA_NEXT="-DRT_SPHERE_MISS_CULL=1 -DRT_SELF_CULL=1 -DRT_EMPTY_LIST_SKIP=1"
nvidia-smi -L > gpurun_out/r2c1_gpu.txt
python tools/gpu_beam_ab.py c3 c1 > gpurun_out/r2c1_ab_product.log 2>&1
RT_NVCC_EXTRA="$RT_NVCC_EXTRA_NEXT" RT_LIB_PATH=$GRAFT_REPO_ROOT/build/librt_next.so python tools/gpu_beam_ab.py c3 c1 > gpurun_out/r2c1_ab_next.log 2>&1
python tools/gpu_perf.py c3 c4 c2 c1 > gpurun_out/r2c1_perf_product.log 2>&1
RT_NVCC_EXTRA="$RT_NVCC_EXTRA_NEXT" RT_LIB_PATH=$GRAFT_REPO_ROOT/build/librt_next.so python tools/gpu_perf.py c3 c4 c2 c1 > gpurun_out/r2c1_perf_next.log 2>&1
RT_NVCC_EXTRA="$RT_NVCC_EXTRA_NEXT" RT_LIB_PATH=$GRAFT_REPO_ROOT/build/librt_next.so timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest_next.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c1_pytest_next.log
tail -3 gpurun_out/r2c1_*.log
