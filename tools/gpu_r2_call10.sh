#!/bin/bash
# round-2 call 10 (8 GPUs): rt_multi on real devices (tests + bench), the torchrun harness at N = 8 / 4 / 2 with the per-scene table at N = 8
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c10_gpus.txt
( time timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or row_tiles or place_tiles" ) > gpurun_out/r2c10_pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c10_pytest_multi.log
timeout 600 python tools/gpu_multi_bench.py c3 20 > gpurun_out/r2c10_multi_c3.log 2>&1
cp gpurun_out/multi_bench.json gpurun_out/r2c10_multi_c3.json
timeout 600 python tools/gpu_multi_bench.py c1 3 > gpurun_out/r2c10_multi_c1.log 2>&1
cp gpurun_out/multi_bench.json gpurun_out/r2c10_multi_c1.json
for n in 8 4 2; do
  sc="none"; [ $n = 8 ] && sc="c1,c2,c4,c5"
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29510+n)) bench.py --gpus $n --steps 20 --warmup 5 --scenes $sc ) > gpurun_out/r2c10_bench_n$n.json 2> gpurun_out/r2c10_bench_n$n.err
  echo "bench n=$n rc=$?" >> gpurun_out/r2c10_bench_n$n.err
done
( time python bench.py --steps 20 --warmup 5 --scenes none --no-cpu-baseline ) > gpurun_out/r2c10_bench_n1.json 2> gpurun_out/r2c10_bench_n1.err
tail -n 4 gpurun_out/r2c10_pytest_multi.log; cat gpurun_out/r2c10_multi_c3.log; tail -n 2 gpurun_out/r2c10_bench_n*.err
