#!/bin/bash
# round-2 final 8-GPU pass: rt_multi on real devices (tests + bench), the torchrun harness at N = 8 (with the per-scene table) / 4 / 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c28_gpus.txt
( time timeout 300 python -m pytest tests -m gpu -q -k "multi_gpu or row_tiles or place_tiles" ) > gpurun_out/r2c28_pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c28_pytest_multi.log
timeout 300 python tools/gpu_multi_bench.py c3 20 > gpurun_out/r2c28_multi_c3.log 2>&1
cp gpurun_out/multi_bench.json gpurun_out/r2c28_multi_c3.json
for n in 8 4 2; do
  sc="none"; [ $n = 8 ] && sc="c1,c4,c5"
  ( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29510+n)) bench.py --gpus $n --steps 20 --warmup 5 --scenes $sc --no-cpu-baseline ) > gpurun_out/r2c28_bench_n$n.json 2> gpurun_out/r2c28_bench_n$n.err
  echo "bench n=$n rc=$?" >> gpurun_out/r2c28_bench_n$n.err
done
tail -n 4 gpurun_out/r2c28_pytest_multi.log; cat gpurun_out/r2c28_multi_c3.log | tail -n 12; tail -n 2 gpurun_out/r2c28_bench_n*.err
