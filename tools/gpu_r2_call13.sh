#!/bin/bash
# Timing bisect of the headline kernel over the round-2 commits (worktrees built under build/bisect/, scratch).
mkdir -p gpurun_out
{
for c in 477ce8c 52ebcf1 cfdfcfc 5a98f94 658ab44 d14ddc8 757ea8b; do
  timeout 300 python tools/gpu_probe.py build/bisect/$c c3 28 2>&1 | tail -n 1
done
timeout 300 python tools/gpu_probe.py . c3 28 2>&1 | tail -n 1
RT_CHAIN_WINDOW=300000 timeout 300 python tools/gpu_probe.py . c3 28 2>&1 | tail -n 1
RT_CHAIN_WINDOW=16384 timeout 300 python tools/gpu_probe.py . c3 28 2>&1 | tail -n 1
for c in 477ce8c 5a98f94 757ea8b; do
  timeout 300 python tools/gpu_probe.py build/bisect/$c c4 60 2>&1 | tail -n 1
done
timeout 300 python tools/gpu_probe.py . c4 60 2>&1 | tail -n 1
} > gpurun_out/r2c13_bisect.txt 2>&1
cat gpurun_out/r2c13_bisect.txt
