#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --profile --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^(bwd_apply_kernel|reduce_kernel|fwd_kernel)$" -s 560 -c 14 -o gpurun_out/prof_lean -f $B > gpurun_out/ncu_lean.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_lean.log
