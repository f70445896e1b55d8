#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --profile --steps 1 --warmup 0 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gather_kernel -s 10 -c 3 -o gpurun_out/prof_gather2 -f $B > gpurun_out/ncu_g2.log 2>&1
echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -s 30 -c 3 -o gpurun_out/prof_wgrad2 -f $B > gpurun_out/ncu_w2.log 2>&1
echo "rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --size 512 --batch 8 --multi_resolution 3 --lambda_smooth 200 --alpha 1.0 > gpurun_out/bench_c4shape.json 2> gpurun_out/bench_c4shape.err
cut -c1-400 gpurun_out/bench_c4shape.json; tail -2 gpurun_out/bench_c4shape.err | cut -c1-300
