#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --profile --steps 1 --warmup 0 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:norm_act_bwd_apply|plane_reduce|norm_act_fwd|gather_taps|sum_taps" -s 150 -c 14 -o gpurun_out/prof_elem -f $B > gpurun_out/ncu_elem.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_elem.ncu-rep
