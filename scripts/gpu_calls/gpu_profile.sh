#!/bin/bash
# ncu passes (1 GPU): launch list of one bench step + one full capture of the dominant conv kernel.
mkdir -p gpurun_out
B="python bench.py --profile --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gather_kernel -s 60 -c 2 -o gpurun_out/prof_gather -f $B > gpurun_out/ncu_full.log 2>&1
echo "full gather rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -s 20 -c 2 -o gpurun_out/prof_wgrad -f $B > gpurun_out/ncu_full2.log 2>&1
echo "full wgrad rc=$?"; ls -la gpurun_out/*.ncu-rep
