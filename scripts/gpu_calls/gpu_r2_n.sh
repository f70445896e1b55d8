#!/bin/bash
# Round 2: where the in-step time goes on the final build — per-call CUDA-event breakdown (eager, single stream, L2 as
# the step leaves it) and a WARM ncu launch list (--cache-control none: one pass, no replay, no flush between kernels)
mkdir -p gpurun_out
O=gpurun_out
echo "== per-op breakdown"; NEMAR_WGRAD_STREAM=0 timeout -s KILL 400 python bench.py --cuda_graph 0 --stream_overlap 0 --steps 10 --warmup 3 --kernel_timing 2 --top 40 --no_cpu_baseline --grid_sample_bench 0 --torch_gpu_reference 0 2>$O/r2n_perop.err | tail -1 > $O/r2n_perop.json; cut -c1-300 $O/r2n_perop.json; tail -2 $O/r2n_perop.err
B="python bench.py --profile --cuda_graph 0 --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --torch_gpu_reference 0 --stream_overlap 0"
echo "== warm ncu launch list"; NEMAR_WGRAD_STREAM=0 timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file $O/r2n_launches_warm.csv $B > $O/r2n_ncu_list.log 2>&1; echo rc=$?
du -sh $O
