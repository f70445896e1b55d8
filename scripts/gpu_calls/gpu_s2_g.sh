#!/bin/bash
# 2 GPUs: the captured step with its two NCCL all-reduces inside the graph; N-rank == 1-rank equivalence check.
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout -s KILL 300 $T bench.py --gpus 2 --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 > $O/g_bench2.json 2> $O/g_bench2.err; echo "graph bench rc=$?"; tail -1 $O/g_bench2.json | cut -c1-700; grep -i "capture\|error\|nccl w" $O/g_bench2.err | head -5
timeout -s KILL 300 $T bench.py --gpus 2 --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --cuda_graph 0 > $O/g_bench2_eager.json 2> $O/g_bench2_eager.err; echo "eager bench rc=$?"; tail -1 $O/g_bench2_eager.json | cut -c1-300
timeout -s KILL 300 $T tests/dist_check.py 2>&1 | tail -4
