#!/bin/bash
# Round 2, 2-GPU call: multi-rank path (NCCL in the captured step + stream overlap), N-rank == 1-rank test on NCCL
mkdir -p gpurun_out
O=gpurun_out
echo "== dist test (NCCL, 2 GPUs)"; timeout -s KILL 300 python -m pytest tests/test_gpu_dist.py -q -s -p no:cacheprovider 2>&1 | grep -E "bucket|DIST_CHECK|passed|failed" | cut -c1-200
echo "== bench C2 x2"; timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 2>$O/r2k2_bench.err | tail -1 > $O/r2k2_bench.json; cut -c1-500 $O/r2k2_bench.json; tail -3 $O/r2k2_bench.err
