#!/bin/bash
# Round 2, 8-GPU call: C3 (= C2 shape, batch 128 over 8 GPUs), C4 and C5 at 8 GPUs
mkdir -p gpurun_out
O=gpurun_out
run() {  # name, extra flags
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --steps 10 --warmup 3 $2 2>$O/r2k8_$1.err | tail -1 > $O/r2k8_$1.json
  echo "== $1"; cut -c1-420 $O/r2k8_$1.json; tail -2 $O/r2k8_$1.err | cut -c1-300
}
run c3 "--workload C2" 29621
run c4 "--workload C4" 29622
run c5 "--workload C5" 29623
