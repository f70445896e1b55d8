#!/bin/bash
# Round 2: weight-gradient residency / split sweep, measured on the step (the wgrad stream shares the SMs with the main stream)
mkdir -p gpurun_out
O=gpurun_out
echo "== kbench wgrad"; timeout -s KILL 300 python scripts/kbench.py --layers resblock d512 down2 --variants "" "NEMAR_WG_PAIR_OCC=1 NEMAR_WG_PAIR_STAGES=6" "NEMAR_WG_PAIR_OCC=1 NEMAR_WG_PAIR_STAGES=4" "NEMAR_WG_OCC_MAX=2" "NEMAR_WG_OCC_MAX=1" > $O/r2r_kbench.txt 2>&1; cut -c1-220 $O/r2r_kbench.txt
B="python bench.py --gpus 1 --steps 20 --warmup 5 --no_cpu_baseline --grid_sample_bench 0 --torch_gpu_reference 0 --kernel_timing 0"
run() { echo "== bench $1"; env $1 timeout -s KILL 300 $B 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"; }
run "NEMAR_WG_PAIR_OCC=1 NEMAR_WG_PAIR_STAGES=6"
run "NEMAR_WG_PAIR_OCC=1 NEMAR_WG_PAIR_STAGES=4"
run "NEMAR_WG_PAIR_OCC=1 NEMAR_WG_PAIR_STAGES=3"
run "NEMAR_WG_PAIR_OCC=1 NEMAR_WG_PAIR_STAGES=6 NEMAR_WG_OCC_MAX=2"
run "NEMAR_WG_PAIR_OCC=1 NEMAR_WG_PAIR_STAGES=6 NEMAR_WG_OCC_MAX=1"
run "NEMAR_WG_PAIR_OCC=1 NEMAR_WG_PAIR_STAGES=6 NEMAR_WGRAD_STREAM=0"
run "NEMAR_TC_PAIR=0"
