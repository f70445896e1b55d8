#!/bin/bash
mkdir -p gpurun_out
echo "== tc FUSED_STATS TPC=4"; NEMAR_FUSED_STATS=1 NEMAR_TC_TPC=4 timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 2 --top 60"
run() { n=$1; echo "== $n"; shift; env "$@" timeout 300 $B 2>/dev/null | tail -1 > gpurun_out/r7_$n.json; python -c "
import sys, json
r = json.load(open('gpurun_out/r7_$n.json')); k = (r.get('roofline') or {}).get('by_kernel', {})
print('ms/step', r['ms_per_step'], {a: round(b['ms'] / r['steps'], 2) for a, b in list(k.items())[:9]})"; }
run default X=1
run fused_stats NEMAR_FUSED_STATS=1
run k128 NEMAR_TC_TPC_K=128
run k256 NEMAR_TC_TPC_K=256
run occ2 NEMAR_WG_OCC=2
run occ1 NEMAR_WG_OCC=1
