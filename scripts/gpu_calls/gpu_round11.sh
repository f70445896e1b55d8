#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 2 --top 60"
run() { n=$1; echo "== $n"; shift; env "$@" timeout 300 $B 2>/dev/null | tail -1 > gpurun_out/r11_$n.json; python -c "
import sys, json
r = json.load(open('gpurun_out/r11_$n.json')); k = (r.get('roofline') or {}).get('by_kernel', {})
print('ms/step', r['ms_per_step'], {a: round(b['ms'] / r['steps'], 2) for a, b in list(k.items())[:7]})"; }
run r4s8 X=1
run r2s8 NEMAR_LEAN_RED_PER_SM=2
run r3s8 NEMAR_LEAN_RED_PER_SM=3
run r1s8 NEMAR_LEAN_RED_PER_SM=1
run r4u4 NEMAR_LEAN_RED_U=4
run r2u4 NEMAR_LEAN_RED_U=4 NEMAR_LEAN_RED_PER_SM=2
run r4s6 NEMAR_LEAN_CTAS_PER_SM=6
run r4s12 NEMAR_LEAN_CTAS_PER_SM=12
