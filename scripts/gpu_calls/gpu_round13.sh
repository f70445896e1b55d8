#!/bin/bash
mkdir -p gpurun_out
echo "== ops+model tests"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 2 --top 60"
run() { n=$1; echo "== $n"; shift; env "$@" timeout 300 $B 2>/dev/null | tail -1 > gpurun_out/r13_$n.json; python -c "
import sys, json
r = json.load(open('gpurun_out/r13_$n.json')); k = (r.get('roofline') or {}).get('by_kernel', {})
print('ms/step', r['ms_per_step'], {a: round(b['ms'] / r['steps'], 2) for a, b in list(k.items())[:7]})"; }
run default X=1
run apply4 NEMAR_LEAN_APPLY_PER_SM=4
run fwd6 NEMAR_LEAN_CTAS_PER_SM=6
