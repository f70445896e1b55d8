#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 2 --top 60"
run() { n=$1; echo "== $n"; shift; env "$@" timeout 300 $B 2>/dev/null | tail -1 > gpurun_out/r9_$n.json; python -c "
import sys, json
r = json.load(open('gpurun_out/r9_$n.json')); k = (r.get('roofline') or {}).get('by_kernel', {})
print('ms/step', r['ms_per_step'], {a: round(b['ms'] / r['steps'], 2) for a, b in list(k.items())[:9]})"; }
run default X=1
run wide NEMAR_TC_WIDE=1

echo "== plain (no kernel timing)"; timeout 300 python bench.py --steps 20 --warmup 5 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 2>/dev/null | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.readline()); print(r['ms_per_step'], r['value'], r['e2e'])"
