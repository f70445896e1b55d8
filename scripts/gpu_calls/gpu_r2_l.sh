#!/bin/bash
# Round 2, call L: weight gradients on their own stream (A/B), full test suite
mkdir -p gpurun_out
O=gpurun_out
echo "== tests"; timeout -s KILL 1500 python -m pytest tests/ -m gpu -q -p no:cacheprovider > $O/r2l_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2l_tests.txt | cut -c1-300
for V in 1 0; do
echo "== bench C2 (NEMAR_WGRAD_STREAM=$V)"; NEMAR_WGRAD_STREAM=$V timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --torch_gpu_reference 0 --grid_sample_bench 0 2>$O/r2l_bench$V.err | tail -1 > $O/r2l_bench$V.json; python - <<PY
import json
d=json.load(open('gpurun_out/r2l_bench$V.json'))
print({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches')})
PY
tail -2 $O/r2l_bench$V.err
done
