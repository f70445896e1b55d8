#!/bin/bash
# Round 2, call I: fused statistics (fixed), conflict-free norm reductions, stream overlap A/B
mkdir -p gpurun_out
O=gpurun_out
echo "== tests"; timeout -s KILL 1500 python -m pytest tests/ -m gpu -q -p no:cacheprovider > $O/r2i_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2i_tests.txt | cut -c1-300
echo "== nbench"; timeout -s KILL 300 python scripts/nbench.py --variants "" --shapes res256 res256r head64 stn32 d512 2>&1 | tee $O/r2i_nbench.txt
for V in "1 1" "0 1" "1 0"; do
set -- $V
echo "== bench C2 (--stream_overlap $1, NEMAR_FUSED_STATS=$2)"; NEMAR_FUSED_STATS=$2 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --torch_gpu_reference 0 --grid_sample_bench 0 --stream_overlap $1 2>$O/r2i_bench$1$2.err | tail -1 > $O/r2i_bench$1$2.json; python - <<PY
import json
d=json.load(open('gpurun_out/r2i_bench$1$2.json'))
print({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches')})
PY
tail -2 $O/r2i_bench$1$2.err
done
