#!/bin/bash
# Round 2, call H: stream overlap A/B, norm-pass micro-benchmark, grid_sample fma form, finalize rewrite
mkdir -p gpurun_out
O=gpurun_out
echo "== tests (ops + conv + model)"; timeout -s KILL 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_conv_tc.py tests/test_gpu_model.py tests/test_gpu_next_rows.py tests/test_gpu_fidelity.py -q -p no:cacheprovider > $O/r2h_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2h_tests.txt | cut -c1-300
for V in 1 0; do
echo "== bench C2 (--stream_overlap $V)"; timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --torch_gpu_reference 0 --stream_overlap $V 2>$O/r2h_bench$V.err | tail -1 > $O/r2h_bench$V.json; python - <<PY
import json
d=json.load(open('gpurun_out/r2h_bench$V.json'))
print({k:d.get(k) for k in ('value','ms_per_step','e2e','grid_sample','gpu_launches')})
PY
tail -2 $O/r2h_bench$V.err
done
echo "== nbench"; timeout -s KILL 600 python scripts/nbench.py --variants "" "NEMAR_LEAN_U=4" "NEMAR_LEAN_RED_U=4" "NEMAR_LEAN_CTAS_PER_SM=16 NEMAR_LEAN_APPLY_PER_SM=12 NEMAR_LEAN_RED_PER_SM=8" "NEMAR_LEAN_U=4 NEMAR_LEAN_RED_U=4 NEMAR_LEAN_CTAS_PER_SM=6 NEMAR_LEAN_APPLY_PER_SM=4 NEMAR_LEAN_RED_PER_SM=3" 2>&1 | tee $O/r2h_nbench.txt
