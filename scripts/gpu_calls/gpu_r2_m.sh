#!/bin/bash
# Round 2, call M: software-pipelined norm passes
mkdir -p gpurun_out
O=gpurun_out
echo "== nbench"; timeout -s KILL 300 python scripts/nbench.py --variants "" "NEMAR_LEAN_APPLY_PER_SM=4" "NEMAR_LEAN_APPLY_PER_SM=2 NEMAR_LEAN_CTAS_PER_SM=6 NEMAR_LEAN_RED_PER_SM=3" --shapes res256 res256r head64 stn32 2>&1 | tee $O/r2m_nbench.txt
echo "== tests"; timeout -s KILL 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_zz_golden_sizes.py -q -p no:cacheprovider > $O/r2m_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2m_tests.txt | cut -c1-300
echo "== bench C2"; timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --torch_gpu_reference 0 --grid_sample_bench 0 2>$O/r2m_bench.err | tail -1 > $O/r2m_bench.json; python - <<PY
import json
d=json.load(open('gpurun_out/r2m_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches')})
PY
tail -2 $O/r2m_bench.err
