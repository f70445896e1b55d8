#!/bin/bash
# Round 2, call J: conflict-free norm reductions
mkdir -p gpurun_out
O=gpurun_out
echo "== nbench"; timeout -s KILL 300 python scripts/nbench.py --variants "" "NEMAR_LEAN_RED_U=4" --shapes res256 res256r head64 stn32 d512 2>&1 | tee $O/r2j_nbench.txt
echo "== tests"; timeout -s KILL 1500 python -m pytest tests/ -m gpu -q -p no:cacheprovider > $O/r2j_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2j_tests.txt | cut -c1-300
echo "== bench C2"; timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --torch_gpu_reference 0 --grid_sample_bench 0 2>$O/r2j_bench.err | tail -1 > $O/r2j_bench.json; python - <<PY
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches')})
PY
tail -2 $O/r2j_bench.err
echo "== launch list"
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2j_launches.csv python bench.py --profile --cuda_graph 0 --steps 2 --warmup 2 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --torch_gpu_reference 0 --stream_overlap 0 > $O/r2j_ncu_list.log 2>&1; echo rc=$?
