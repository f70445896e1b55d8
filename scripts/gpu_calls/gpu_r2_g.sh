#!/bin/bash
# Round 2, call G: k7 head/tail as column-tap gather + 7x1 conv, tiled grid_sample, rn tap arithmetic; C4 / C5 single-GPU lines
mkdir -p gpurun_out
O=gpurun_out
echo "== conv engine tests"; timeout -s KILL 900 python -m pytest tests/test_gpu_conv_tc.py -q -x -p no:cacheprovider 2>&1 | tail -4 | cut -c1-300
echo "== tests"; timeout -s KILL 1500 python -m pytest tests/ -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_conv_tc.py > $O/r2g_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2g_tests.txt | cut -c1-300
for V in 1 0; do
echo "== bench C2 (NEMAR_K7_XTAPS=$V)"; NEMAR_K7_XTAPS=$V timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --torch_gpu_reference 0 2>$O/r2g_bench$V.err | tail -1 > $O/r2g_bench$V.json; python - <<PY
import json
d=json.load(open('gpurun_out/r2g_bench$V.json'))
print({k:d.get(k) for k in ('value','ms_per_step','e2e','grid_sample','gpu_launches')})
r=d.get('roofline',{})
for k,v in r.get('by_kernel',{}).items(): print(k, round(v['ms'],2), v['n'], v['tflops'], {a:round(b,2) for a,b in v['top'].items()})
PY
done
echo "== bench C4 (1 GPU)"; timeout -s KILL 600 python bench.py --workload C4 --steps 5 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --torch_gpu_reference 0 2>$O/r2g_c4.err | tail -1 > $O/r2g_c4.json; cut -c1-700 $O/r2g_c4.json; tail -2 $O/r2g_c4.err
echo "== bench C5 (1 GPU)"; timeout -s KILL 600 python bench.py --workload C5 --steps 5 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --torch_gpu_reference 0 2>$O/r2g_c5.err | tail -1 > $O/r2g_c5.json; cut -c1-700 $O/r2g_c5.json; tail -2 $O/r2g_c5.err
