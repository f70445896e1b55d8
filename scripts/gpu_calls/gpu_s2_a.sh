#!/bin/bash
# Session-2 call A (1 GPU): hardware probe, CTA-pair kernels (region debug -> pytest), A/B benches of the opt-in paths.
mkdir -p gpurun_out
O=gpurun_out
echo "== probe"; timeout -s KILL 60 scripts/probe/umma_shift_probe.bin > $O/probe.txt 2>&1; echo "rc=$? match=$(grep -c MATCH $O/probe.txt) mismatch=$(grep -c mismatch $O/probe.txt)"
echo "== pair gather (case 2)"; PAIR_MODE=2 timeout -s KILL 150 python scripts/pair_debug.py 2 > $O/pair_gather.txt 2>&1; G=$?; echo "rc=$G"; tail -12 $O/pair_gather.txt | cut -c1-300
echo "== pair wgrad (case 2)"; PAIR_MODE=3 timeout -s KILL 150 python scripts/pair_debug.py 2 > $O/pair_wgrad.txt 2>&1; W=$?; echo "rc=$W"; tail -16 $O/pair_wgrad.txt | cut -c1-300
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
PM=0
if [ $G -eq 0 ] && [ $W -eq 0 ]; then PM=1; elif [ $G -eq 0 ]; then PM=2; elif [ $W -eq 0 ]; then PM=3; fi
echo "usable pair mode: $PM"
if [ $PM -ne 0 ]; then
  echo "== pair cases (mode $PM)"
  NEMAR_TC_PAIR=$PM timeout -s KILL 400 python scripts/tc_check.py 2 5 6 9 10 23 24 25 26 2>&1 | cut -c1-330 | tee $O/pair_cases.txt
fi
B="python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 1 --top 6"
run() { n=$1; echo "== bench $n"; shift; env "$@" timeout -s KILL 300 $B $EXTRA 2>$O/a_$n.err | tail -1 > $O/a_$n.json; python - <<PY
import json
try:
    r = json.load(open('$O/a_$n.json')); k = (r.get('roofline') or {}).get('by_kernel', {})
    print('ms/step', r['ms_per_step'], 'e2e', r['e2e']['ms_per_step'], 'launches', r['gpu_launches'])
    for a, b in k.items():
        print('   ', a, round(b['ms'] / r['steps'], 2), 'ms/step', b['tflops'], 'TF/s', {x: round(y / r['steps'], 2) for x, y in list(b['top'].items())[:3]})
except Exception as e:
    print('bench $n failed:', e); print(open('$O/a_$n.err').read()[-1500:])
PY
}
EXTRA="" run default X=1
if [ $PM -ne 0 ]; then
  EXTRA="" run pair NEMAR_TC_PAIR=$PM
  if [ $G -eq 0 ]; then EXTRA="" run pair_deep NEMAR_TC_PAIR=$PM NEMAR_TC_PAIR_STAGES=6 NEMAR_TC_PAIR_TPC=2; fi
  if [ $W -eq 0 ]; then EXTRA="" run pair_wg3 NEMAR_TC_PAIR=$PM NEMAR_WG_PAIR_STAGES=3; fi
fi
EXTRA="--batch_d 1" run batchd X=1
EXTRA="--cuda_graph 1" run graph X=1; grep -i "capture" $O/a_graph.err | head -3
echo "== batch_d equivalence test"; timeout -s KILL 400 python -m pytest tests/test_gpu_model.py -q -x -k batched -p no:cacheprovider 2>&1 | tail -5
