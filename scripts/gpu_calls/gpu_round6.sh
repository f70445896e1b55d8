#!/bin/bash
# multi-tile gather CTAs + occupancy-aware wgrad splits: correctness under forced settings, then A/B timing
mkdir -p gpurun_out
echo "== tc default"; timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q 2>&1 | tail -3
echo "== tc TPC=8"; NEMAR_TC_TPC=8 timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q 2>&1 | tail -3
echo "== tc TPC=2 OCC=2"; NEMAR_TC_TPC=2 NEMAR_WG_OCC=2 timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q 2>&1 | tail -3
echo "== model"; timeout 600 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 2"
run() { echo "== $1"; shift; env "$@" timeout 300 $B 2>/dev/null | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.readline()); k = (r.get('roofline') or {}).get('by_kernel', {})
print('ms/step', r['ms_per_step'], {a: round(b['ms'] / r['steps'], 2) for a, b in list(k.items())[:9]})"; }
run legacy NEMAR_TC_TPC=1 NEMAR_WG_LEGACY_SPLITS=1 NEMAR_WG_OCC=1
run wgrad_only NEMAR_TC_TPC=1
run new X=1
run new_waves1 NEMAR_TC_TPC_WAVES=1
run new_waves2 NEMAR_TC_TPC_WAVES=2
