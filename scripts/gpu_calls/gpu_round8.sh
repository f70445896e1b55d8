#!/bin/bash
mkdir -p gpurun_out
echo "== tc default"; timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q 2>&1 | tail -3
echo "== tc RING 40 TPC 4"; NEMAR_TC_RING_KB=40 NEMAR_TC_TPC=4 timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q 2>&1 | tail -3
echo "== model"; timeout 600 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 2 --top 60"
run() { n=$1; echo "== $n"; shift; env "$@" timeout 300 $B 2>/dev/null | tail -1 > gpurun_out/r8_$n.json; python -c "
import sys, json
r = json.load(open('gpurun_out/r8_$n.json')); k = (r.get('roofline') or {}).get('by_kernel', {})
print('ms/step', r['ms_per_step'], {a: round(b['ms'] / r['steps'], 2) for a, b in list(k.items())[:9]})"; }
run default X=1
run legacy_tiles NEMAR_TC_LEGACY_TILES=1
run ring48 NEMAR_TC_RING_KB=48
run ring64 NEMAR_TC_RING_KB=64
run wgocc6 NEMAR_WG_OCC_MAX=6 NEMAR_WG_MIN_STAGES=3
run wgst3 NEMAR_WG_MIN_STAGES=3
run wgst6 NEMAR_WG_MIN_STAGES=6
