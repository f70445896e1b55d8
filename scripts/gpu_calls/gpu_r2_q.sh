#!/bin/bash
# Round 2: apply pass with the mirrored tap in the ring, 8-warp statistics finalize, wgrad-pair residency A/B
mkdir -p gpurun_out
O=gpurun_out
echo "== tests (ops, model)"; timeout -s KILL 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_conv_tc.py -m gpu -q -x -p no:cacheprovider -k "not tiled_kernels and not pair_mode" > $O/r2q_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2q_tests.txt | cut -c1-300
echo "== nbench"; timeout -s KILL 300 python scripts/nbench.py --by_variant --reps 10 --shapes res256 res256r up128 stn32 --variants "" "NEMAR_LEAN_PIPE_APPLY=2" "NEMAR_LEAN_PIPE_APPLY=4" > $O/r2q_nbench.txt 2>&1; cut -c1-220 $O/r2q_nbench.txt
echo "== kbench wgrad pair"; timeout -s KILL 300 python scripts/kbench.py --layers resblock d512 --variants "" "NEMAR_WG_PAIR_OCC=1 NEMAR_WG_PAIR_STAGES=6" "NEMAR_WG_PAIR_OCC=1 NEMAR_WG_PAIR_STAGES=4" "NEMAR_WG_PAIR_STAGES=2" > $O/r2q_kbench.txt 2>&1; cut -c1-220 $O/r2q_kbench.txt
B="python bench.py --gpus 1 --steps 20 --warmup 5 --no_cpu_baseline --grid_sample_bench 0 --torch_gpu_reference 0 --kernel_timing 0"
echo "== bench default"; timeout -s KILL 300 $B 2>/dev/null | tail -1 | cut -c1-200
echo "== bench default again"; timeout -s KILL 300 $B 2>/dev/null | tail -1 | cut -c1-200
echo "== bench WG occ1 st6"; NEMAR_WG_PAIR_OCC=1 NEMAR_WG_PAIR_STAGES=6 timeout -s KILL 300 $B 2>/dev/null | tail -1 | cut -c1-200
