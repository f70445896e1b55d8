#!/bin/bash
# Round 2: ring-kernel build — op/model parity subset, nbench of the new defaults, and the step time by ring depth
mkdir -p gpurun_out
O=gpurun_out
echo "== tests (ops, model, conv defaults)"; timeout -s KILL 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_fidelity.py -m gpu -q -x -p no:cacheprovider > $O/r2p_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2p_tests.txt | cut -c1-300
echo "== nbench"; timeout -s KILL 300 python scripts/nbench.py --by_variant --reps 10 --shapes res256 res256r up128 head64 stn32 --variants "" "NEMAR_LEAN_PIPE_RED=2 NEMAR_LEAN_PIPE_APPLY=2" "NEMAR_LEAN_PIPE_RED=4 NEMAR_LEAN_PIPE_FWD=3" > $O/r2p_nbench.txt 2>&1; cut -c1-220 $O/r2p_nbench.txt
B="python bench.py --gpus 1 --steps 20 --warmup 5 --no_cpu_baseline --grid_sample_bench 0 --torch_gpu_reference 0"
echo "== bench default"; timeout -s KILL 300 $B 2>$O/r2p_bench.err | tail -1 > $O/r2p_bench.json; cut -c1-220 $O/r2p_bench.json; tail -2 $O/r2p_bench.err
echo "== bench PIPE=3"; NEMAR_LEAN_PIPE=3 timeout -s KILL 300 $B --kernel_timing 0 2>/dev/null | tail -1 | cut -c1-200
echo "== bench PIPE=2"; NEMAR_LEAN_PIPE=2 timeout -s KILL 300 $B --kernel_timing 0 2>/dev/null | tail -1 | cut -c1-200
echo "== bench PIPE=4"; NEMAR_LEAN_PIPE=4 timeout -s KILL 300 $B --kernel_timing 0 2>/dev/null | tail -1 | cut -c1-200
